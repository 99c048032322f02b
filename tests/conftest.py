import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("natoms", "ntrain", "ntest", "ntrval", "ntrgr"):
        d[k] = int(d[k])
    for k in ("sigma", "lambdav", "lambdagradxyz", "prior"):
        d[k] = float(d[k])
    return d


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


@pytest.fixture(scope="session")
def oracle():
    from oracle import kreg_oracle

    kreg_oracle.load()
    return kreg_oracle


def reference_mlatom_data():
    """The reference's own mlatom.data module (build container only; h5py is absent here and only needed for
    features these tests do not touch, so an empty stand-in module satisfies the import)."""
    import os
    import sys
    import types

    if not os.path.isdir("/root/reference/mlatom"):
        return None
    os.environ["MLATOM_NO_VERSION_CHECK"] = "1"
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    try:
        import mlatom.data as ref_data
    except Exception:  # pragma: no cover - reference not importable here
        return None
    return ref_data



def pytest_sessionfinish(session, exitstatus):
    """With the bounds-checking build (KREG_LIB=.../libkreg_b200_dbg.so) the in-kernel failure counter must be zero at
    the end of the session (compute-sanitizer is closed on this GPU pool)."""
    import ctypes
    import os

    if not os.environ.get("KREG_LIB", "").endswith("_dbg.so"):
        return
    try:
        from mlatom_b200 import _lib

        code, line = ctypes.c_int(0), ctypes.c_int(0)
        n = _lib.load().kreg_debug_bounds_failures(ctypes.byref(code), ctypes.byref(line))
    except Exception as err:  # pragma: no cover
        print(f"\nkreg_debug_bounds_failures could not be read: {err}")
        return
    print(f"\nkreg_debug_bounds_failures = {n} (first code {code.value}, source line {line.value})")
    if n > 0:
        session.exitstatus = 1
