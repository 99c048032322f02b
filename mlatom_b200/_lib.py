"""ctypes binding of libkreg_b200.so (include/kreg_b200.h).

There is no fallback of any kind: if the shared library is missing or the device is not
an sm_100 part, the first call raises.  Nothing here imports anything from ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# KREG_LIB selects another build of the same ABI (the bounds-checking debug library, libkreg_b200_dbg.so)
LIB_PATH = os.environ.get("KREG_LIB") or os.path.join(_HERE, "libkreg_b200.so")

KREG_OK = 0
KREG_E_BADARG = -1
KREG_E_UNSUPPORTED = -2
KREG_E_WORKSPACE = -3
KREG_E_CUDA = -4
FILL_LOWER = 0
FILL_FULL = 1
PREDICT_ENERGY = 1
PREDICT_GRADIENT = 2
FUSED_MAX_NATOMS = 24  # KREG_FUSED_MAX_NATOMS: largest molecule of the register-resident prediction kernels

_ERRORS = {
    KREG_E_BADARG: "bad argument",
    KREG_E_UNSUPPORTED: "unsupported (Natoms outside the compiled set, or not an sm_100 device)",
    KREG_E_WORKSPACE: "workspace too small",
    KREG_E_CUDA: "CUDA / cuSOLVER failure",
}

# every symbol include/kreg_b200.h declares: (restype, argtypes)
PROTOTYPES = {
    "kreg_abi_version": (c_int, []),
    "kreg_last_error": (c_char_p, []),
    "kreg_device_check": (c_int, []),
    "kreg_debug_bounds_failures": (ctypes.c_longlong, [POINTER(c_int), POINTER(c_int)]),
    "kreg_stream_synchronize": (c_int, [c_void_p]),
    "kreg_measure_fp64_peak": (c_int, [c_int, c_int, c_double, POINTER(c_double)]),
    "kreg_xeq": (c_int, [c_int, c_void_p, c_void_p, c_void_p]),
    "kreg_descriptors": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kreg_build_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64]),
    "kreg_build_workspace_bytes_rows": (c_size_t, [c_int, c_int64, c_int64, c_int64, c_int64]),
    "kreg_build_kernel_matrix": (
        c_int,
        [c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_double, c_double, c_double, c_int64, c_int64,
         c_int, c_void_p, c_int64, c_void_p, c_size_t, c_void_p],
    ),
    "kreg_kernel_block": (
        c_int,
        [c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_double, c_void_p, c_int64, c_void_p, c_size_t,
         c_void_p],
    ),
    "kreg_kernel_block_ex_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64]),
    "kreg_kernel_block_ex": (
        c_int,
        [c_int, c_int, c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_double, c_int, c_double, c_void_p,
         c_int64, c_void_p, c_size_t, c_void_p],
    ),
    "kreg_radial_packed_model_bytes": (c_size_t, [c_int, c_int64]),
    "kreg_pack_model_radial": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kreg_predict_radial_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64]),
    "kreg_predict_radial": (
        c_int,
        [c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int64, c_void_p, c_int,
         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "kreg_block_dot": (c_int, [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_double, c_void_p, c_void_p]),
    "kreg_laplacian_predict": (
        c_int,
        [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int64, c_void_p, c_void_p, c_int64, c_int,
         c_void_p, c_void_p, c_void_p],
    ),
    "kreg_perm_descriptors": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "kreg_perm_self_terms": (
        c_int, [c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "kreg_perm_kernel_block_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int]),
    "kreg_perm_kernel_block": (
        c_int,
        [c_int, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_double, c_int, c_double, c_void_p,
         c_int64, c_void_p, c_size_t, c_void_p],
    ),
    "kreg_perm_scale_alpha": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kreg_perm_finish_prediction": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "kreg_solve_workspace_bytes": (c_int, [c_int64, POINTER(c_size_t), POINTER(c_size_t)]),
    "kreg_solve": (c_int, [c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "kreg_solve_indefinite_workspace_bytes": (c_int, [c_int64, c_int, POINTER(c_size_t), POINTER(c_size_t)]),
    "kreg_solve_indefinite": (
        c_int, [c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "kreg_precompute_v": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kreg_packed_model_bytes": (c_size_t, [c_int, c_int64, c_int]),
    "kreg_pack_model": (
        c_int,
        [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
    "kreg_pack_model_x": (
        c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kreg_md_first_half": (c_int, [c_int, c_int64, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kreg_md_second_half": (
        c_int, [c_int, c_int64, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kreg_md_uncertainty": (
        c_int, [c_int64, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "kreg_hessian": (c_int, [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_double, c_int64, c_void_p, c_void_p, c_void_p]),
    "kreg_predict_tile_rows": (c_int, [c_int, c_int]),
    "kreg_predict_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int64]),
    "kreg_predict": (
        c_int,
        [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_int64, c_void_p, c_int,
         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    ),
}


class KregError(RuntimeError):
    def __init__(self, fn: str, code: int, detail: str = ""):
        self.function, self.code = fn, code
        if code > 0:
            msg = f"{fn}: factorisation failed, leading minor {code} is not positive definite"
        else:
            msg = f"{fn}: {_ERRORS.get(code, 'error ' + str(code))}"
        if detail:
            msg += f" [{detail}]"
        super().__init__(msg)


_lib = None


def load() -> ctypes.CDLL:
    """dlopen libkreg_b200.so and set every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C mlatom_b200/csrc). mlatom_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(fn: str, rc: int) -> None:
    if rc != KREG_OK:
        detail = ""
        if rc == KREG_E_CUDA:
            detail = (load().kreg_last_error() or b"").decode(errors="replace")
        raise KregError(fn, rc, detail)


def call(fn: str, *args) -> None:
    check(fn, getattr(load(), fn)(*args))
