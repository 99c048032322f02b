"""TEST INFRASTRUCTURE ONLY -- ctypes access to the reference's own compiled KREG code.

``oracle/_ref/KREG.so`` is a byte-for-byte copy of the reference's prebuilt
``mlatom/fortran/np2/KREG.so`` (ifort 2019, ``mlatom/fortran/compile.py:7-8``) made by
``oracle/Makefile``.  Its f2py Python module targets another CPython ABI and is never
imported; instead the raw Fortran entry points ``kreg_mp_<name>_`` (module ``kreg`` of
``mlatom/fortran/KREG.f90``) are called directly, every argument by reference,
arrays in Fortran order, ``logical`` = 4-byte int with ifort's true = -1.

Only ``tests/``, ``__graft_entry__.smoke()``, ``tools/`` (fixture generation) and the
CPU-baseline legs of ``bench.py`` may import this module.  The product (``mlatom_b200``)
never does.
"""
from __future__ import annotations

import ctypes
import glob
import os
from ctypes import POINTER, byref, c_double, c_int

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_lib = None


def available() -> bool:
    return os.path.isfile(os.path.join(_REF_DIR, "KREG.so")) and os.path.isfile(
        os.path.join(_REF_DIR, "libiomp5.so")
    )


def _find_openblas() -> str:
    import scipy

    root = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    hits = sorted(glob.glob(os.path.join(root, "libscipy_openblas*.so")))
    if not hits:
        raise RuntimeError("SciPy's bundled OpenBLAS not found under " + root)
    return hits[0]


def load():
    """dlopen the shims (by path, so LD_LIBRARY_PATH is not needed) and then KREG.so."""
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref is not built; run `make -C oracle ref` where /root/reference exists")
    os.environ.setdefault("KREG_REF_LAPACK", _find_openblas())
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    for name in ("libiomp5.so", "libmkl_core.so", "libmkl_intel_thread.so", "libmkl_intel_lp64.so"):
        ctypes.CDLL(os.path.join(_REF_DIR, name), mode=ctypes.RTLD_GLOBAL)
    _lib = ctypes.CDLL(os.path.join(_REF_DIR, "KREG.so"), mode=ctypes.RTLD_GLOBAL)
    return _lib


def set_threads(n: int) -> None:
    """The pthread stand-in reads OMP_NUM_THREADS at every parallel region."""
    os.environ["OMP_NUM_THREADS"] = str(int(n))
    os.environ["OPENBLAS_NUM_THREADS"] = str(int(n))


def _ci(v):
    return byref(c_int(int(v)))


def _cd(v):
    return byref(c_double(float(v)))


def _p(a, t=c_double):
    return a.ctypes.data_as(POINTER(t))


def _logical(v):
    return byref(c_int(-1 if v else 0))


def xyz_to_f(xyz: np.ndarray) -> np.ndarray:
    """(N, Nat, 3) C-order -> Fortran XYZ(3, Nat, N); the C-order bytes already are that layout."""
    return np.ascontiguousarray(xyz, dtype=np.float64)


def get_ac2darray(natoms: int) -> np.ndarray:
    """kreg_mp_get_ac2darray_ (KREG.f90:34-52). Returns int32 (Nat,Nat) (symmetric)."""
    lib = load()
    xsize = natoms * (natoms - 1) // 2
    ac2d = np.zeros((natoms, natoms), dtype=np.int32, order="F")
    lib.kreg_mp_get_ac2darray_(_ci(natoms), _ci(xsize), _p(ac2d, c_int))
    return ac2d


def get_xeq(req: np.ndarray) -> np.ndarray:
    """kreg_mp_get_xeq_ (KREG.f90:223-239). req: (Nat,3)."""
    lib = load()
    natoms = req.shape[0]
    xsize = natoms * (natoms - 1) // 2
    ac2d = get_ac2darray(natoms)
    reqf = np.ascontiguousarray(req, dtype=np.float64)  # bytes == F(3,Nat)
    xeq = np.zeros(xsize)
    lib.kreg_mp_get_xeq_(_ci(natoms), _ci(xsize), _p(ac2d, c_int), _p(reqf), _p(xeq))
    return xeq


def calc_re_descriptors(xyz: np.ndarray, xeq: np.ndarray) -> np.ndarray:
    """kreg_mp_calc_re_descriptors_ (KREG.f90:257-271). xyz (N,Nat,3) -> X (N,D)."""
    lib = load()
    n, natoms, _ = xyz.shape
    xsize = natoms * (natoms - 1) // 2
    ac2d = get_ac2darray(natoms)
    xyzf = xyz_to_f(xyz)
    x = np.zeros((n, xsize))  # bytes == F(D,N)
    lib.kreg_mp_calc_re_descriptors_(
        _ci(xsize), _ci(natoms), _ci(n), _p(ac2d, c_int), _p(xyzf), _p(x), _p(np.ascontiguousarray(xeq))
    )
    return x


def _itrgrxyz(ntrgr: int, ntrain: int, natoms: int):
    """get_itrgrxyz (KREG.f90:273-291) -- or the trivial (1,3) zero array when there are no gradients."""
    lib = load()
    if ntrgr == 0:
        return np.zeros((1, 3), dtype=np.int32, order="F"), 1
    it = np.zeros((ntrgr, 3), dtype=np.int32, order="F")
    lib.kreg_mp_get_itrgrxyz_(_ci(ntrgr), _ci(ntrain), _ci(natoms), _p(it, c_int))
    return it, ntrgr


def calc_kernel_matrix(xyz, x, sigma, ntrval, ntrgr) -> np.ndarray:
    """kreg_mp_calc_kernel_matrix_ (KREG.f90:293-346). Returns full symmetric K (M,M)."""
    lib = load()
    n, natoms, _ = xyz.shape
    xsize = x.shape[1]
    m = ntrval + ntrgr
    it, itdim = _itrgrxyz(ntrgr, n, natoms)
    ac2d = get_ac2darray(natoms)
    k = np.zeros((m, m), order="F")
    lib.kreg_mp_calc_kernel_matrix_(
        _ci(natoms), _ci(xsize), _ci(n), _ci(ntrval), _ci(ntrgr), _ci(itdim), _p(it, c_int),
        _p(ac2d, c_int), _p(xyz_to_f(xyz)), _p(np.ascontiguousarray(x)), _p(k), _cd(sigma),
    )
    return k


def train(xyz, x, ytrain, sigma, lambdav, lambdagradxyz, ntrval, ntrgr, calc_kernel=True):
    """kreg_mp_train_ (KREG.f90:495-532). Returns (K, alpha)."""
    lib = load()
    n, natoms, _ = xyz.shape
    xsize = x.shape[1]
    m = ntrval + ntrgr
    ac2d = get_ac2darray(natoms)
    k = np.zeros((m, m), order="F")
    alpha = np.zeros(m)
    y = np.ascontiguousarray(ytrain, dtype=np.float64)
    assert y.shape == (m,)
    lib.kreg_mp_train_(
        _ci(natoms), _ci(xsize), _ci(n), _ci(ntrval), _ci(ntrgr), _p(ac2d, c_int), _p(xyz_to_f(xyz)),
        _p(np.ascontiguousarray(x)), _p(y), _p(k), _p(alpha), _cd(sigma), _cd(lambdav),
        _cd(lambdagradxyz), _logical(calc_kernel),
    )
    return k, alpha


def predict(xyz_tr, x_tr, xyz_te, x_te, alpha, sigma, ntrval, ntrgr, calc_val=True, calc_grad=True, calc_hess=False):
    """kreg_mp_predict_ (KREG.f90:534-596). Returns (Yest (Np,), Ygrad (Np,Nat,3)) -- plus Yhess (Np,3Nat,3Nat) when
    calc_hess; no prior added."""
    lib = load()
    ntr, natoms, _ = xyz_tr.shape
    npred = xyz_te.shape[0]
    xsize = x_tr.shape[1]
    ac2d = get_ac2darray(natoms)
    yest = np.zeros(npred)
    ygrad = np.zeros((npred, natoms, 3))  # bytes == F(3,Nat,Np)
    yhess = np.zeros((npred, 3 * natoms, 3 * natoms)) if calc_hess else np.zeros(1)  # bytes == F(3Nat,3Nat,Np)
    a = np.ascontiguousarray(alpha, dtype=np.float64).reshape(-1)
    assert a.shape[0] == ntrval + ntrgr
    lib.kreg_mp_predict_(
        _ci(natoms), _ci(xsize), _ci(ntr), _ci(npred), _ci(ntrval), _ci(ntrgr), _p(ac2d, c_int),
        _p(np.ascontiguousarray(x_tr)), _p(xyz_to_f(xyz_tr)), _p(np.ascontiguousarray(x_te)),
        _p(xyz_to_f(xyz_te)), _p(a), _logical(calc_val), _logical(calc_grad), _logical(calc_hess),
        _p(yest), _p(ygrad), _p(yhess), _cd(sigma),
    )
    if calc_hess:
        return yest, ygrad, yhess
    return yest, ygrad
