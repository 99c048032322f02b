"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/kreg_oracle.c).

The oracle restates the reference's KREG algorithm (mlatom/fortran/KREG.f90 and the
Gaussian/RE/unpermuted branch of mlatom/MLatomF_src/A_KRR_kernel.f90, A_KRR.f90) on the
CPU.  It is pinned against the reference's own compiled library (oracle/_ref/KREG.so, see
oracle/ref_kreg.py) by tests/test_oracle_vs_ref.py and by the golden vectors in
tests/golden/ (generated from that library by tools/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
module; the product package mlatom_b200 never does.

The linear solve is LAPACK dpotrf/dpotrs('U') exactly as mlatom/fortran/mathUtils.f90:65,79,
through SciPy (OpenBLAS).  The reference links MKL; no reference test pins alpha at that
boundary, so alpha parity is *unpinned* (checked through residuals and fixed-alpha
predictions instead).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_long

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkreg_oracle.so")
_lib = None
_DP = POINTER(c_double)


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "kreg_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libkreg_oracle.so"], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.kro_num_threads.restype = c_int
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_DP)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_threads(n: int) -> None:
    load().kro_set_num_threads(c_int(int(n)))


def num_threads() -> int:
    return int(load().kro_num_threads())


def ac2d(nat: int) -> np.ndarray:
    out = np.zeros((nat, nat), dtype=np.int32)
    load().kro_ac2d(c_int(nat), out.ctypes.data_as(POINTER(c_int)))
    return out


def xeq(req: np.ndarray) -> np.ndarray:
    req = _c(req)
    nat = req.shape[0]
    out = np.zeros(nat * (nat - 1) // 2)
    load().kro_xeq(c_int(nat), _d(req), _d(out))
    return out


def descriptors(xyz: np.ndarray, xeq_: np.ndarray) -> np.ndarray:
    xyz = _c(xyz)
    n, nat, _ = xyz.shape
    out = np.zeros((n, nat * (nat - 1) // 2))
    load().kro_descriptors(c_int(nat), c_long(n), _d(xyz), _d(_c(xeq_)), _d(out))
    return out


def kernel_matrix(xyz, x, sigma, ntrval, ntrgr, factored=False) -> np.ndarray:
    xyz, x = _c(xyz), _c(x)
    n, nat, _ = xyz.shape
    m = ntrval + ntrgr
    k = np.zeros((m, m))
    fn = load().kro_kernel_matrix_factored if factored else load().kro_kernel_matrix
    fn(c_int(nat), c_int(n), c_int(ntrval), c_int(ntrgr), _d(xyz), _d(x), c_double(sigma), _d(k))
    return k


def solve(k: np.ndarray, y: np.ndarray, ntrval: int, ntrgr: int, lambdav: float, lambdag: float):
    """calcAlpha (KREG.f90:598-627): copy K, shift the two diagonal blocks, dpotrf/dpotrs 'U'."""
    from scipy.linalg import cho_factor, cho_solve

    a = np.array(k, dtype=np.float64, order="F", copy=True)
    load().kro_add_lambda(c_int(ntrval), c_int(ntrgr), c_double(lambdav), c_double(lambdag), _d(a))
    c, low = cho_factor(a, lower=False, overwrite_a=True, check_finite=False)
    return cho_solve((c, low), _c(y), check_finite=False)


def ytrain(y=None, ygrad=None, prior=0.0) -> np.ndarray:
    """kreg_api.py:51-72: [Y - prior ; gradients flattened (point, atom, xyz)]."""
    parts = []
    if y is not None:
        parts.append(np.asarray(y, dtype=np.float64) - prior)
    if ygrad is not None:
        parts.append(np.asarray(ygrad, dtype=np.float64).reshape(-1))
    return np.concatenate(parts)


def predict_kregf90(xyz_tr, x_tr, alpha, sigma, ntrval, ntrgr, xyz_te, x_te, calc_val=True, calc_grad=True):
    xyz_tr, x_tr, xyz_te, x_te, alpha = map(_c, (xyz_tr, x_tr, xyz_te, x_te, alpha))
    ntr, nat, _ = xyz_tr.shape
    npred = xyz_te.shape[0]
    e = np.zeros(npred)
    g = np.zeros((npred, nat, 3))
    load().kro_predict_kregf90(
        c_int(nat), c_int(ntr), c_int(ntrval), c_int(ntrgr), _d(xyz_tr), _d(x_tr), _d(alpha.reshape(-1)),
        c_double(sigma), c_long(npred), _d(xyz_te), _d(x_te), c_int(calc_val), c_int(calc_grad), _d(e), _d(g),
    )
    return e, g


def predict_mlatomf(xyz_tr, x_tr, alpha, sigma, prior, ntrval, ntrgr, xyz_te, x_te, calc_val=True, calc_grad=True):
    xyz_tr, x_tr, xyz_te, x_te, alpha = map(_c, (xyz_tr, x_tr, xyz_te, x_te, alpha))
    ntr, nat, _ = xyz_tr.shape
    npred = xyz_te.shape[0]
    e = np.zeros(npred)
    g = np.zeros((npred, nat, 3))
    load().kro_predict_mlatomf(
        c_int(nat), c_int(ntr), c_int(ntrval), c_int(ntrgr), _d(xyz_tr), _d(x_tr), _d(alpha.reshape(-1)),
        c_double(sigma), c_double(prior), c_long(npred), _d(xyz_te), _d(x_te), c_int(calc_val),
        c_int(calc_grad), _d(e), _d(g),
    )
    return e, g


def precompute_v(xyz_tr, x_tr, alpha_g):
    xyz_tr, x_tr = _c(xyz_tr), _c(x_tr)
    ntr, nat, _ = xyz_tr.shape
    ag = _c(alpha_g).reshape(ntr, nat, 3)
    v = np.zeros_like(x_tr)
    load().kro_precompute_v(c_int(nat), c_int(ntr), _d(xyz_tr), _d(x_tr), _d(ag), _d(v))
    return v


def predict_factored(x_tr, alpha_v, v, sigma, prior, xyz_te, x_te, calc_grad=True, want_scale=False):
    x_tr, xyz_te, x_te = _c(x_tr), _c(xyz_te), _c(x_te)
    ntr = x_tr.shape[0]
    npred, nat, _ = xyz_te.shape
    av = None if alpha_v is None else _c(alpha_v).reshape(-1)
    vv = None if v is None else _c(v)
    e = np.zeros(npred)
    g = np.zeros((npred, nat, 3))
    s = np.zeros(npred) if want_scale else None
    load().kro_predict_factored(
        c_int(nat), c_int(ntr), _d(x_tr), _d(av), _d(vv), c_double(sigma), c_double(prior), c_long(npred),
        _d(xyz_te), _d(x_te), c_int(calc_grad), _d(e), _d(g), _d(s),
    )
    return (e, g, s) if want_scale else (e, g)


def hessian_kregf90(x_tr, alpha, sigma, ntrval, xyz_te, x_te):
    """Literal KREG.f90 Hessian (values part of alpha only). Returns (Np, 3Nat, 3Nat)."""
    x_tr, xyz_te, x_te = _c(x_tr), _c(xyz_te), _c(x_te)
    npred, nat, _ = xyz_te.shape
    a = _c(alpha).reshape(-1)
    h = np.zeros((npred, 3 * nat, 3 * nat))
    load().kro_hessian_kregf90(c_int(nat), c_int(ntrval), _d(x_tr), _d(a), c_double(sigma), c_long(npred), _d(xyz_te),
                               _d(x_te), _d(h))
    return h


KERNEL_KINDS = {"gaussian": 0, "exponential": 1, "matern": 2, "laplacian": 3}


def kernel_block_generic(kind: str, x_a, x_b, sigma: float, nn: int = 2) -> np.ndarray:
    """MLatomF's Kfunk (A_KRR_kernel.f90:806-851) between two descriptor sets: (na, D) x (nb, D) -> (na, nb).
    Parity unpinned (no executable MLatomF here); restated literally from the Fortran."""
    x_a, x_b = _c(x_a), _c(x_b)
    out = np.zeros((x_a.shape[0], x_b.shape[0]))
    load().kro_kernel_block_generic(c_int(KERNEL_KINDS[kind.lower()]), c_int(nn), c_int(x_a.shape[1]), c_long(x_a.shape[0]),
                                    _d(x_a), c_long(x_b.shape[0]), _d(x_b), c_double(sigma), _d(out))
    return out


# ---- permutationally invariant kernel (MLatomF molDescrType=permuted permInvKernel) -- PARITY UNPINNED ----------------
def perm_atom_maps(natoms: int, groups) -> np.ndarray:
    """Full 0-based atom maps of every permutation of the atoms inside each group (MLatomF ``permInvNuclei=1-2.5-6``:
    groups separated by '.', atoms by '-'; Nperm = prod (group size)!, dataset.f90:1952-1977).  A map follows
    permuteAtoms / getPermIatoms (molDescr.f90:197-220): walking the atoms in order, the k-th atom that belongs to the
    permuted set takes the k-th entry of the permuted index list.  Row 0 is the identity."""
    import itertools

    groups = [sorted(int(a) for a in g) for g in groups]
    maps = []
    for combo in itertools.product(*[itertools.permutations(g) for g in groups]):
        perm_list = [a for g in combo for a in g]          # permIatomsNucl(iperm): the groups' permuted atoms in sequence
        members = set(perm_list)
        m, icount = list(range(natoms)), 0
        for ii in range(natoms):
            if ii in members:
                m[ii] = perm_list[icount]
                icount += 1
        maps.append(m)
    return np.asarray(maps, dtype=np.int32)


def inverse_atom_maps(pmap: np.ndarray) -> np.ndarray:
    """mod_inversePermInds (A_KRR_kernel.f90:944-950)."""
    inv = np.empty_like(pmap)
    for p in range(pmap.shape[0]):
        inv[p, pmap[p]] = np.arange(pmap.shape[1], dtype=pmap.dtype)
    return inv


def perm_descriptors(xyz: np.ndarray, req: np.ndarray, pmap: np.ndarray) -> np.ndarray:
    """Permuted RE descriptor (calcX, molDescr.f90:128-160): (N, Nperm * D), block P = RE descriptor of permutation P."""
    xyz, req = _c(xyz), _c(req)
    pmap = np.ascontiguousarray(pmap, dtype=np.int32)
    n, nat, _ = xyz.shape
    out = np.zeros((n, pmap.shape[0] * nat * (nat - 1) // 2))
    load().kro_perm_descriptors(c_int(nat), c_long(n), _d(xyz), _d(req), c_int(pmap.shape[0]),
                                pmap.ctypes.data_as(ctypes.c_void_p), _d(out))
    return out


def perm_kernel_block(x_a: np.ndarray, x_b: np.ndarray, nperm: int, sigma: float) -> np.ndarray:
    """Normalised permutationally invariant Gaussian kernel (kernelFunction, A_KRR_kernel.f90:773-804)."""
    x_a, x_b = _c(x_a), _c(x_b)
    out = np.zeros((x_a.shape[0], x_b.shape[0]))
    load().kro_perm_kernel_block(c_int(x_a.shape[1] // nperm), c_int(nperm), c_long(x_a.shape[0]), _d(x_a),
                                 c_long(x_b.shape[0]), _d(x_b), c_double(sigma), _d(out))
    return out


def perm_predict(x_tr, alpha, sigma, prior, xyz_te, x_te, pmap, calc_grad=True):
    """Yest_KRR and the permInvKernel branch of YgradXYZest_KRR (A_KRR_kernel.f90:225-245, 414-503), values-trained."""
    x_tr, x_te, xyz_te, a = _c(x_tr), _c(x_te), _c(xyz_te), _c(alpha).reshape(-1)
    inv = np.ascontiguousarray(inverse_atom_maps(np.asarray(pmap)), dtype=np.int32)
    npred, nat, _ = xyz_te.shape
    e, g = np.zeros(npred), np.zeros((npred, nat, 3))
    load().kro_perm_predict(c_int(nat), c_int(inv.shape[0]), inv.ctypes.data_as(ctypes.c_void_p), c_long(x_tr.shape[0]),
                            _d(x_tr), _d(a), c_double(sigma), c_double(prior), c_long(npred), _d(xyz_te), _d(x_te),
                            c_int(1 if calc_grad else 0), _d(e), _d(g))
    return e, (g if calc_grad else None)


def predict_generic(kind: str, x_tr, alpha, sigma, prior, xyz_te, x_te, nn: int = 2, calc_grad=True):
    """Values-trained model on any of MLatomF's kernel functions: Yest_KRR (A_KRR_kernel.f90:225-245) and the generic
    branch of YgradXYZest_KRR (:731-744 -> dKdMiat_unpermuted :1407-1447 -> dKdXid :1178-1233; exponential and Laplacian
    derivatives by central differences, as the reference does).  Parity unpinned (no executable MLatomF here)."""
    x_tr, x_te, xyz_te, a = _c(x_tr), _c(x_te), _c(xyz_te), _c(alpha).reshape(-1)
    npred, nat, _ = xyz_te.shape
    e, g = np.zeros(npred), np.zeros((npred, nat, 3))
    load().kro_predict_generic(c_int(KERNEL_KINDS[kind.lower()]), c_int(nn), c_int(nat), c_long(x_tr.shape[0]), _d(x_tr),
                               _d(a), c_double(sigma), c_double(prior), c_long(npred), _d(xyz_te), _d(x_te),
                               c_int(1 if calc_grad else 0), _d(e), _d(g))
    return e, (g if calc_grad else None)
