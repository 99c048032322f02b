"""Permutationally invariant KREG: MLatomF's ``molDescrType=permuted permInvKernel`` on the device.

Reference: the permuted RE descriptor (calcX, MLatomF_src/molDescr.f90:128-160; permuteAtoms :197-220; option parsing
``permInvNuclei=1-2.5-6`` dataset.f90:1952-1977) and the normalised kernel

    k(M_i, M_j) = S_ij / (sqrt(S_ii) sqrt(S_jj)),   S_ij = sum_P exp(-|X_i - X_j^P|^2 / (2 sigma^2))

(kernelFunction, A_KRR_kernel.f90:773-804), training ``calcKernel`` :149-159 + Cholesky, prediction ``Yest_KRR``
:225-245 and the permInvKernel branch of ``YgradXYZest_KRR`` :414-503.  In the reference this path is reached through
``kreg(ml_program='MLatomF')`` with ``additional_mlatomf_args=['permInvKernel', 'permInvNuclei=...',
'molDescrType=permuted']`` (models.py:494-495); `mlatom_b200.models.kreg` routes the same arguments here.

Scope: models trained on VALUES (energies), predicting energies and gradients -- the second-order part of the reference
(gradient-trained permutationally invariant models, :505-725) is not built.

Host logic only in this file (permutation tables); every number comes from the CUDA kernels behind
``include/kreg_b200.h`` (kreg_perm_*).  No CPU fallback.
"""
from __future__ import annotations

import itertools
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib, engine
from ._lib import KregError, call
from .engine import F64, _dev, _on, _ptr, _stream
from .kreg_api import KREG_API, ac2d_array


# ---- permutation tables (host logic) ------------------------------------------------------------------------------
def parse_perm_inv_nuclei(spec: str) -> list:
    """``'4-5.6-7-8'`` -> [[3, 4], [5, 6, 7]]: groups separated by '.', atoms (1-based in the option string, as MLatomF
    reads them: dataset.f90:1952-1973) by '-'; returned 0-based."""
    groups = []
    for part in str(spec).split("."):
        part = part.strip()
        if part:
            groups.append([int(tok) - 1 for tok in part.split("-")])
    return groups


def atom_maps(natoms: int, groups: Iterable[Sequence[int]]) -> np.ndarray:
    """Full 0-based atom maps (Nperm, Nat) of every permutation of the atoms inside each group; Nperm = prod (size)!.
    A map is built as permuteAtoms does (molDescr.f90:197-220): walking the atoms in order, the k-th atom that belongs to
    the permuted set takes the k-th entry of the permuted index list.  Row 0 is the identity (the reference pairs block 1
    of one descriptor with every block of the other)."""
    groups = [sorted(int(a) for a in g) for g in groups]
    flat = [a for g in groups for a in g]
    if len(set(flat)) != len(flat) or any(a < 0 or a >= natoms for a in flat):
        raise ValueError("permutation groups must be disjoint sets of atom indices inside the molecule")
    maps = []
    for combo in itertools.product(*[itertools.permutations(g) for g in groups]):
        perm_list = [a for g in combo for a in g]
        members, m, k = set(perm_list), list(range(natoms)), 0
        for ii in range(natoms):
            if ii in members:
                m[ii] = perm_list[k]
                k += 1
        maps.append(m)
    out = np.asarray(maps, dtype=np.int32).reshape(-1, natoms)
    assert np.array_equal(out[0], np.arange(natoms))
    return out


def descriptor_permutations(pmap: np.ndarray):
    """Entry permutations of the RE descriptor induced by atom maps: pidx[P][d(a,c)] = d(p(a), p(c)) and the tables of
    the inverse maps (mod_inversePermInds, A_KRR_kernel.f90:944-950).  Both int32 (Nperm, D)."""
    pmap = np.asarray(pmap, dtype=np.int64)
    nperm, nat = pmap.shape
    for p in range(nperm):
        if sorted(pmap[p].tolist()) != list(range(nat)):
            raise ValueError("every atom map must be a permutation of 0..Nat-1")
    iu = np.triu_indices(nat, 1)                  # descriptor order: row-major upper triangle (KREG.f90:48)
    index = np.zeros((nat, nat), dtype=np.int64)
    index[iu] = np.arange(len(iu[0]))
    index = index + index.T
    inv = np.empty_like(pmap)
    for p in range(nperm):
        inv[p, pmap[p]] = np.arange(nat)
    pidx = index[pmap[:, iu[0]], pmap[:, iu[1]]]
    pinv = index[inv[:, iu[0]], inv[:, iu[1]]]
    return pidx.astype(np.int32), pinv.astype(np.int32)


# ---- device operations ------------------------------------------------------------------------------------------
class Permutations:
    """Atom maps and their descriptor-index tables, device resident."""

    def __init__(self, pmap: np.ndarray, device):
        self.pmap = np.asarray(pmap, dtype=np.int32)
        self.nperm, self.natoms = self.pmap.shape
        if not np.array_equal(self.pmap[0], np.arange(self.natoms)):
            raise ValueError("the first permutation must be the identity")
        pidx, pinv = descriptor_permutations(self.pmap)
        self.pidx = torch.as_tensor(pidx).to(device).contiguous()
        self.pinv = torch.as_tensor(pinv).to(device).contiguous()


def perm_descriptors(xyz, xeq_: torch.Tensor, perms: Permutations) -> torch.Tensor:
    """Permuted RE descriptor (N, Nperm * D): MLatomF's X for molDescrType=permuted."""
    xyz = _dev(xyz, xeq_.device)
    n, nat, _ = xyz.shape
    out = torch.empty((n, perms.nperm * nat * (nat - 1) // 2), dtype=F64, device=xyz.device)
    with _on(xyz):
        call("kreg_perm_descriptors", nat, n, _ptr(xyz), _ptr(xeq_), perms.nperm, _ptr(perms.pidx), _ptr(out),
             _stream(xyz.device))
    return out


def self_terms(xyz, xeq_: torch.Tensor, perms: Permutations, sigma: float, gradient: bool = False):
    """S_ii (N,) and optionally dS_ii/dxyz (N, Nat, 3)."""
    xyz = _dev(xyz, xeq_.device)
    n, nat, _ = xyz.shape
    s = torch.empty(n, dtype=F64, device=xyz.device)
    ds = torch.empty((n, nat, 3), dtype=F64, device=xyz.device) if gradient else None
    with _on(xyz):
        call("kreg_perm_self_terms", nat, n, _ptr(xyz), _ptr(xeq_), perms.nperm, _ptr(perms.pidx), _ptr(perms.pinv),
             float(sigma), _ptr(s), _ptr(ds), _stream(xyz.device))
    return s, ds


def kernel_block(xyz_a, xyz_b, xeq_: torch.Tensor, perms: Permutations, sigma: float, symmetric: bool = False,
                 lam: float = 0.0) -> torch.Tensor:
    """Normalised permutationally invariant kernel block (na, nb); symmetric=True puts exactly 1 + lam on the diagonal."""
    xyz_a, xyz_b = _dev(xyz_a, xeq_.device), _dev(xyz_b, xeq_.device)
    na, nat, _ = xyz_a.shape
    nb = xyz_b.shape[0]
    out = torch.empty((na, nb), dtype=F64, device=xyz_a.device)
    ws_bytes = _lib.load().kreg_perm_kernel_block_workspace_bytes(nat, na, nb, perms.nperm)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xyz_a.device)
    with _on(xyz_a):
        call("kreg_perm_kernel_block", nat, na, _ptr(xyz_a), nb, _ptr(xyz_b), _ptr(xeq_), perms.nperm, _ptr(perms.pidx),
             float(sigma), 1 if symmetric else 0, float(lam), _ptr(out), out.stride(0), _ptr(ws), ws_bytes,
             _stream(xyz_a.device))
    return out


class PermInvModel:
    """A values-trained permutationally invariant model, device resident: the expanded training set (Ntrain * Nperm
    permuted descriptor rows with alpha_j / sqrt(S_jj)) packed for the ordinary prediction kernels."""

    def __init__(self, xyz_tr, xeq_, alpha, sigma: float, prior: float, perms: Permutations):
        xeq_ = _dev(xeq_)
        xyz_tr = _dev(xyz_tr, xeq_.device)
        alpha = _dev(alpha, xeq_.device).reshape(-1)
        ntr, nat, _ = xyz_tr.shape
        assert alpha.numel() == ntr and perms.natoms == nat
        self.natoms, self.ntrain, self.sigma, self.prior = nat, ntr, float(sigma), float(prior)
        self.perms, self.xeq = perms, xeq_
        s_tr, _ = self_terms(xyz_tr, xeq_, perms, sigma)
        a_exp = torch.empty(ntr * perms.nperm, dtype=F64, device=xeq_.device)
        with _on(xeq_):
            call("kreg_perm_scale_alpha", ntr, perms.nperm, _ptr(alpha), _ptr(s_tr), _ptr(a_exp), _stream(xeq_.device))
        x_exp = perm_descriptors(xyz_tr, xeq_, perms).view(ntr * perms.nperm, -1)
        self.expanded = engine.PackedModel.from_descriptors(x_exp, xeq_, a_exp, sigma, 0.0, nat)

    @property
    def device(self):
        return self.xeq.device

    def predict(self, xyz: torch.Tensor, energy: bool = True, gradient: bool = True):
        """xyz (N, Nat, 3) cuda fp64 -> (E (N,) or None, dE/dxyz (N, Nat, 3) or None)."""
        xyz = _dev(xyz, self.device)
        n = xyz.shape[0]
        e, g = self.expanded.predict(xyz, True, gradient)  # E' is needed for the gradient's quotient-rule term too
        s, ds = self_terms(xyz, self.xeq, self.perms, self.sigma, gradient)
        with _on(xyz):
            call("kreg_perm_finish_prediction", self.natoms, n, _ptr(s), _ptr(ds), self.prior, _ptr(e), _ptr(g),
                 _stream(xyz.device))
        return (e if energy else None), g


class PermInvKREG(KREG_API):
    """`KREG_API` with the permutationally invariant kernel: same methods (train / train_arrays / predict /
    predict_arrays / save_model / load_model), same attributes; ``X`` holds the permuted descriptor (Ntrain, Nperm * D)
    like MLatomF's X array, ``pmap`` the atom maps."""

    def __init__(self, groups=None, pmap: Optional[np.ndarray] = None, device: Optional[torch.device] = None):
        super().__init__(device)
        self.groups, self.pmap, self.Nperm = groups, pmap, None

    def _perms(self, natoms: int) -> Permutations:
        pmap = self.pmap if self.pmap is not None else atom_maps(natoms, self.groups or [])
        self.pmap = np.asarray(pmap, dtype=np.int32)
        return Permutations(self.pmap, self._dev())

    def train_arrays(self, sigma, lambdav, lambdagradxyz, xyz, req, y=None, ygrad=None, prior=None, shutup=True):
        if ygrad is not None:
            raise NotImplementedError("permutationally invariant models are trained on values only here "
                                      "(the reference's second-order part, A_KRR_kernel.f90:505-725, is not built)")
        if y is None:
            raise ValueError("nothing to learn: give property_to_learn")
        self._model, self._graphs = None, {}
        dev = self._dev()
        d_xyz = _dev(xyz, dev)
        self.Ntrain, self.Natoms = int(d_xyz.shape[0]), int(d_xyz.shape[1])
        self.Xsize = self.Natoms * (self.Natoms - 1) // 2
        self.NtrVal, self.NtrGrXYZ, self.Ksize = self.Ntrain, 0, self.Ntrain
        yv = y.detach().cpu().numpy() if torch.is_tensor(y) else np.asarray(y, dtype=np.float64)
        self.Yref = yv
        if prior is None:
            self.prior = 0.0
        elif isinstance(prior, str):
            if prior.casefold() != "mean":
                raise ValueError(f"Unsupported prior: {prior}")
            self.prior = float(np.mean(yv))
        else:
            self.prior = float(prior)
        self.Ytrain = yv - self.prior
        self.sigma, self.lambdav, self.lambdagradxyz = float(sigma), float(lambdav), float(lambdagradxyz)
        perms = self._perms(self.Natoms)
        self.Nperm = perms.nperm
        self.ac2dArray = ac2d_array(self.Natoms)
        self.Req = np.asarray(req.detach().cpu().numpy() if torch.is_tensor(req) else req, dtype=np.float64)
        d_xeq = engine.xeq(_dev(self.Req, dev))
        d_y = _dev(self.Ytrain, dev)
        k = kernel_block(d_xyz, d_xyz, d_xeq, perms, self.sigma, symmetric=True, lam=self.lambdav)
        try:
            d_alpha = engine.solve(k, d_y)
        except KregError as err:
            if err.code <= 0:
                raise
            k = kernel_block(d_xyz, d_xyz, d_xeq, perms, self.sigma, symmetric=True, lam=self.lambdav)
            d_alpha = engine.solve_indefinite(k, d_y)  # the reference's Bunch-Kaufman fallback (mathUtils.f90:24-26)
        del k
        self._model = PermInvModel(d_xyz, d_xeq, d_alpha, self.sigma, self.prior, perms)
        self.XYZ = d_xyz.cpu().numpy()
        self.X = perm_descriptors(d_xyz, d_xeq, perms).cpu().numpy()
        self.Xeq = d_xeq.cpu().numpy()
        self.alpha = d_alpha.cpu().numpy().reshape(1, -1)
        return self

    def _ensure_model(self) -> PermInvModel:
        if self._model is None:
            if self.alpha is None or self.XYZ is None:
                raise RuntimeError("permutationally invariant KREG model not found: train or load a model first")
            dev = self._dev()
            perms = self._perms(int(self.Natoms))
            self._model = PermInvModel(_dev(self.XYZ, dev), engine.xeq(_dev(self.Req, dev)),
                                       _dev(np.asarray(self.alpha).reshape(-1), dev), float(self.sigma),
                                       float(self.prior), perms)
        return self._model

    def predict_arrays(self, xyz, calculate_energy=True, calculate_energy_gradients=True):
        model = self._ensure_model()
        on_device = torch.is_tensor(xyz) and xyz.is_cuda
        e, g = model.predict(_dev(xyz, model.device), calculate_energy, calculate_energy_gradients)
        if on_device:
            return e, g
        return (None if e is None else e.cpu().numpy()), (None if g is None else g.cpu().numpy())

    def hessian_arrays(self, xyz):
        raise NotImplementedError("the reference has no Hessian for the permutationally invariant kernel")

    def clean_up(self):
        self.__init__(self.groups, self.pmap, self.device)

    _KEYS = KREG_API._KEYS + ["Nperm", "pmap"]

    def save_unf(self, filename, atomic_numbers=None):
        raise NotImplementedError("permuted-descriptor models are stored as .npz (no MLatomF-written permuted .unf file "
                                  "exists to check a writer against)")

    def load_unf(self, filename):
        raise NotImplementedError("permuted-descriptor .unf files are not read")
