"""KREG-style models on MLatomF's other kernel functions (exponential, Matern, Laplacian) over the RE descriptor.

Reference: ``Kfunk`` (MLatomF_src/A_KRR_kernel.f90:806-851, ``distance`` :2040-2074), training ``calcKernel`` :149-159 +
``calcAlpha`` (A_KRR.f90:873-959), prediction ``Yest_KRR`` :225-245 and the generic branch of ``YgradXYZest_KRR``
:731-744 (``dKdMiat_unpermuted`` :1407-1447, ``dKdXid`` :1178-1233).  In the reference these kernels are reached through
``kreg(ml_program='MLatomF')`` with ``additional_mlatomf_args=['kernel=Matern', 'nn=2']`` (models.py:494-495);
`mlatom_b200.models.kreg` routes the same arguments here.

Scope: models trained on VALUES.  Energies for all kernels; analytic gradients for the kernels of the Euclidean
distance (Matern as the reference; exponential analytically where the reference takes central differences); the
Laplacian kernel (L1 distance, no GEMM form) predicts through its materialised kernel block.  Everything runs in the CUDA kernels behind ``include/kreg_b200.h``
(kreg_kernel_block_ex, kreg_solve, kreg_pack_model_radial, kreg_predict_radial, kreg_block_dot, kreg_laplacian_predict);
no CPU fallback.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib, engine
from ._lib import PREDICT_ENERGY, PREDICT_GRADIENT, KregError, call
from .engine import F64, KERNEL_KINDS, _dev, _on, _ptr, _stream
from .kreg_api import KREG_API, ac2d_array

_CANONICAL = {"gaussian": "Gaussian", "exponential": "exponential", "matern": "Matern", "laplacian": "Laplacian"}


class RadialModel:
    """A values-trained model on one of the kernel functions, device resident."""

    def __init__(self, kind: str, nn: int, xyz_tr, xeq_, alpha, sigma: float, prior: float):
        self.kind, self.nn = kind.lower(), int(nn)
        if self.kind not in KERNEL_KINDS:
            raise ValueError(f"unknown kernel {kind!r}")
        xeq_ = _dev(xeq_)
        self.xyz_tr = _dev(xyz_tr, xeq_.device)
        self.alpha = _dev(alpha, xeq_.device).reshape(-1)
        self.ntrain, self.natoms = int(self.xyz_tr.shape[0]), int(self.xyz_tr.shape[1])
        assert self.alpha.numel() == self.ntrain
        self.sigma, self.prior, self.xeq = float(sigma), float(prior), xeq_
        self.packed = self.center = None
        self.x_tr = None  # training descriptors (Laplacian gradients only; built on first use)
        if self.kind != "laplacian":
            nbytes = _lib.load().kreg_radial_packed_model_bytes(self.natoms, self.ntrain)
            self.packed = torch.empty(nbytes, dtype=torch.uint8, device=xeq_.device)
            self.center = torch.empty_like(xeq_)
            with _on(xeq_):
                call("kreg_pack_model_radial", self.natoms, self.ntrain, _ptr(self.xyz_tr), _ptr(xeq_), _ptr(self.alpha),
                     _ptr(self.center), _ptr(self.packed), nbytes, _stream(xeq_.device))

    @property
    def device(self):
        return self.xeq.device

    def workspace(self, n: int) -> Optional[torch.Tensor]:
        """Scratch of a batch of n geometries (the small-batch split launch; the tiled path above 24 atoms), for a caller
        that must own it -- a CUDA graph keeps the raw pointer (engine.GraphPredictor)."""
        if self.kind == "laplacian":
            return None
        with _on(self.xeq):
            nbytes = int(_lib.load().kreg_predict_radial_workspace_bytes(self.natoms, self.ntrain, int(n)))
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)

    def predict(self, xyz: torch.Tensor, energy: bool = True, gradient: bool = True, out_e: Optional[torch.Tensor] = None,
                out_g: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None):
        """xyz (N, Nat, 3) cuda fp64 -> (E (N,) or None, dE/dxyz (N, Nat, 3) or None).  out_e / out_g / ws: caller-owned
        result and scratch buffers (the Euclidean-distance kernels; the same protocol as engine.PackedModel.predict)."""
        xyz = _dev(xyz, self.device)
        n = int(xyz.shape[0])
        if self.kind == "laplacian":
            # L1 distance: no GEMM form.  The kernel block goes through HBM in slabs of at most 256 MB; energies are its
            # product with alpha, gradients the analytic sign form (the reference differentiates this kernel numerically,
            # A_KRR_kernel.f90:1222-1230)
            e = torch.empty(n, dtype=F64, device=xyz.device) if energy or not gradient else None
            g = torch.empty((n, self.natoms, 3), dtype=F64, device=xyz.device) if gradient else None
            if gradient and self.x_tr is None:
                self.x_tr = engine.descriptors(self.xyz_tr, self.xeq)
            step = max(1, (256 << 20) // (8 * max(self.ntrain, 1)))
            for lo in range(0, n, step):
                x = xyz[lo:lo + step]
                kt = engine.kernel_block_ex("laplacian", x, self.xyz_tr, self.xeq, self.sigma)
                with _on(xyz):
                    if gradient:
                        flags = PREDICT_GRADIENT | (PREDICT_ENERGY if energy else 0)
                        call("kreg_laplacian_predict", self.natoms, self.ntrain, _ptr(self.x_tr), _ptr(self.alpha),
                             _ptr(self.xeq), self.sigma, self.prior, kt.shape[0], _ptr(x), _ptr(kt), kt.stride(0), flags,
                             _ptr(e[lo:lo + step]) if energy else None, _ptr(g[lo:lo + step]), _stream(xyz.device))
                    else:
                        call("kreg_block_dot", kt.shape[0], kt.shape[1], _ptr(kt), kt.stride(0), _ptr(self.alpha), self.prior,
                             _ptr(e[lo:lo + step]), _stream(xyz.device))
            return (e if energy else None), g
        e = (out_e if out_e is not None else torch.empty(n, dtype=F64, device=xyz.device)) if energy else None
        g = (out_g if out_g is not None else torch.empty((n, self.natoms, 3), dtype=F64, device=xyz.device)) if gradient else None
        if n == 0:
            return e, g
        flags = (PREDICT_ENERGY if energy else 0) | (PREDICT_GRADIENT if gradient else 0)
        with _on(xyz):
            if ws is not None:
                ws_bytes = ws.numel()
            else:
                ws_bytes = int(_lib.load().kreg_predict_radial_workspace_bytes(self.natoms, self.ntrain, n))
                ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=xyz.device)
            call("kreg_predict_radial", KERNEL_KINDS[self.kind], self.nn, self.natoms, self.ntrain, _ptr(self.packed),
                 _ptr(self.center), _ptr(self.xeq), self.sigma, self.prior, n, _ptr(xyz), flags, _ptr(e), _ptr(g), _ptr(ws),
                 ws_bytes, _stream(xyz.device))
        return e, g


class RadialKREG(KREG_API):
    """`KREG_API` on another kernel function: same methods and attributes plus ``kernel`` and ``nn``."""

    def __init__(self, kernel: str = "Matern", nn: int = 2, device: Optional[torch.device] = None):
        super().__init__(device)
        if str(kernel).lower() not in KERNEL_KINDS:
            raise ValueError(f"kernel={kernel!r}: one of Gaussian, exponential, Matern, Laplacian")
        self.kernel, self.nn = _CANONICAL[str(kernel).lower()], int(nn)

    def clean_up(self):
        self.__init__(self.kernel, self.nn, self.device)

    def train_arrays(self, sigma, lambdav, lambdagradxyz, xyz, req, y=None, ygrad=None, prior=None, shutup=True):
        if ygrad is not None:
            raise NotImplementedError("models on the non-Gaussian kernels are trained on values only here")
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
        self.ac2dArray = ac2d_array(self.Natoms)
        self.Req = np.asarray(req.detach().cpu().numpy() if torch.is_tensor(req) else req, dtype=np.float64)
        d_xeq = engine.xeq(_dev(self.Req, dev))
        d_y = _dev(self.Ytrain, dev)

        def assemble():
            return engine.kernel_block_ex(self.kernel, d_xyz, d_xyz, d_xeq, self.sigma, nn=self.nn, symmetric=True,
                                          lam=self.lambdav)

        try:
            d_alpha = engine.solve(assemble(), d_y)
        except KregError as err:
            if err.code <= 0:
                raise
            d_alpha = engine.solve_indefinite(assemble(), d_y)  # the reference's Bunch-Kaufman fallback (mathUtils.f90:24-26)
        self._model = RadialModel(self.kernel, self.nn, d_xyz, d_xeq, d_alpha, self.sigma, self.prior)
        self.XYZ = d_xyz.cpu().numpy()
        self.X = engine.descriptors(d_xyz, d_xeq).cpu().numpy()
        self.Xeq = d_xeq.cpu().numpy()
        self.alpha = d_alpha.cpu().numpy().reshape(1, -1)
        return self

    def _ensure_model(self) -> RadialModel:
        if self._model is None:
            if self.alpha is None or self.XYZ is None:
                raise RuntimeError("model not found: train or load a model first")
            dev = self._dev()
            self._model = RadialModel(str(self.kernel), int(self.nn), _dev(self.XYZ, dev), engine.xeq(_dev(self.Req, dev)),
                                      _dev(np.asarray(self.alpha).reshape(-1), dev), float(self.sigma), float(self.prior))
        return self._model

    def predict_arrays(self, xyz, calculate_energy=True, calculate_energy_gradients=True):
        model = self._ensure_model()
        on_device = torch.is_tensor(xyz) and xyz.is_cuda
        if not on_device and model.kind != "laplacian" and model.device.type == "cuda":
            n = int(np.shape(xyz)[0])
            if 0 < n <= self.graph_batch_limit:
                # MD / geometry optimisation: the same small batch every step -> replay a captured CUDA graph, as the
                # Gaussian model does (KREG_API.predict_arrays)
                key = (n, bool(calculate_energy), bool(calculate_energy_gradients))
                gp = self._graphs.get(key)
                if gp is None or gp.model is not model:
                    gp = self._graphs[key] = engine.GraphPredictor(model, n, calculate_energy, calculate_energy_gradients)
                e, g = gp(np.asarray(xyz.detach().cpu().numpy() if torch.is_tensor(xyz) else xyz, dtype=np.float64))
                return (None if e is None else e.copy()), (None if g is None else g.copy())
        e, g = model.predict(_dev(xyz, model.device), calculate_energy, calculate_energy_gradients)
        if on_device:
            return e, g
        return (None if e is None else e.cpu().numpy()), (None if g is None else g.cpu().numpy())

    def hessian_arrays(self, xyz):
        raise NotImplementedError("the reference's Hessian exists for the Gaussian KREG model only (KREG.f90:393-420)")

    _KEYS = KREG_API._KEYS + ["kernel", "nn"]

    def load_model(self, filename):
        super().load_model(filename)
        self.kernel, self.nn = _CANONICAL[str(self.kernel).lower()], int(self.nn)

    def save_unf(self, filename, atomic_numbers=None):
        raise NotImplementedError("models on the non-Gaussian kernels are stored as .npz")

    def load_unf(self, filename):
        raise NotImplementedError("models on the non-Gaussian kernels are stored as .npz")
