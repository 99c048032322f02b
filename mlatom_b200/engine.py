"""Device-side KREG operations on PyTorch fp64 CUDA tensors.

PyTorch is plumbing here (device memory, streams, pinned host buffers); every number is
produced by the hand-written sm_100a kernels behind the C ABI in ``include/kreg_b200.h``.
There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import FILL_FULL, FILL_LOWER, PREDICT_ENERGY, PREDICT_GRADIENT, KregError, call

F64 = torch.float64


try:  # the raw handle of the current stream without building a torch.cuda.Stream object (7 us per call saved)
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # pragma: no cover - older / newer torch without the private getter
    _raw_stream = None


def _stream(device=None) -> int:
    """Raw handle of the current stream of ``device`` (default: the current device)."""
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    if _raw_stream is not None:
        return _raw_stream(idx)
    return torch.cuda.current_stream(idx).cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(t: torch.Tensor):
    """Device guard for an ABI call: every kernel launch, cuSOLVER handle and SM-count query in the library binds to
    the CURRENT device, so a call on tensors of another GPU (``kreg(device='cuda:1')`` while device 0 is current) must
    switch to the tensors' device first.  No-op (no context-manager cost on the MD step) when it already is current;
    pinned host tensors (zero-copy graph path) run on the current device."""
    if not t.is_cuda or t.device.index is None or t.device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(t.device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dev(t, device=None) -> torch.Tensor:
    """fp64, contiguous, on a CUDA device (host arrays are copied once)."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64))
    if t.dtype != F64:
        t = t.to(F64)
    if not t.is_cuda:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def device_check() -> None:
    call("kreg_device_check")


def measure_fp64_peak(kind: str = "dmma", reps: int = 5, ms_target: float = 20.0) -> float:
    """TFLOP/s of the FP64 pipe through DFMA or DMMA (roofline denominator of the prediction kernel)."""
    out = ctypes.c_double(0.0)
    call("kreg_measure_fp64_peak", 0 if kind == "dfma" else 1, reps, float(ms_target), ctypes.byref(out))
    return out.value


def xeq(req) -> torch.Tensor:
    """Equilibrium pair distances (KREG.f90:223-239). req (Nat,3) -> (D,)."""
    req = _dev(req)
    nat = req.shape[0]
    out = torch.empty(nat * (nat - 1) // 2, dtype=F64, device=req.device)
    with _on(req):
        call("kreg_xeq", nat, _ptr(req), _ptr(out), _stream(req.device))
    return out


def descriptors(xyz, xeq_: torch.Tensor) -> torch.Tensor:
    """RE descriptors (KREG.f90:241-271). xyz (N,Nat,3) -> (N,D); bit-identical to the reference."""
    xyz = _dev(xyz, xeq_.device)
    n, nat, _ = xyz.shape
    out = torch.empty((n, nat * (nat - 1) // 2), dtype=F64, device=xyz.device)
    with _on(xyz):
        call("kreg_descriptors", nat, n, _ptr(xyz), _ptr(xeq_), _ptr(out), _stream(xyz.device))
    return out


def build_workspace_bytes(nat: int, ntr: int, ntrgr: int, rows: Optional[Tuple[int, int]] = None) -> int:
    """Scratch bytes of one assembly call; with ``rows`` only the pair-term block of the geometries those rows touch."""
    lib = _lib.load()
    if rows is None:
        return int(lib.kreg_build_workspace_bytes(nat, ntr, ntrgr))
    return int(lib.kreg_build_workspace_bytes_rows(nat, ntr, ntrgr, int(rows[0]), int(rows[1])))


def empty_kernel_matrix(m: int, ntrval: int, device) -> torch.Tensor:
    """Uninitialised (M, M) float64 matrix laid out for the assembly kernels' store stream: row stride a multiple of
    16 doubles and the base shifted so that column `ntrval` -- where every row segment of the gradient blocks starts --
    sits on a 128-byte line.  The blocks are written as full 32-byte sectors then; with an odd M or NtrVal every segment
    starts and ends in a partial sector and the assembly runs up to 25 % slower (3.87 -> 5.01 TB/s at
    14 atoms).  The result is a strided view: ``k.stride(0)`` is the ldk/lda the C ABI takes."""
    ld = (m + 15) // 16 * 16
    shift = (16 - ntrval % 16) % 16
    flat = torch.empty(m * ld + shift, dtype=F64, device=device)
    return flat[shift:].view(m, ld)[:, :m]


def build_kernel_matrix(
    xyz: torch.Tensor,
    xeq_: torch.Tensor,
    sigma: float,
    lambdav: float,
    lambdagradxyz: float,
    use_values: bool = True,
    use_gradients: bool = False,
    fill: int = FILL_LOWER,
    rows: Optional[Tuple[int, int]] = None,
    out: Optional[torch.Tensor] = None,
    row_offset: int = 0,
    workspace: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """K + diag(lambda) assembled on the device (KREG.f90:293-346 + :616-622).

    Returns the (M, M) row-major matrix; with ``fill=FILL_LOWER`` only j <= i is defined.
    ``rows=(r0, r1)`` restricts the assembly to a row slab (multi-GPU sharding); ``out`` is then either the
    whole matrix or, with ``row_offset=r0``, a slab buffer whose first row is matrix row r0.
    ``workspace`` (uint8, >= build_workspace_bytes) lets a caller that assembles repeatedly (grid search) reuse one
    scratch allocation.
    """
    xyz = _dev(xyz, xeq_.device)
    ntr, nat, _ = xyz.shape
    ntrval = ntr if use_values else 0
    ntrgr = 3 * nat * ntr if use_gradients else 0
    m = ntrval + ntrgr
    if out is None:
        out = empty_kernel_matrix(m, ntrval if use_gradients else 0, xyz.device)
    r0, r1 = (0, m) if rows is None else rows
    # a slab buffer may be narrower than M under FILL_LOWER: only columns j < (row range end) are ever written
    assert out.is_cuda and out.dtype == F64 and out.stride(1) == 1
    assert out.shape[1] >= (m if fill != FILL_LOWER else min(m, r1))
    assert out.shape[0] >= r1 - row_offset and row_offset <= r0
    ws_bytes = build_workspace_bytes(nat, ntr, ntrgr, None if rows is None else (r0, r1))
    if workspace is not None and workspace.numel() >= ws_bytes and workspace.device == xyz.device:
        ws = workspace
    else:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device)
    with _on(xyz):
        call(
            "kreg_build_kernel_matrix", nat, ntr, ntrval, ntrgr, _ptr(xyz), _ptr(xeq_), float(sigma), float(lambdav),
            float(lambdagradxyz), r0, r1, fill, _ptr(out) - row_offset * out.stride(0) * 8, out.stride(0), _ptr(ws),
            ws.numel(), _stream(xyz.device),
        )
    return out


def kernel_block(xyz_a: torch.Tensor, xyz_b: torch.Tensor, xeq_: torch.Tensor, sigma: float) -> torch.Tensor:
    """Rectangular value-value kernel block k(X_a[i], X_b[j]) (A_KRR_kernel.f90:190-223)."""
    xyz_a, xyz_b = _dev(xyz_a, xeq_.device), _dev(xyz_b, xeq_.device)
    na, nat, _ = xyz_a.shape
    nb = xyz_b.shape[0]
    out = torch.empty((na, nb), dtype=F64, device=xyz_a.device)
    lib = _lib.load()
    ws_bytes = lib.kreg_build_workspace_bytes(nat, na, 0) + lib.kreg_build_workspace_bytes(nat, nb, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz_a.device)
    with _on(xyz_a):
        call("kreg_kernel_block", nat, na, _ptr(xyz_a), nb, _ptr(xyz_b), _ptr(xeq_), float(sigma), _ptr(out),
             out.stride(0), _ptr(ws), ws_bytes, _stream(xyz_a.device))
    return out


KERNEL_KINDS = {"gaussian": 0, "exponential": 1, "matern": 2, "laplacian": 3}


def kernel_block_ex(kind: str, xyz_a: torch.Tensor, xyz_b: torch.Tensor, xeq_: torch.Tensor, sigma: float, nn: int = 2,
                    symmetric: bool = False, lam: float = 0.0) -> torch.Tensor:
    """Value-value kernel block for MLatomF's other kernel functions on the RE descriptor ('Gaussian', 'exponential',
    'Matern' of order nn, 'Laplacian'; Kfunk, A_KRR_kernel.f90:806-851).  symmetric=True (same geometry set on both
    sides) puts exactly 1 + lam on the diagonal: the matrix ``solve`` factorises for a values-trained model."""
    xyz_a, xyz_b = _dev(xyz_a, xeq_.device), _dev(xyz_b, xeq_.device)
    na, nat, _ = xyz_a.shape
    nb = xyz_b.shape[0]
    out = torch.empty((na, nb), dtype=F64, device=xyz_a.device)
    ws_bytes = _lib.load().kreg_kernel_block_ex_workspace_bytes(nat, na, nb)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xyz_a.device)
    with _on(xyz_a):
        call("kreg_kernel_block_ex", KERNEL_KINDS[kind.lower()], int(nn), nat, na, _ptr(xyz_a), nb, _ptr(xyz_b), _ptr(xeq_),
             float(sigma), 1 if symmetric else 0, float(lam), _ptr(out), out.stride(0), _ptr(ws), ws_bytes,
             _stream(xyz_a.device))
    return out


def solve(k: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """alpha = (K)^-1 y by in-place Cholesky (cuSOLVER potrf/potrs; mathUtils.f90:8-104).

    ``k`` (row-major, lower triangle valid, lambda already on the diagonal) is DESTROYED;
    returns alpha.  Raises KregError(code > 0) if K is not positive definite."""
    m = k.shape[0]
    alpha = y.detach().clone().to(F64).contiguous()
    with _on(k):
        dbytes, hbytes = ctypes.c_size_t(0), ctypes.c_size_t(0)
        call("kreg_solve_workspace_bytes", m, ctypes.byref(dbytes), ctypes.byref(hbytes))
        dws = torch.empty(max(dbytes.value, 16), dtype=torch.uint8, device=k.device)
        hws = _pinned_scratch(max(hbytes.value, 16))
        call("kreg_solve", m, _ptr(k), k.stride(0), _ptr(alpha), _ptr(dws), dbytes.value, hws.data_ptr(), hbytes.value,
             _stream(k.device))
    return alpha


SOLVE_AUTO, SOLVE_BUNCH_KAUFMAN, SOLVE_LU = 0, 1, 2


def solve_indefinite(k: torch.Tensor, y: torch.Tensor, method: int = SOLVE_AUTO) -> torch.Tensor:
    """alpha = K^-1 y for a symmetric matrix that is NOT positive definite: Bunch-Kaufman (cuSOLVER sytrf/sytrs) or
    pivoted LU, behind the C ABI (kreg_solve_indefinite; the reference: solveSysLinEqsBK, mathUtils.f90:106-186).
    ``k`` must be freshly assembled (lower triangle valid) -- not the matrix a failed ``solve`` left behind -- and is
    destroyed.  Raises KregError(code > 0) for an exactly singular pivot."""
    m = k.shape[0]
    alpha = y.detach().clone().to(F64).contiguous()
    with _on(k):
        dbytes, hbytes = ctypes.c_size_t(0), ctypes.c_size_t(0)
        call("kreg_solve_indefinite_workspace_bytes", m, int(method), ctypes.byref(dbytes), ctypes.byref(hbytes))
        dws = torch.empty(max(dbytes.value, 16), dtype=torch.uint8, device=k.device)
        hws = _pinned_scratch(max(hbytes.value, 16))
        call("kreg_solve_indefinite", m, _ptr(k), k.stride(0), _ptr(alpha), int(method), _ptr(dws), dbytes.value,
             hws.data_ptr(), hbytes.value, _stream(k.device))
    return alpha


_PINNED: dict = {}


def _pinned_scratch(nbytes: int) -> torch.Tensor:
    """One small pinned host buffer per size class, reused across solves (pinning costs ~100 us per allocation)."""
    size = 1 << max(8, (int(nbytes) - 1).bit_length())
    buf = _PINNED.get(size)
    if buf is None:
        buf = _PINNED[size] = torch.empty(size, dtype=torch.uint8).pin_memory()
    return buf


def hessian(x_train: torch.Tensor, alpha_val: Optional[torch.Tensor], xeq_: torch.Tensor, sigma: float,
            xyz: torch.Tensor) -> torch.Tensor:
    """Reference-defined Hessian of the predicted energy (KREG.f90:180-221, 393-420): (N, 3Nat, 3Nat)."""
    xyz = _dev(xyz, xeq_.device)
    n, nat, _ = xyz.shape
    out = torch.zeros((n, 3 * nat, 3 * nat), dtype=F64, device=xyz.device)
    ntr = 0 if alpha_val is None else int(alpha_val.numel())
    with _on(xyz):
        call("kreg_hessian", nat, ntr, _ptr(x_train), _ptr(alpha_val), _ptr(xeq_), float(sigma), n, _ptr(xyz), _ptr(out),
             _stream(xyz.device))
    return out


def precompute_v(xyz_tr: torch.Tensor, xeq_: torch.Tensor, alpha_grad: torch.Tensor) -> torch.Tensor:
    """v_j = J_j alpha_grad,j. alpha_grad (Ntr, Nat, 3) -> (Ntr, D)."""
    ntr, nat, _ = xyz_tr.shape
    out = torch.empty((ntr, nat * (nat - 1) // 2), dtype=F64, device=xyz_tr.device)
    with _on(xyz_tr):
        call("kreg_precompute_v", nat, ntr, _ptr(xyz_tr), _ptr(xeq_), _ptr(alpha_grad.contiguous()), _ptr(out),
             _stream(xyz_tr.device))
    return out


@dataclass
class PackedModel:
    """A trained KREG model resident in HBM in the fragment order the prediction kernel streams."""

    natoms: int
    ntrain: int
    sigma: float
    prior: float
    has_gradient_terms: bool
    xeq: torch.Tensor      # (D,)
    center: torch.Tensor   # (D,)
    packed: torch.Tensor   # uint8 blob

    @property
    def device(self):
        return self.packed.device

    @classmethod
    def from_training_set(cls, xyz_tr, xeq_, alpha, sigma: float, prior: float, ntrval: int, ntrgr: int, device=None):
        """alpha: (NtrVal + NtrGrXYZ,) in the reference order [values ; (point, atom, xyz)]."""
        xeq_ = _dev(xeq_, device)
        xyz_tr = _dev(xyz_tr, xeq_.device)
        alpha = _dev(alpha, xeq_.device).reshape(-1)
        ntr, nat, _ = xyz_tr.shape
        assert alpha.numel() == ntrval + ntrgr
        alpha_v = alpha[:ntrval].contiguous() if ntrval else None
        v = None
        if ntrgr:
            v = precompute_v(xyz_tr, xeq_, alpha[ntrval:].reshape(ntr, nat, 3))
        nbytes = _lib.load().kreg_packed_model_bytes(nat, ntr, 1 if v is not None else 0)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=xeq_.device)
        center = torch.empty_like(xeq_)
        with _on(xeq_):
            call("kreg_pack_model", nat, ntr, _ptr(xyz_tr), _ptr(xeq_), _ptr(alpha_v), _ptr(v), float(sigma),
                 _ptr(center), _ptr(packed), nbytes, _stream(xeq_.device))
        return cls(nat, ntr, float(sigma), float(prior), v is not None, xeq_, center, packed)

    @classmethod
    def from_descriptors(cls, x_tr, xeq_, alpha_v, sigma: float, prior: float, natoms: int, device=None):
        """Values-only model given by its training DESCRIPTORS (Ntrain, D) instead of geometries -- what MLatomF's
        .unf files carry (MLmodel.f90:880-886)."""
        xeq_ = _dev(xeq_, device)
        x_tr = _dev(x_tr, xeq_.device)
        alpha_v = _dev(alpha_v, xeq_.device).reshape(-1)
        ntr = int(x_tr.shape[0])
        assert x_tr.shape[1] == natoms * (natoms - 1) // 2 and alpha_v.numel() == ntr
        nbytes = _lib.load().kreg_packed_model_bytes(natoms, ntr, 0)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=xeq_.device)
        center = torch.empty_like(xeq_)
        with _on(xeq_):
            call("kreg_pack_model_x", natoms, ntr, _ptr(x_tr), _ptr(alpha_v), None, float(sigma), _ptr(center),
                 _ptr(packed), nbytes, _stream(xeq_.device))
        return cls(natoms, ntr, float(sigma), float(prior), False, xeq_, center, packed)

    def predict(self, xyz: torch.Tensor, energy: bool = True, gradient: bool = True, out_e=None, out_g=None, ws=None):
        """xyz (N, Nat, 3) cuda fp64 -> (E (N,) or None, dE/dxyz (N, Nat, 3) or None).  Pinned host tensors are accepted
        too (unified addressing: the kernels then read / write host memory directly -- only sensible for a few
        geometries, where it saves the copy operations; see GraphPredictor).

        ``ws``: scratch tensor owned by the CALLER (``workspace(n)``).  Anything that records this call in a CUDA graph
        must pass its own: the graph keeps the raw pointer, so the scratch has to live as long as the graph does."""
        assert (xyz.is_cuda or xyz.is_pinned()) and xyz.dtype == F64 and xyz.is_contiguous()
        n = xyz.shape[0]
        assert xyz.shape[1] == self.natoms
        odev = xyz.device if xyz.is_cuda else self.device
        e = (out_e if out_e is not None else torch.empty(n, dtype=F64, device=odev)) if energy else None
        g = (out_g if out_g is not None else torch.empty((n, self.natoms, 3), dtype=F64, device=odev)) if gradient else None
        flags = (PREDICT_ENERGY if energy else 0) | (PREDICT_GRADIENT if gradient else 0)
        hv = 1 if self.has_gradient_terms else 0
        ws_bytes = 0
        if ws is not None:
            ws_bytes = ws.numel()
        elif n > 0 and (self._always_ws or n <= self._small_batch_limit()):
            need = self.workspace_bytes(n)
            if need:
                ws = self._shared_ws(need, odev)
                ws_bytes = ws.numel()
        with _on(self.packed):
            call("kreg_predict", self.natoms, self.ntrain, _ptr(self.packed), _ptr(self.center), _ptr(self.xeq),
                 self.sigma, self.prior, hv, n, _ptr(xyz), flags, _ptr(e), _ptr(g), _ptr(ws), ws_bytes,
                 _stream(self.packed.device))
        return e, g

    _ws: Optional[torch.Tensor] = None       # scratch shared by eager calls, grown on demand
    _ws_retired: Optional[list] = None       # outgrown scratch buffers: kept alive (see _shared_ws)
    _wave_rows: Optional[int] = None

    @property
    def _always_ws(self) -> bool:
        """Molecules above the register-resident kernels' range (> 24 atoms) run the tiled three-kernel path, which
        needs scratch for every batch size."""
        return self.natoms > _lib.FUSED_MAX_NATOMS

    def workspace_bytes(self, n: int) -> int:
        """Scratch bytes kreg_predict can use for a batch of n geometries (0: none needed)."""
        with _on(self.packed):
            return int(_lib.load().kreg_predict_workspace_bytes(self.natoms, self.ntrain,
                                                                1 if self.has_gradient_terms else 0, int(n)))

    def workspace(self, n: int) -> Optional[torch.Tensor]:
        """A scratch tensor the caller owns (CUDA-graph owners: one per captured batch shape)."""
        need = self.workspace_bytes(n)
        return torch.empty(need, dtype=torch.uint8, device=self.device) if need else None

    def _shared_ws(self, need: int, device) -> torch.Tensor:
        # A larger batch replaces the shared scratch, but the outgrown buffer is never handed back to the allocator
        # while the model lives: a CUDA graph captured around an eager predict() (torch.cuda.graph over user code)
        # holds its raw pointer.  The total is bounded (the small-batch scratch tops out at ~2 x SMs x tile bytes).
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            if self._ws is not None:
                if self._ws_retired is None:
                    self._ws_retired = []
                self._ws_retired.append(self._ws)
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def _small_batch_limit(self) -> int:
        return self.wave_rows() // 2

    def wave_rows(self) -> int:
        """Test geometries in one full wave of the prediction kernel on this device (SM count x rows per CTA)."""
        if self._wave_rows is None:
            rows = _lib.load().kreg_predict_tile_rows(self.natoms, 1 if self.has_gradient_terms else 0)
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            self._wave_rows = int(rows) * int(sms)
        return self._wave_rows

    _host_predictor = None

    def predict_host(self, xyz_host: np.ndarray, energy: bool = True, gradient: bool = True, chunk: Optional[int] = None):
        """Host-buffer entry point: numpy (N, Nat, 3) in, numpy out.  Streams chunks through pinned
        staging buffers with H2D copy / kernel / D2H copy overlapped on three streams."""
        hp = self._host_predictor
        if hp is None or (chunk is not None and hp.chunk != int(chunk)):
            hp = self._host_predictor = HostPredictor(self, chunk)  # streams, events, staging buffers: built once
        return hp.run(xyz_host, energy, gradient)


class HostPredictor:
    """Double-buffered host <-> device pipeline around PackedModel.predict."""

    def __init__(self, model: PackedModel, chunk: Optional[int] = None, nbuf: int = 3):
        if chunk is None:
            chunk = 4 * model.wave_rows()  # whole waves per chunk (up to 2 CTAs per SM): no partially filled last wave
        self.model, self.chunk, self.nbuf = model, int(chunk), nbuf
        dev, nat = model.device, model.natoms
        self.s_in, self.s_k, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.d_xyz = [torch.empty((chunk, nat, 3), dtype=F64, device=dev) for _ in range(nbuf)]
        self.d_e = [torch.empty(chunk, dtype=F64, device=dev) for _ in range(nbuf)]
        self.d_g = [torch.empty((chunk, nat, 3), dtype=F64, device=dev) for _ in range(nbuf)]
        self.ev_in = [torch.cuda.Event() for _ in range(nbuf)]
        self.ev_k = [torch.cuda.Event() for _ in range(nbuf)]
        self.ev_out = [torch.cuda.Event() for _ in range(nbuf)]

    def run(self, xyz_host, energy=True, gradient=True, out_e=None, out_g=None):
        m = self.model
        xyz_t = xyz_host if torch.is_tensor(xyz_host) else torch.from_numpy(np.ascontiguousarray(xyz_host, dtype=np.float64))
        n = xyz_t.shape[0]
        if out_e is None and energy:
            out_e = torch.empty(n, dtype=F64).pin_memory()
        if out_g is None and gradient:
            out_g = torch.empty((n, m.natoms, 3), dtype=F64).pin_memory()
        with torch.cuda.device(m.device):
            start = torch.cuda.current_stream()
            for s in (self.s_in, self.s_k, self.s_out):
                s.wait_stream(start)
            for ci, lo in enumerate(range(0, n, self.chunk)):
                hi = min(n, lo + self.chunk)
                b = ci % self.nbuf
                with torch.cuda.stream(self.s_in):
                    if ci >= self.nbuf:
                        self.s_in.wait_event(self.ev_k[b])  # kernel that read this input buffer is done
                    self.d_xyz[b][: hi - lo].copy_(xyz_t[lo:hi], non_blocking=True)
                    self.ev_in[b].record()
                with torch.cuda.stream(self.s_k):
                    self.s_k.wait_event(self.ev_in[b])
                    if ci >= self.nbuf:
                        self.s_k.wait_event(self.ev_out[b])  # previous results left this output buffer
                    m.predict(self.d_xyz[b][: hi - lo], energy, gradient,
                              out_e=self.d_e[b][: hi - lo] if energy else None,
                              out_g=self.d_g[b][: hi - lo] if gradient else None)
                    self.ev_k[b].record()
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(self.ev_k[b])
                    if energy:
                        out_e[lo:hi].copy_(self.d_e[b][: hi - lo], non_blocking=True)
                    if gradient:
                        out_g[lo:hi].copy_(self.d_g[b][: hi - lo], non_blocking=True)
                    self.ev_out[b].record()
            start.wait_stream(self.s_out)
            start.wait_stream(self.s_k)
            start.wait_stream(self.s_in)
        return out_e, out_g


class GraphPredictor:
    """Fixed-shape prediction captured ONCE in a CUDA graph (pinned-host xyz -> device -> kernels -> pinned-host E, G).

    For the loops that call ``predict`` on the same number of geometries every step -- molecular dynamics
    (mlatom/md.py:169,198: one geometry per step; md_parallel.py:168,194: one batch per step) and geometry
    optimisation -- replaying the graph removes the per-call launch and Python overhead of the four operations.
    """

    ZERO_COPY_MAX = 0  # kernels reading / writing the pinned host buffers directly (two graph nodes instead of five) measured
    #                    41-60 us against 42-46 us with copy nodes for one geometry, and slower from 8 geometries on: off by default

    def __init__(self, model: PackedModel, n: int, energy: bool = True, gradient: bool = True, zero_copy=None):
        self.model, self.n, self.energy, self.gradient = model, int(n), energy, gradient
        self.zero_copy = (n <= self.ZERO_COPY_MAX) if zero_copy is None else bool(zero_copy)
        dev, nat = model.device, model.natoms
        self.x_host = torch.zeros((n, nat, 3), dtype=F64).pin_memory()
        # E and dE/dxyz share one buffer on each side: one device->host copy node instead of two
        self.out_host = torch.zeros(n * (1 + 3 * nat), dtype=F64).pin_memory()
        self.e_host, self.g_host = self.out_host[:n], self.out_host[n:].view(n, nat, 3)
        self._x_np, self._e_np, self._g_np = self.x_host.numpy(), self.e_host.numpy(), self.g_host.numpy()
        self.x = torch.zeros((n, nat, 3), dtype=F64, device=dev)
        self.out = torch.zeros(n * (1 + 3 * nat), dtype=F64, device=dev)
        self.e, self.g = self.out[:n], self.out[n:].view(n, nat, 3)
        self.ws = model.workspace(n)  # owned by this graph: the captured kernels keep its raw pointer
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside capture: allocates the small-batch workspace, sets attributes
                    self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            # capture on a stream of THIS device: torch's default capture stream is one process-wide object created
            # on whichever device captured first, and a graph captured through it on another device comes out empty
            with torch.cuda.graph(self.graph, stream=side):
                self._body()

    def _current_stream(self) -> int:
        """Raw handle of the stream the replay was just enqueued on (the capture device's current stream)."""
        idx = self.model.device.index
        if _raw_stream is not None:
            return _raw_stream(torch.cuda.current_device() if idx is None else idx)
        return torch.cuda.current_stream(self.model.device).cuda_stream

    def _body(self):
        if self.zero_copy:  # two graph nodes instead of five: no copy operations around the kernels
            self.model.predict(self.x_host, self.energy, self.gradient, out_e=self.e_host, out_g=self.g_host, ws=self.ws)
            return
        self.x.copy_(self.x_host, non_blocking=True)
        self.model.predict(self.x, self.energy, self.gradient, out_e=self.e, out_g=self.g, ws=self.ws)
        if self.energy and self.gradient:
            self.out_host.copy_(self.out, non_blocking=True)
        elif self.energy:
            self.e_host.copy_(self.e, non_blocking=True)
        elif self.gradient:
            self.g_host.copy_(self.g, non_blocking=True)

    def __call__(self, xyz: np.ndarray):
        """xyz (n, Nat, 3) numpy -> (E (n,), dE/dxyz (n, Nat, 3)) numpy views of the pinned result buffers."""
        self._x_np[...] = xyz
        with _on(self.x):  # a graph replays on the CURRENT device's stream: make that the device it was captured on
            self.graph.replay()
            call("kreg_stream_synchronize", self._current_stream())
        return (self._e_np if self.energy else None), (self._g_np if self.gradient else None)
