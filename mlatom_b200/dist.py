"""Multi-GPU sharding of the KREG path: one process per GPU, ``torch.distributed`` for plumbing
(NCCL over NVLink on the GPUs; the same code runs on gloo/CPU tensors in the unit tests).

What shards, and how:
  * prediction   -- test geometries are independent: contiguous batches per rank, the model (alpha) is
                    broadcast once, no data-path collective;
  * K assembly   -- row slabs of the row-major lower triangle, balanced by AREA (boundaries ~ sqrt), each
                    slab a contiguous region; the factorising rank receives them with point-to-point
                    send/recv straight into its M x M buffer, then ``alpha`` is broadcast;
  * grid search  -- (sigma) grid points are independent: round-robin over ranks, losses all-gathered.
There is no exchange step anywhere on the path, so nothing here fuses compute with a collective.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n items for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _snap(r: int, align: int, offset: int) -> int:
    """Nearest row boundary that does not split a geometry's gradient rows: offset + k * align for rows past `offset`
    (offset = NtrVal, align = 3 Nat); rows of the value block are left where they are."""
    if align <= 1 or r <= offset:
        return r
    return offset + int(round((r - offset) / align)) * align


def triangle_row_slabs(m: int, world_size: int, align: int = 1, offset: int = 0) -> List[Tuple[int, int]]:
    """Row ranges of an m x m LOWER triangle with (nearly) equal element counts: boundary r_k solves
    r(r+1)/2 = k/world * m(m+1)/2.  `align` / `offset` snap interior boundaries to offset + k * align (3*Nat rows per
    geometry after the NtrVal value rows) so that a geometry's gradient rows are never split."""
    cuts = [0]
    total = m * (m + 1) / 2.0
    for k in range(1, world_size):
        r = int(round((-1.0 + math.sqrt(1.0 + 8.0 * total * k / world_size)) / 2.0))
        r = _snap(r, align, offset)
        cuts.append(min(max(r, cuts[-1]), m))
    cuts.append(m)
    return [(cuts[i], cuts[i + 1]) for i in range(world_size)]


def sub_slabs(slab: Tuple[int, int], pieces: int, align: int = 1, offset: int = 0) -> List[Tuple[int, int]]:
    """Split one rank's row slab [r0, r1) of the lower triangle into `pieces` consecutive row ranges of (nearly) equal
    AREA -- the unit of the pipelined gather: a piece is sent while the next one is still being assembled."""
    r0, r1 = slab
    if r1 <= r0:
        return []
    cuts = [r0]
    a0, a1 = r0 * (r0 + 1) / 2.0, r1 * (r1 + 1) / 2.0
    for k in range(1, pieces):
        t = a0 + (a1 - a0) * k / pieces
        r = _snap(int(round((-1.0 + math.sqrt(1.0 + 8.0 * t)) / 2.0)), align, offset)
        cuts.append(min(max(r, cuts[-1]), r1))
    cuts.append(r1)
    return [(cuts[i], cuts[i + 1]) for i in range(pieces) if cuts[i + 1] > cuts[i]]


def piece_width(b: int, m: int) -> int:
    """Stored row length of a piece that ends at row b: only columns j < b of the lower triangle are defined, so a piece
    travels as a (rows x b) trapezoid instead of full M-wide rows (half the NVLink bytes on average); a multiple of 16
    doubles, so that with `piece_buffer`'s base shift the row segments the assembly kernels write start on 128-byte lines
    (engine.empty_kernel_matrix has the measurement)."""
    up = lambda x: (x + 15) // 16 * 16
    return min(up(m), up(b))


def piece_buffer(rows: int, width: int, offset: int, device) -> torch.Tensor:
    """Contiguous (rows x width) send buffer whose column `offset` (NtrVal of a gradient matrix) is 128-byte aligned."""
    shift = (16 - offset % 16) % 16
    flat = torch.empty(rows * width + shift, dtype=torch.float64, device=device)
    return flat[shift:].view(rows, width)


def gather_row_slabs(k_full: Optional[torch.Tensor], slab: torch.Tensor, slabs: Sequence[Tuple[int, int]],
                     root: int = 0) -> None:
    """Point-to-point gather of whole, full-width row slabs into the root's matrix (kept for callers that hold their
    slab as (r1-r0) x M; the sharded assembly below sends trapezoid pieces instead)."""
    rank, ws = world()
    if ws == 1:
        return
    ops = []
    if rank == root:
        for r, (r0, r1) in enumerate(slabs):
            if r != root and r1 > r0:
                ops.append(dist.P2POp(dist.irecv, k_full[r0:r1], r))
    else:
        r0, r1 = slabs[rank]
        if r1 > r0:
            ops.append(dist.P2POp(dist.isend, slab, root))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def build_kernel_matrix_sharded(build_rows: Callable[[int, int, torch.Tensor, int], None], m: int, device,
                                root: int = 0, align: int = 1, pieces: int = 4,
                                stats: Optional[dict] = None, offset: int = 0) -> Optional[torch.Tensor]:
    """Assemble the lower triangle of K on all ranks, gather on `root`.  `build_rows(r0, r1, out, row_offset)` must fill
    rows [r0, r1) (columns j <= i) into `out`, whose row 0 is matrix row `row_offset` and whose row stride may be
    SHORTER than M (>= the last row's defined length).  Returns K on the root, None elsewhere.

    Every rank cuts its slab into `pieces` equal-area sub-slabs, each in its own contiguous trapezoid buffer; a piece
    is handed to the communication stream (NCCL runs P2P on its own stream, ordered after the work already queued on
    the caller's stream) as soon as its assembly kernels are queued, so the transfer of piece s overlaps the assembly
    of piece s+1.  The root receives into staging buffers and scatters them into its M x M matrix with one strided
    device copy per piece.  `stats` (optional dict) receives the byte count sent to the root."""
    rank, ws = world()
    # boundaries fall on offset + k * align (gradient matrices: offset = NtrVal, align = 3 Nat): a piece that ends at row b
    # then also ends at COLUMN b, so its rows fit a buffer of width b
    slabs = triangle_row_slabs(m, ws, align, offset)
    plans = [sub_slabs(slabs[r], pieces, align, offset) for r in range(ws)]
    if rank == root:
        if device.type == "cuda":
            from . import engine

            k = engine.empty_kernel_matrix(m, offset, device)  # sector-aligned row segments (engine.empty_kernel_matrix)
        else:
            k = torch.empty((m, m), dtype=torch.float64, device=device)
        # Post every receive first: the senders' pieces arrive while the root assembles its own slab.  Piece s of ALL
        # senders forms one batch (one NCCL group): the transfers of a batch run concurrently, so the root's inbound
        # links carry several senders at once -- receives posted one by one are serialised on the communicator and the
        # root then takes in one sender at a time (384 GB/s at 8 GPUs).  Batches complete in order; the pieces of a
        # finished batch are scattered while the next one is still arriving.
        batches = []
        for s in range(max([len(plans[r]) for r in range(ws) if r != root] or [0])):
            stages, ops = [], []
            for r in range(ws):
                if r == root or s >= len(plans[r]):
                    continue
                a, b = plans[r][s]
                st = torch.empty((b - a, piece_width(b, m)), dtype=torch.float64, device=device)
                stages.append((a, b, st))
                ops.append(dist.P2POp(dist.irecv, st, r))
            if ops:
                batches.append((stages, dist.batch_isend_irecv(ops)))
        r0, r1 = slabs[root]
        if r1 > r0:
            build_rows(r0, r1, k, 0)
        nbytes = 0
        for stages, reqs in batches:
            for req in reqs:
                req.wait()
            for (a, b, st) in stages:
                k[a:b, :b].copy_(st[:, :b])
                nbytes += st.numel() * 8
        if stats is not None:
            stats["gather_bytes"] = nbytes
        return k
    reqs, keep = [], []
    for (a, b) in plans[rank]:
        buf = piece_buffer(b - a, piece_width(b, m), offset, device)
        build_rows(a, b, buf, a)
        reqs.append(dist.isend(buf, root))
        keep.append(buf)
    for req in reqs:
        req.wait()
    if stats is not None:
        stats["gather_bytes"] = sum(t.numel() * 8 for t in keep)
    return None


def broadcast_tensor(t: torch.Tensor, root: int = 0) -> torch.Tensor:
    _, ws = world()
    if ws > 1:
        dist.broadcast(t, src=root)
    return t


def predict_sharded(predict: Callable[[torch.Tensor], Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]],
                    xyz: torch.Tensor, gather: bool = False):
    """Each rank predicts its contiguous shard of `xyz` (N,Nat,3).  Returns (lo, hi, E, G) for the shard, or
    the all-gathered full arrays when gather=True."""
    rank, ws = world()
    n = xyz.shape[0]
    lo, hi = shard_range(n, ws, rank)
    e, g = predict(xyz[lo:hi])
    if not gather or ws == 1:
        return lo, hi, e, g
    sizes = [shard_range(n, ws, r) for r in range(ws)]
    width = max(b - a for a, b in sizes)  # all_gather needs equal shapes: pad shards to the widest
    out = []
    for t in (e, g):
        if t is None:
            out.append(None)
            continue
        mine = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        mine[: hi - lo] = t
        parts = [torch.empty_like(mine) for _ in range(ws)]
        dist.all_gather(parts, mine)
        out.append(torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], 0))
    return 0, n, out[0], out[1]


def grid_shard(points: Sequence, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Indices of the grid points this rank evaluates (round robin)."""
    r, ws = world()
    rank = r if rank is None else rank
    world_size = ws if world_size is None else world_size
    return list(range(rank, len(points), world_size))


def allgather_losses(local: dict, n_points: int) -> List[float]:
    """local: {point index: loss} -> list of all losses on every rank.  Which points a rank evaluated travels as an
    explicit mask (a NaN loss -- a NaN alpha -- is a RESULT, not a marker): evaluated NaN losses come back as +inf so
    that optimize_grid's strict '<' stays well defined and a failed point can never win; points nobody evaluated stay
    NaN."""
    _, ws = world()
    vals = torch.zeros(n_points, dtype=torch.float64)
    mask = torch.zeros(n_points, dtype=torch.float64)
    bad = torch.zeros(n_points, dtype=torch.float64)  # evaluated, but +inf or NaN
    for i, v in local.items():
        mask[i] = 1.0
        if v != v or v == float("inf"):
            bad[i] = 1.0
        else:
            vals[i] = v
    if ws > 1:
        backend = dist.get_backend()
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        packed = torch.stack((vals, mask, bad)).to(dev)
        dist.all_reduce(packed)
        vals, mask, bad = packed.cpu()
    out = torch.where(bad > 0, torch.full_like(vals, float("inf")), vals)
    return torch.where(mask > 0, out, torch.full_like(vals, float("nan"))).tolist()
