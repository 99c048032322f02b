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


def triangle_row_slabs(m: int, world_size: int, align: int = 1) -> List[Tuple[int, int]]:
    """Row ranges of an m x m LOWER triangle with (nearly) equal element counts: boundary r_k solves
    r(r+1)/2 = k/world * m(m+1)/2.  `align` rounds interior boundaries (e.g. to 3*Nat so a geometry's
    gradient rows are never split)."""
    cuts = [0]
    total = m * (m + 1) / 2.0
    for k in range(1, world_size):
        r = int(round((-1.0 + math.sqrt(1.0 + 8.0 * total * k / world_size)) / 2.0))
        if align > 1:
            r = int(round(r / align)) * align
        cuts.append(min(max(r, cuts[-1]), m))
    cuts.append(m)
    return [(cuts[i], cuts[i + 1]) for i in range(world_size)]


def gather_row_slabs(k_full: Optional[torch.Tensor], slab: torch.Tensor, slabs: Sequence[Tuple[int, int]],
                     root: int = 0) -> None:
    """Point-to-point gather of row slabs into the root's matrix.  Rank r owns rows slabs[r] in `slab`
    ((r1-r0) x M, contiguous); the root receives each directly into k_full[r0:r1] (no staging copy)."""
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
                                root: int = 0, align: int = 1) -> Optional[torch.Tensor]:
    """Assemble K on all ranks, gather on `root`.  `build_rows(r0, r1, out, row_offset)` must fill rows
    [r0, r1) of the matrix into `out` whose row 0 is matrix row `row_offset`.  Returns K on the root, None
    elsewhere."""
    rank, ws = world()
    slabs = triangle_row_slabs(m, ws, align)
    r0, r1 = slabs[rank]
    if rank == root:
        k = torch.empty((m, m), dtype=torch.float64, device=device)
        if r1 > r0:
            build_rows(r0, r1, k, 0)
        gather_row_slabs(k, k[r0:r1], slabs, root)
        return k
    slab = torch.empty((max(r1 - r0, 0), m), dtype=torch.float64, device=device)
    if r1 > r0:
        build_rows(r0, r1, slab, r0)
    gather_row_slabs(None, slab, slabs, root)
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
    """local: {point index: loss} -> list of all losses on every rank."""
    _, ws = world()
    vals = torch.full((n_points,), float("nan"), dtype=torch.float64)
    for i, v in local.items():
        vals[i] = v
    if ws == 1:
        return vals.tolist()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    vals = vals.to(dev)
    mask = (~torch.isnan(vals)).to(torch.float64)
    # +inf (a grid point whose matrix was not positive definite) must survive the sum: reduce finite values and
    # the "is infinite" flags separately
    inf = torch.isinf(vals).to(torch.float64)
    filled = torch.nan_to_num(vals, nan=0.0, posinf=0.0, neginf=0.0)
    dist.all_reduce(filled)
    dist.all_reduce(mask)
    dist.all_reduce(inf)
    out = torch.where(inf > 0, torch.full_like(filled, float("inf")), filled)
    return torch.where(mask > 0, out, torch.full_like(filled, float("nan"))).cpu().tolist()
