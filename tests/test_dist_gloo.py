"""CPU, world_size 2 over gloo: the multi-GPU host logic (shard arithmetic, slab gather, alpha broadcast,
sharded prediction, grid sharding) with CPU tensors standing in for device buffers."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from mlatom_b200 import dist as kd  # noqa: E402


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1000003):
        for ws in (1, 2, 3, 8):
            spans = [kd.shard_range(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_triangle_slabs_balanced_and_aligned():
    for m, ws, align, offset in ((50000, 8, 1, 0), (56000, 8, 27, 2000), (100, 3, 1, 0), (5, 8, 1, 0), (54000, 4, 27, 0)):
        slabs = kd.triangle_row_slabs(m, ws, align, offset)
        assert slabs[0][0] == 0 and slabs[-1][1] == m and len(slabs) == ws
        assert all(slabs[i][1] == slabs[i + 1][0] for i in range(ws - 1))
        if align > 1:  # boundaries inside the gradient part never split a geometry's 3 Nat rows
            assert all(a <= offset or (a - offset) % align == 0 for a, _ in slabs)
        if m >= 1000:
            areas = [b * (b + 1) / 2 - a * (a + 1) / 2 for a, b in slabs]
            assert max(areas) / (sum(areas) / ws) < 1.02


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        m = 37
        ref = torch.arange(m * m, dtype=torch.float64).reshape(m, m)

        def build_rows(r0, r1, out, row_offset):
            # like the assembly kernels under FILL_LOWER: only j <= i is written; the buffer may be narrower than m
            w = min(out.shape[1], m)
            view = out[r0 - row_offset:r1 - row_offset, :w]
            view.fill_(float("nan"))
            keep = torch.tril(torch.ones(r1 - r0, w, dtype=torch.bool), diagonal=r0)
            view[keep] = ref[r0:r1, :w][keep]

        stats = {}
        k = kd.build_kernel_matrix_sharded(build_rows, m, torch.device("cpu"), root=0, pieces=3, stats=stats)
        if rank == 0:
            assert torch.equal(torch.tril(k), torch.tril(ref))
            assert 0 < stats["gather_bytes"] < 8 * m * m  # trapezoids, not full-width rows
        else:
            assert k is None and stats["gather_bytes"] > 0
        # pieces: equal-area split of a slab, aligned
        for slab, pieces, align, offset in (((1000, 5000), 4, 1, 0), ((200 + 27 * 10, 200 + 27 * 200), 5, 27, 200),
                                            ((3, 4), 4, 1, 0)):
            sub = kd.sub_slabs(slab, pieces, align, offset)
            assert sub[0][0] == slab[0] and sub[-1][1] == slab[1]
            assert all(sub[i][1] == sub[i + 1][0] for i in range(len(sub) - 1))
            assert all((a - offset) % align == 0 for a, _ in sub)
            if slab[1] - slab[0] > 100:
                areas = [b * (b + 1) / 2 - a * (a + 1) / 2 for a, b in sub]
                assert max(areas) / (sum(areas) / len(areas)) < 1.1
        # a NaN loss is a result, not "not evaluated": it must come back as +inf, unevaluated points as NaN
        got = kd.allgather_losses({0: float("nan")} if rank == 0 else {1: 2.0}, 3)
        assert got[0] == float("inf") and got[1] == 2.0 and got[2] != got[2]
        # alpha broadcast
        alpha = torch.arange(5, dtype=torch.float64) if rank == 0 else torch.zeros(5, dtype=torch.float64)
        kd.broadcast_tensor(alpha, 0)
        assert torch.equal(alpha, torch.arange(5, dtype=torch.float64))
        # sharded prediction with gather: a fake model E = sum(xyz), G = 2*xyz
        xyz = torch.arange(11 * 3 * 3, dtype=torch.float64).reshape(11, 3, 3)
        lo, hi, e, g = kd.predict_sharded(lambda x: (x.sum((1, 2)), 2 * x), xyz, gather=True)
        assert (lo, hi) == (0, 11) and torch.equal(e, xyz.sum((1, 2))) and torch.equal(g, 2 * xyz)
        lo, hi, e, g = kd.predict_sharded(lambda x: (x.sum((1, 2)), None), xyz)
        assert (lo, hi) == kd.shard_range(11, ws, rank) and g is None and e.shape[0] == hi - lo
        # grid sharding
        pts = list(range(9))
        mine = kd.grid_shard(pts)
        losses = kd.allgather_losses({i: float(i * i) for i in mine}, len(pts))
        assert losses == [float(i * i) for i in pts]
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))
