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
    for m, ws, align in ((50000, 8, 1), (56000, 8, 27), (100, 3, 1), (5, 8, 1)):
        slabs = kd.triangle_row_slabs(m, ws, align)
        assert slabs[0][0] == 0 and slabs[-1][1] == m and len(slabs) == ws
        assert all(slabs[i][1] == slabs[i + 1][0] for i in range(ws - 1))
        if align > 1:
            assert all(a % align == 0 for a, _ in slabs)
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
            out[r0 - row_offset:r1 - row_offset] = ref[r0:r1]

        k = kd.build_kernel_matrix_sharded(build_rows, m, torch.device("cpu"), root=0)
        if rank == 0:
            assert torch.equal(k, ref)
        else:
            assert k is None
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
