"""Two ranks, two GPUs, NCCL (VERDICT r01 missing #4): the sharded paths of SURVEY.md 8e on real devices.

Spawned from pytest with torch.multiprocessing; skipped on a box with fewer than two GPUs (the single-GPU round-end
run) -- run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`.  The same host logic runs on
CPU tensors over gloo in tests/test_dist_gloo.py.
"""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    import datetime

    # a rank that raises must not leave its peer waiting for NCCL's default 10-minute watchdog
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=dev, timeout=datetime.timedelta(seconds=120))
    try:
        from mlatom_b200 import dist as kd, engine, synthetic as S

        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        for nat, ntr, grads in ((9, 3000, False), (9, 200, True), (12, 75, True), (26, 20, True)):
            eq = S.equilibrium(nat)
            xyz_h = S.geometries(eq, ntr, 1)
            e_h, g_h = S.pair_potential(xyz_h, eq)
            xyz, xeq = t(xyz_h), engine.xeq(t(eq))
            nv, ng = ntr, (3 * nat * ntr if grads else 0)
            m = nv + ng

            def build_rows(r0, r1, out, row_offset):
                engine.build_kernel_matrix(xyz, xeq, 1.1, 1e-6, 2e-6, use_values=True, use_gradients=grads,
                                           rows=(r0, r1), out=out, row_offset=row_offset)

            stats = {}
            k = kd.build_kernel_matrix_sharded(build_rows, m, dev, root=0, align=3 * nat if grads else 1, pieces=3,
                                               stats=stats, offset=nv if grads else 0)
            y = t(np.concatenate([e_h - e_h.mean()] + ([g_h.reshape(-1)] if grads else [])))
            alpha = torch.empty(m, dtype=torch.float64, device=dev)
            if rank == 0:
                whole = engine.build_kernel_matrix(xyz, xeq, 1.1, 1e-6, 2e-6, use_values=True, use_gradients=grads)
                assert torch.equal(torch.tril(k), torch.tril(whole)), (nat, ntr, grads)  # bit for bit
                assert stats["gather_bytes"] < 8 * m * m * 0.75
                alpha.copy_(engine.solve(k, y))
            else:
                assert k is None
            kd.broadcast_tensor(alpha, 0)
            # every rank now holds the root's alpha: identical checksums and identical predictions from it
            parts = [torch.empty_like(alpha) for _ in range(ws)]
            dist.all_gather(parts, alpha)
            assert all(torch.equal(p, parts[0]) for p in parts)
            model = engine.PackedModel.from_training_set(xyz, xeq, alpha, 1.1, float(e_h.mean()), nv, ng)
            xte = t(S.geometries(eq, 1001, 2))
            lo, hi, e, g = kd.predict_sharded(lambda x: model.predict(x.contiguous()), xte, gather=True)
            e1, g1 = model.predict(xte)
            assert (lo, hi) == (0, 1001)
            assert (e - e1).abs().max().item() <= 1e-12 * e1.abs().max().item() and (g - g1).abs().max().item() <= 1e-11
        # grid search sharded over the ranks == the table one rank computes alone
        from mlatom_b200 import data as D, models as M

        def make_db(nat, n, seed):
            eq = S.equilibrium(nat)
            xyz = S.geometries(eq, n, seed)
            e, g = S.pair_potential(xyz, eq)
            return D.molecular_database.from_arrays(S.atomic_numbers(nat), xyz, energy=e, energy_gradients=g)

        sub, val = make_db(9, 40, 1), make_db(9, 20, 3)
        model = M.kreg(model_file=os.path.join(tmp, f"grid{rank}"), device=dev)
        model.optimize_hyperparameters(hyperparameters=["lambda", "sigma"], optimization_algorithm="grid",
                                       optimization_algorithm_kwargs={"grid_size": 4},
                                       training_kwargs={"property_to_learn": "energy"},
                                       subtraining_molecular_database=sub, validation_molecular_database=val)
        table = torch.tensor([model.grid_losses[k] for k in sorted(model.grid_losses)], dtype=torch.float64, device=dev)
        tabs = [torch.empty_like(table) for _ in range(ws)]
        dist.all_gather(tabs, table)
        assert all(torch.equal(x, tabs[0]) for x in tabs) and torch.isfinite(table).all()
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_ranks_nccl(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))
