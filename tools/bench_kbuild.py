#!/usr/bin/env python3
"""K-matrix assembly benchmark (BASELINE.json configs[2] and configs[4]).

    python tools/bench_kbuild.py --ntrain 50000                  # cfg3: values K, 50k x 50k (10 GB triangle)
    python tools/bench_kbuild.py --ntrain 2000 --grads           # cfg5: gradient-augmented K, M = 56 000
    torchrun --nproc-per-node N ... tools/bench_kbuild.py ...    # row slabs over N GPUs + gather + 1-GPU potrf

Prints one JSON line: assembly time (max over ranks, CUDA events), GB/s of the stored triangle, gather time,
potrf/potrs time (separate timer), and the residual of the solved system on sampled rows.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mlatom_b200 import dist as kd, engine, synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--natoms", type=int, default=9)
    ap.add_argument("--ntrain", type=int, default=50000)
    ap.add_argument("--grads", action="store_true")
    ap.add_argument("--sigma", type=float, default=1.0)
    ap.add_argument("--lam", type=float, default=1e-6)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--plain", action="store_true", help="K as a dense torch.empty((M, M)) instead of the engine's layout")
    ap.add_argument("--ldk", type=int, default=0, help="row stride of K in doubles (default M): alignment experiments")
    ap.add_argument("--shift", type=int, default=0, help="offset of K's first element from a 256-byte boundary, doubles")
    args = ap.parse_args()

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=dev)
    nat, ntr = args.natoms, args.ntrain
    eq = S.equilibrium(nat)
    xyz_h = S.geometries(eq, ntr, seed=1)
    e_h, g_h = S.pair_potential(xyz_h, eq)
    xyz = torch.from_numpy(xyz_h).to(dev)
    xeq = engine.xeq(torch.from_numpy(eq).to(dev))
    nv, ng = ntr, (3 * nat * ntr if args.grads else 0)
    m = nv + ng
    prior = float(e_h.mean())
    y = torch.from_numpy(np.concatenate([e_h - prior] + ([g_h.reshape(-1)] if args.grads else []))).to(dev)
    align = 3 * nat if args.grads else 1
    slabs = kd.triangle_row_slabs(m, ws, align, nv if args.grads else 0)

    def build_rows(r0, r1, out, row_offset):
        engine.build_kernel_matrix(xyz, xeq, args.sigma, args.lam, args.lam, use_values=True, use_gradients=args.grads,
                                   rows=(r0, r1), out=out, row_offset=row_offset)

    def sync():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    r0, r1 = slabs[rank]
    if rank == 0:
        if args.ldk or args.shift:
            ldk = max(args.ldk, m)
            flat = torch.empty(m * ldk + args.shift, dtype=torch.float64, device=dev)
            k = flat[args.shift:].view(m, ldk)[:, :m]
        elif args.plain or ws > 1:  # gather_row_slabs receives into row ranges of K: dense rows
            k = torch.empty((m, m), dtype=torch.float64, device=dev)
        else:
            k = engine.empty_kernel_matrix(m, nv if args.grads else 0, dev)  # what the engine allocates for itself
        target, off = k, 0
    else:
        k = None
        target, off = torch.empty((r1 - r0, m), dtype=torch.float64, device=dev), r0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    build_ms = []
    for it in range(args.reps + 1):
        sync()
        ev0.record()
        if r1 > r0:
            build_rows(r0, r1, target, off)
        ev1.record()
        sync()
        if it:
            build_ms.append(ev0.elapsed_time(ev1))
    t_build = torch.tensor([float(np.mean(build_ms))], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(t_build, op=dist.ReduceOp.MAX)
    if ws > 1:  # warm the NCCL point-to-point channels up (lazy connection setup is not part of the gather)
        tiny = torch.zeros(8, dtype=torch.float64, device=dev)
        if rank == 0:
            for r in range(1, ws):
                dist.recv(tiny, r)
        else:
            dist.send(tiny, 0)
    sync()
    t0 = time.perf_counter()
    kd.gather_row_slabs(k, target if rank else k[r0:r1], slabs, 0)
    sync()
    gather_ms = 1e3 * (time.perf_counter() - t0)

    out = None
    if rank == 0:
        tri_bytes = 4.0 * m * (m + 1)
        out = {"workload": f"K assembly, natoms={nat}, ntrain={ntr}, gradients={args.grads}, M={m}", "n_gpus": ws,
               "M": m, "bytes_lower_triangle": tri_bytes, "build_ms_max_over_ranks": t_build.item(),
               "build_gb_per_s": tri_bytes / t_build.item() * 1e-6, "gather_ms": gather_ms if ws > 1 else 0.0,
               "slabs": slabs if ws > 1 else None}
        if not args.no_solve:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            alpha = engine.solve(k, y)
            torch.cuda.synchronize()
            out["potrf_potrs_ms"] = 1e3 * (time.perf_counter() - t0)
            if not args.grads:
                idx = torch.from_numpy(np.random.default_rng(0).choice(ntr, size=min(256, ntr), replace=False)).to(dev)
                rows = engine.kernel_block(xyz[idx].contiguous(), xyz, xeq, args.sigma)
                res = rows @ alpha + args.lam * alpha[idx] - y[idx]
                out["residual_rel_sampled_rows"] = (res.norm() / y[idx].norm()).item()
            del k
        print(json.dumps(out))
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
