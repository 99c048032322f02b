#!/usr/bin/env python3
"""BASELINE configs[4]: KREG with forces in training (gradient-augmented kernel) + sigma/lambda grid search.

    python tools/bench_gridsearch.py --ntrain 2000 --grid 9                 # 56k-dim system, 81 grid points
    torchrun --nproc-per-node N ... tools/bench_gridsearch.py ...          # grid points sharded over N GPUs

Uses models.kreg.optimize_hyperparameters(optimization_algorithm='grid') -- the fused device path: K(sigma) assembled
once per sigma, re-factorised per lambda (cuSOLVER potrf, the dominant cost), validation predictions on the GPU.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
from mlatom_b200 import data as D, models, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--natoms", type=int, default=9)
ap.add_argument("--ntrain", type=int, default=2000)
ap.add_argument("--nval", type=int, default=500)
ap.add_argument("--grid", type=int, default=9)
a = ap.parse_args()
ws, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eq = S.equilibrium(a.natoms)
def db(n, seed):
    xyz = S.geometries(eq, n, seed)
    e, g = S.pair_potential(xyz, eq)
    return D.molecular_database.from_arrays(S.atomic_numbers(a.natoms), xyz, energy=e, energy_gradients=g)
sub, val = db(a.ntrain, 1), db(a.nval, 3)
model = models.kreg(model_file=f"/tmp/grid_{rank}.npz")
torch.cuda.synchronize()
t0 = time.perf_counter()
table = model._fused_grid_losses(["lambda", "sigma"], model._grid(["lambda", "sigma"], a.grid),
                                 {"property_to_learn": "energy", "xyz_derivative_property_to_learn": "energy_gradients"},
                                 None, sub, val)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    best = min(table, key=table.get)
    finite = sum(np.isfinite(v) for v in table.values())
    print(json.dumps({"workload": f"grid search {a.grid}x{a.grid} (lambda, sigma), natoms={a.natoms}, ntrain={a.ntrain} with "
                                  f"gradients -> M={a.ntrain * (1 + 3 * a.natoms)}, nval={a.nval}",
                      "n_gpus": ws, "seconds": dt, "grid_points": len(table), "positive_definite_points": int(finite),
                      "best_lambda": best[0], "best_sigma": best[1], "best_loss": table[best]}))
if ws > 1:
    dist.barrier(); dist.destroy_process_group()
