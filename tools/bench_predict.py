#!/usr/bin/env python3
"""Prediction-kernel benchmark for any molecule size / model kind (BASELINE configs[3] = 21 atoms).

    python tools/bench_predict.py --natoms 21 --ntrain 10000 --ntest 200000 [--grads] [--energy-only]
Prints ms per pass, geometries/s and algorithmic TFLOP/s (SURVEY.md 8d flop counts)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlatom_b200 import engine, synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--natoms", type=int, default=21)
ap.add_argument("--ntrain", type=int, default=10000)
ap.add_argument("--ntest", type=int, default=200000)
ap.add_argument("--grads", action="store_true", help="gradient-trained model (v_j terms)")
ap.add_argument("--energy-only", action="store_true")
ap.add_argument("--sigma", type=float, default=None)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--host", action="store_true", help="stream the test set from pinned host memory (predict_host)")
a = ap.parse_args()
nat, D = a.natoms, a.natoms * (a.natoms - 1) // 2
sigma = a.sigma or (1.0 if nat <= 12 else 2.0)
eq = S.equilibrium(nat)
dev = torch.device("cuda")
xtr = torch.from_numpy(S.geometries(eq, a.ntrain, 1)).to(dev)
xte_h = torch.from_numpy(S.geometries(eq, a.ntest, 2))
xte = xte_h.pin_memory() if a.host else xte_h.to(dev)
out_e = torch.empty(a.ntest, dtype=torch.float64).pin_memory() if a.host else None
out_g = torch.empty((a.ntest, nat, 3), dtype=torch.float64).pin_memory() if a.host and not a.energy_only else None
hp = None
nv, ng = a.ntrain, (3 * nat * a.ntrain if a.grads else 0)
alpha = torch.from_numpy(np.random.default_rng(0).standard_normal(nv + ng) * 1e-3).to(dev)
model = engine.PackedModel.from_training_set(xtr, engine.xeq(torch.from_numpy(eq).to(dev)), alpha, sigma, 0.0, nv, ng)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ms = []
for it in range(a.reps + 2):
    torch.cuda.synchronize(); ev[0].record()
    if a.host:
        if hp is None:
            hp = engine.HostPredictor(model)
        hp.run(xte, True, not a.energy_only, out_e, out_g)
    else:
        model.predict(xte, True, not a.energy_only)
    ev[1].record(); torch.cuda.synchronize()
    if it >= 2: ms.append(ev[0].elapsed_time(ev[1]))
t = float(np.mean(ms))
pairs = float(a.ntest) * a.ntrain
per_pair = (3 * D + 4) if a.energy_only else ((9 * D + 8) if a.grads else (5 * D + 4))
flop = pairs * per_pair + (0 if a.energy_only else a.ntest * 30.0 * D)
print(json.dumps({"natoms": nat, "D": D, "ntrain": a.ntrain, "ntest": a.ntest, "gradient_model": a.grads,
                  "energy_only": a.energy_only, "from_host": a.host, "ms": t, "geom_per_s": a.ntest / t * 1e3,
                  "ps_per_pair": t * 1e9 / pairs, "algorithmic_tflops": flop / t * 1e-9}))
