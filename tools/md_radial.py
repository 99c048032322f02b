#!/usr/bin/env python3
"""Batched NVE (md.BatchedNVE, CUDA-graphed fused velocity Verlet) driven by a Matern-kernel model: energy conservation and
graph replay == eager stepping."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlatom_b200 import kernels, md, synthetic as S

nat, ntr, b = 9, 400, 64
eq = S.equilibrium(nat)
xtr = S.geometries(eq, ntr, 1, spread=0.08)
e, _ = S.pair_potential(xtr, eq)
api = kernels.RadialKREG("Matern", 2)
api.train_arrays(2.0, 1e-8, 0.0, xtr, eq, y=e, prior="mean")
pm = api._ensure_model()
x0 = torch.from_numpy(S.geometries(eq, b, 5, spread=0.02)).cuda()
v0 = torch.from_numpy(np.random.default_rng(6).standard_normal((b, nat, 3)) * 0.02).cuda()
masses = torch.from_numpy(np.where(S.atomic_numbers(nat) == 1, 1.0, 12.0)).cuda()
runs, out_rec = {}, {}
for use_graph in (True, False):
    sim = md.BatchedNVE(pm, x0, v0, masses, dt=0.01, use_graph=use_graph)
    e0 = (sim.epot + sim.ekin).clone()
    out = sim.run(300, record_every=50)
    etot = out["energies"].sum(1)
    out_rec[f"drift_graph_{use_graph}"] = (etot - e0).abs().max().item()
    out_rec["ekin_max"] = sim.ekin.abs().max().item()
    runs[use_graph] = (out["xyz"].clone(), etot.clone())
out_rec["graph_equals_eager"] = bool(torch.equal(runs[True][0], runs[False][0]) and torch.equal(runs[True][1], runs[False][1]))
print(json.dumps(out_rec))
