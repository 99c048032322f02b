#!/usr/bin/env python3
"""Per-kernel device time of one prediction call at several batch sizes (run under ncu, or alone to get the totals)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlatom_b200 import engine, synthetic as S
nat, ntr = 9, 10000
eq = S.equilibrium(nat)
dev = torch.device("cuda")
xtr = torch.from_numpy(S.geometries(eq, ntr, 1)).to(dev)
alpha = torch.from_numpy(np.random.default_rng(0).standard_normal(ntr) * 1e-3).to(dev)
model = engine.PackedModel.from_training_set(xtr, engine.xeq(torch.from_numpy(eq).to(dev)), alpha, 1.0, 0.0, ntr, 0)
for n in (1, 8, 64, 512):
    x = torch.from_numpy(S.geometries(eq, n, 2)).to(dev)
    for _ in range(6):
        model.predict(x)
    torch.cuda.synchronize()
