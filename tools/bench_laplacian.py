#!/usr/bin/env python3
"""Laplacian-kernel model (L1 distance, no GEMM form): E only (kernel block + kreg_block_dot) and E + analytic gradients
(kernel block + kreg_laplacian_predict) on one GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlatom_b200 import engine, kernels, synthetic as S

dev = torch.device("cuda")
out = []
for nat, ntr, nte in ((9, 10000, 20000), (21, 10000, 8000)):
    eq = S.equilibrium(nat)
    xeq = engine.xeq(torch.from_numpy(eq).to(dev))
    xtr = torch.from_numpy(S.geometries(eq, ntr, 1)).to(dev)
    xte = torch.from_numpy(S.geometries(eq, nte, 2)).to(dev)
    model = kernels.RadialModel("laplacian", 0, xtr, xeq, torch.randn(ntr, dtype=torch.float64, device=dev), 3.0, 0.0)
    rec = {"natoms": nat, "ntrain": ntr, "ntest": nte}
    for name, grad in (("energy_ms", False), ("energy_gradient_ms", True)):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ms = []
        for it in range(4):
            torch.cuda.synchronize()
            ev[0].record()
            model.predict(xte, True, grad)
            ev[1].record()
            torch.cuda.synchronize()
            if it:
                ms.append(ev[0].elapsed_time(ev[1]))
        rec[name] = float(np.mean(ms))
    rec["geometries_per_s_with_gradients"] = nte / rec["energy_gradient_ms"] * 1e3
    out.append(rec)
print(json.dumps(out))
