#!/usr/bin/env python3
"""Host-to-host latency of one prediction step (numpy in, numpy out) on a Matern model through RadialKREG.predict_arrays:
the CUDA-graph replay that small batches take, beside the eager device path + copies."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlatom_b200 import kernels, synthetic as S

nat, ntr = 9, 10000
eq = S.equilibrium(nat)
xtr = S.geometries(eq, ntr, 1)
api = kernels.RadialKREG("Matern", 2)
api.train_arrays(1.0, 1e-6, 0.0, xtr[:200], eq, y=np.zeros(200))           # shapes only: the model below is what is timed
api._model = kernels.RadialModel("matern", 2, torch.from_numpy(xtr).cuda(), api._model.xeq,
                                 torch.randn(ntr, dtype=torch.float64, device="cuda") * 1e-3, 1.0, 0.0)
out = {"natoms": nat, "ntrain": ntr}
for n in (1, 64):
    x = S.geometries(eq, n, 2)
    for name, limit in (("graph_us", 64), ("eager_us", 0)):
        api.graph_batch_limit = limit
        for _ in range(30): api.predict_arrays(x)
        t0 = time.perf_counter()
        for _ in range(500): api.predict_arrays(x)
        out[f"batch_{n}_{name}"] = round((time.perf_counter() - t0) / 500 * 1e6, 1)
print(json.dumps(out))
