"""Throughput of E + dE/dxyz prediction with a Matern-kernel model (kreg_predict_radial, three-kernel path) on one GPU."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mlatom_b200 import engine, kernels  # noqa: E402
from mlatom_b200 import synthetic as S  # noqa: E402

out = []
for nat, ntr, nte in ((9, 10000, 151552), (21, 10000, 75776)):
    dev = torch.device("cuda")
    eq = S.equilibrium(nat)
    xeq = engine.xeq(torch.from_numpy(eq).to(dev))
    xtr = torch.from_numpy(S.geometries(eq, ntr, 1)).to(dev)
    xte = torch.from_numpy(S.geometries(eq, nte, 2)).to(dev)
    alpha = torch.randn(ntr, dtype=torch.float64, device=dev)
    d = nat * (nat - 1) // 2
    for kind, nn in (("matern", 2), ("gaussian", 0)):
        model = kernels.RadialModel(kind, nn, xtr, xeq, alpha, 2.0, 0.0)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ms = []
        for it in range(4):
            torch.cuda.synchronize()
            ev[0].record()
            model.predict(xte)
            ev[1].record()
            torch.cuda.synchronize()
            if it:
                ms.append(ev[0].elapsed_time(ev[1]))
        t = float(np.mean(ms))
        out.append({"natoms": nat, "kernel": kind, "nn": nn, "ntrain": ntr, "ntest": nte, "ms": t,
                    "geometries_per_s": nte / t * 1e3, "algorithmic_tflops": float(nte) * ntr * (5 * d + 4) / t * 1e-9})
print(json.dumps(out))
