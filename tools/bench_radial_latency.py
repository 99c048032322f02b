#!/usr/bin/env python3
"""Latency of small-batch predictions with a model on a radial kernel function (Matern, n = 2): the split launch
(training set divided over the chip, kreg_predict_radial with its workspace) beside the unsplit one (same entry point,
no workspace), device-resident input, back-to-back launches."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlatom_b200 import _lib, engine, kernels, synthetic as S
from mlatom_b200._lib import call
from mlatom_b200.engine import _ptr, _stream

dev = torch.device("cuda")
out = []
for nat in (6, 9, 21, 22):
    ntr = 10000
    eq = S.equilibrium(nat)
    xtr = torch.from_numpy(S.geometries(eq, ntr, 1)).to(dev)
    alpha = torch.from_numpy(np.random.default_rng(0).standard_normal(ntr) * 1e-3).to(dev)
    model = kernels.RadialModel("matern", 2, xtr, engine.xeq(torch.from_numpy(eq).to(dev)), alpha, 1.0, 0.0)
    for n in (1, 8, 64, 512):
        x = torch.from_numpy(S.geometries(eq, n, 2)).to(dev)
        e = torch.empty(n, dtype=torch.float64, device=dev)
        g = torch.empty((n, nat, 3), dtype=torch.float64, device=dev)
        nbytes = int(_lib.load().kreg_predict_radial_workspace_bytes(nat, ntr, n))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
        rec = {"natoms": nat, "ntrain": ntr, "n": n, "workspace_bytes": nbytes}
        for name, wb in (("unsplit_us", 0), ("split_us", nbytes)):
            def run():
                call("kreg_predict_radial", kernels.KERNEL_KINDS["matern"], 2, nat, ntr, _ptr(model.packed), _ptr(model.center),
                     _ptr(model.xeq), model.sigma, model.prior, n, _ptr(x), 3, _ptr(e), _ptr(g), _ptr(ws), wb, _stream(dev))
            for _ in range(20): run()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(200): run()
            ev1.record(); torch.cuda.synchronize()
            rec[name] = round(ev0.elapsed_time(ev1) / 200 * 1e3, 1)
            rec[name.replace("_us", "_e0")] = float(e[0])
        out.append(rec)
print(json.dumps(out))
