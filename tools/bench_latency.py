#!/usr/bin/env python3
"""Latency of small-batch predictions (the MD / geometry-optimisation step: md.py:169,198 calls predict on ONE
geometry per step).  Times kreg_predict on device-resident input for several batch sizes, with the stream kept busy
(back-to-back launches) and per call including a device->host read of E."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlatom_b200 import engine, synthetic as S
nat, ntr = 9, 10000
eq = S.equilibrium(nat)
dev = torch.device("cuda")
xtr = torch.from_numpy(S.geometries(eq, ntr, 1)).to(dev)
alpha = torch.from_numpy(np.random.default_rng(0).standard_normal(ntr) * 1e-3).to(dev)
model = engine.PackedModel.from_training_set(xtr, engine.xeq(torch.from_numpy(eq).to(dev)), alpha, 1.0, 0.0, ntr, 0)
out = {}
for n in (1, 8, 64, 512, 4096, 18944):
    x = torch.from_numpy(S.geometries(eq, n, 2)).to(dev)
    for _ in range(20): model.predict(x)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(200): model.predict(x)
    ev1.record(); torch.cuda.synchronize()
    back_to_back_us = ev0.elapsed_time(ev1) / 200 * 1e3
    t0 = time.perf_counter()
    for _ in range(200):
        e, g = model.predict(x); e[0].item()
    sync_us = (time.perf_counter() - t0) / 200 * 1e6
    rec = {"device_us": round(back_to_back_us, 1), "call_plus_readback_us": round(sync_us, 1)}
    if n <= 512:
        gp = engine.GraphPredictor(model, n)
        xh = x.cpu().numpy()
        for _ in range(20): gp(xh)
        t0 = time.perf_counter()
        for _ in range(500): gp(xh)
        rec["graph_host_to_host_us"] = round((time.perf_counter() - t0) / 500 * 1e6, 1)
        if n <= 64:
            for zc in (False, True):
                gp = engine.GraphPredictor(model, n, zero_copy=zc)
                for _ in range(20): gp(xh)
                t0 = time.perf_counter()
                for _ in range(500): gp(xh)
                rec["graph_zero_copy_us" if zc else "graph_copies_us"] = round((time.perf_counter() - t0) / 500 * 1e6, 1)
    out[n] = rec
print(json.dumps({"natoms": nat, "ntrain": ntr, "latency_by_batch": out}))
