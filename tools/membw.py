#!/usr/bin/env python3
"""Write-only / read-only / copy HBM bandwidth on this box (context for the K-assembly roofline: the
MEASURED_PEAKS.json figure is a copy, i.e. half reads and half writes)."""
import torch
n = 1 << 30  # 8 GiB of fp64
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def t(fn, reps=5):
    best = 1e9
    for _ in range(reps + 2):
        torch.cuda.synchronize(); ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    return best
w = t(lambda: a.fill_(1.5)); print(f"write-only fill   : {8*n/w*1e-6:8.1f} GB/s")
w = t(lambda: a.zero_());    print(f"write-only memset : {8*n/w*1e-6:8.1f} GB/s")
w = t(lambda: a.sum());      print(f"read-only sum     : {8*n/w*1e-6:8.1f} GB/s")
w = t(lambda: b.copy_(a));   print(f"copy (r+w bytes)  : {16*n/w*1e-6:8.1f} GB/s")
