"""Write-only HBM ceiling: a fill of 8 GiB (torch's vectorised fill kernel), best of 10, CUDA events.
The gradient-block emit kernel is a pure store stream, so this -- not the copy figure in MEASURED_PEAKS.json -- is what
bounds it from above."""
import json
import torch

x = torch.empty(1 << 30, dtype=torch.float64, device="cuda")
best = 1e9
for _ in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    x.fill_(1.5)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(json.dumps({"fill_bytes": x.numel() * 8, "ms": best, "write_gb_per_s": x.numel() * 8 / best / 1e6}))
