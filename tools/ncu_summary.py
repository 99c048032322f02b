#!/usr/bin/env python3
"""Summaries of ncu captures for profiles/.

    python tools/ncu_summary.py launches <launches.csv> [title]     # from: ncu --metrics gpu__time_duration.sum --csv
    python tools/ncu_summary.py full <capture.ncu-rep> [title]      # from: ncu --set full (read here with ncu -i --page raw --csv)
    python tools/ncu_summary.py traffic <capture.ncu-rep> <profiles/traffic.json>   # dram bytes per launch -> bench.py's roofline.traffic
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
]


def launches(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if not r[vi].replace(".", "").replace(",", "").isdigit():
            continue
        ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        n, t = agg.get(r[ki], (0, 0.0))
        agg[r[ki]] = (n + 1, t + ns)
    total = sum(t for _, t in agg.values())
    print(f"# {title}")
    print(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} {n:8d} {t * 1e-6:10.3f} {t / n * 1e-3:10.1f} {100 * t / total:6.2f}%")


def traffic(path, out_json):
    """dram bytes per launch of every kernel in an `ncu --set full` capture -> profiles/traffic.json (what bench.py's
    roofline.traffic reads; never a constant kept in bench.py)."""
    import json
    import os

    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    table = json.load(open(out_json)) if os.path.exists(out_json) else {}
    for r in rows[2:]:
        name = "kreg::" + r[ki].replace("void ", "").split("(")[0].replace(", ", ",")
        rd, wr = float(r[ri]) * scale[units[ri]], float(r[wi]) * scale[units[wi]]
        table[name] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "source": f"{os.path.basename(path)}: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                 f"--clock-control none (summary: profiles/{os.path.basename(path).replace('_prof_', '_').replace('.ncu-rep', '_ncu.txt')})"}
    json.dump(table, open(out_json, "w"), indent=1)
    print(json.dumps(table, indent=1))


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {title}")
    for r in rows[2:]:
        print("-" * 60)
        print(f"{'Kernel Name':112s} {r[ki]}")
        for extra in ("Block Size", "Grid Size"):
            if extra in hdr:
                print(f"{extra:112s} {r[hdr.index(extra)]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:95s} {units[i]:16s} {r[i]}")


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else path
    if mode == "traffic":
        traffic(path, sys.argv[3])
    else:
        (launches if mode == "launches" else full)(path, title)
