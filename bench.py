#!/usr/bin/env python3
"""bench.py -- KREG E+F prediction throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] -- KREG energies + gradients, 9-atom
ethanol-shaped synthetic geometries, Ntrain = 10 000 (values-trained, alpha solved on the GPU
before the timed region), Ntest = 1 000 000 geometries PER GPU (weak scaling: test batches are
independent).  One "step" = one pass of the fused E + dE/dxyz prediction kernel over
the rank's 1M-geometry batch.  Inputs + outputs of a step are 440 MB per GPU (> 126 MB L2), so no
explicit L2 flush is needed between timed iterations.

JSON keys beyond the base contract:
  roofline     dominant kernel (predict_kernel) against the FP64 tensor pipe (DMMA) peak measured
               live by kreg_measure_fp64_peak() -- MEASURED_PEAKS.json has no FP64 entry
  kbuild       K-matrix assembly GB/s (values block, Ntrain=10k) and its two ceilings; potrf timed apart
  cpu_baseline the oracle's port of MLatomF's prediction path on the host cores (bounded sample)
  e2e          same metric through PackedModel.predict_host: pinned host xyz in, host E/G out, copies
               inside the timed region
  step_latency one geometry / 64 geometries per call (the MD step): device and CUDA-graph host-to-host microseconds
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

NAT, NTRAIN, NTEST, SIGMA, LAMBDA = 9, 10000, 1000000, 1.0, 1e-6
D = NAT * (NAT - 1) // 2
METRIC = "KREG E+F predictions/s (9 atoms, Ntrain=10k, Ntest=1M per GPU)"
UNIT = "geometries/s"
WORKLOAD = "cfg2: KREG energies+forces prediction, 9 atoms, Ntrain=10000, Ntest=1000000 per GPU"
# algorithmic work: E + dE/dxyz, values-trained
FLOP_PER_STEP = float(NTEST) * NTRAIN * (5 * D + 4) + float(NTEST) * 30 * D
SCALING = "weak"
KERNEL = "kreg::predict_kernel<9,2,false,false,3,3,true,8,false>"


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the ncu capture recorded in
    profiles/traffic.json (written by tools/ncu_summary.py next to the summary it was read from); None if the kernel
    has no entry -- never a constant kept in this file."""
    try:
        norm = lambda k: k.replace("false", "0").replace("true", "1").replace(" ", "")  # ncu prints bool arguments as 0 / 1
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            rec = {norm(k): v for k, v in json.load(f).items()}.get(norm(kernel))
        return (rec["dram_bytes"], rec["source"]) if rec else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


TRAFFIC, TRAFFIC_SOURCE = measured_traffic(KERNEL)


def set_workload(name, world):
    """--workload cfg4: BASELINE.json configs[3], 10M 21-atom geometries in total, split over the ranks (strong scaling).
    The default (cfg2) is the configuration the metric is quoted on and is what the driver runs."""
    global NAT, NTRAIN, NTEST, SIGMA, D, METRIC, WORKLOAD, FLOP_PER_STEP, SCALING, KERNEL, TRAFFIC, TRAFFIC_SOURCE
    if name == "cfg2":
        return
    NAT, NTRAIN, NTEST, SIGMA = 21, 10000, 10_000_000 // world, 2.0
    D = NAT * (NAT - 1) // 2
    METRIC = "KREG E+F predictions/s (21 atoms, Ntrain=10k, 10M geometries in total)"
    WORKLOAD = (f"cfg4: KREG batched E+F screening of 10M 21-atom aspirin-shaped geometries, Ntrain=10000 (assumed, "
                f"SURVEY.md 8d), {NTEST} per GPU")
    FLOP_PER_STEP = float(NTEST) * NTRAIN * (5 * D + 4) + float(NTEST) * 30 * D
    SCALING = "strong"
    KERNEL = "kreg::predict_kernel<21,1,false,true,1,3,true,8,false>"
    TRAFFIC, TRAFFIC_SOURCE = measured_traffic(KERNEL)


def shared_config():
    """The workload description, IDENTICAL in both arms (`--impl ours` / `--impl reference`): what is arm-specific goes to
    the top-level `notes` key, so that the driver's config comparison sees the same object."""
    return {"workload": WORKLOAD, "natoms": NAT, "descriptor_size": D, "ntrain": NTRAIN, "ntest_per_gpu": NTEST,
            "sigma": SIGMA, "lambda": LAMBDA, "prior": "mean",
            "model": "values-trained KREG model (RE descriptor, Gaussian kernel), energies + gradients predicted"}


def training_set():
    from mlatom_b200 import synthetic as S

    eq = S.equilibrium(NAT)
    xyz = S.geometries(eq, NTRAIN, seed=1)
    e, _ = S.pair_potential(xyz, eq)
    return eq, xyz, e


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(eq, xyz_tr, alpha, prior, seconds=12.0, threads=None):
    """The oracle's port of MLatomF's prediction path (A_KRR.f90:985-1029, A_KRR_kernel.f90:225-245,
    :327-352) on the host cores, on a bounded sample of the same workload."""
    from mlatom_b200 import synthetic as S
    from oracle import kreg_oracle as O

    threads = threads or os.cpu_count() or 1
    O.set_threads(threads)
    xeq = O.xeq(eq)
    x_tr = O.descriptors(xyz_tr, xeq)

    last = {}

    def run(n):
        xte = S.geometries(eq, n, seed=2)  # the first n geometries of rank 0's test batch (same generator stream)
        t0 = time.perf_counter()
        x_te = O.descriptors(xte, xeq)  # descriptor build is part of the path
        last["e"], last["g"] = O.predict_mlatomf(xyz_tr, x_tr, alpha, SIGMA, prior, NTRAIN, 0, xte, x_te, True, True)
        return time.perf_counter() - t0

    n0 = 4 * threads
    t = run(n0)
    n = int(max(n0, min(NTEST, n0 * seconds / max(t, 1e-6))))
    n = max(threads, n // threads * threads)
    t = run(n)
    return {"value": n / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} of {NTEST} test geometries, full Ntrain={NTRAIN}, {t:.1f} s wall",
            "results": (last["e"], last["g"])}, n, t


def step_latency(engine, model, dev):
    """What an MD / geometry-optimisation driver pays per step (md.py:169,198 predicts ONE geometry per step;
    md_parallel.py one batch per step): device time of kreg_predict in latency mode and host-to-host time of the
    CUDA-graphed predictor, E + dE/dxyz, against this run's Ntrain = 10k model."""
    import torch

    from mlatom_b200 import synthetic as S

    out = {}
    eq = S.equilibrium(NAT)
    for n in (1, 64):
        x = torch.from_numpy(S.geometries(eq, n, seed=3)).to(dev)
        for _ in range(20):
            model.predict(x)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(200):
            model.predict(x)
        ev[1].record()
        torch.cuda.synchronize()
        gp = engine.GraphPredictor(model, n)
        xh = x.cpu().numpy()
        for _ in range(20):
            gp(xh)
        t0 = time.perf_counter()
        for _ in range(500):
            gp(xh)
        out[f"batch_{n}"] = {"device_us": ev[0].elapsed_time(ev[1]) / 200 * 1e3,
                             "graph_host_to_host_us": (time.perf_counter() - t0) / 500 * 1e6}
    return out


def kbuild_configs(engine, dev, hbm_peak, fp64_peak):
    """K assembly at BASELINE configs[2] (values, 50k) and configs[4] (gradient-augmented, 2000 x 9 atoms -> 56k)."""
    import torch

    from mlatom_b200 import synthetic as S

    out = {}
    for name, nat, ntr, grads in (("cfg3_values_50k", NAT, 50000, False), ("cfg5_gradients_56k", NAT, 2000, True),
                                  ("aspirin_gradients_55k", 21, 857, True), ("nat12_gradients_55k", 12, 1500, True)):
        d = nat * (nat - 1) // 2
        eq = S.equilibrium(nat)
        xeq = engine.xeq(torch.from_numpy(eq).to(dev))
        sigma = SIGMA if nat <= 12 else 2.0
        xyz = torch.from_numpy(S.geometries(eq, ntr, seed=1)).to(dev)
        m = ntr + (3 * nat * ntr if grads else 0)
        k = engine.empty_kernel_matrix(m, ntr if grads else 0, dev)  # the layout the engine itself allocates
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ms = []
        for it in range(4):
            torch.cuda.synchronize()
            ev[0].record()
            engine.build_kernel_matrix(xyz, xeq, sigma, LAMBDA, LAMBDA, use_gradients=grads, out=k)
            ev[1].record()
            torch.cuda.synchronize()
            if it:
                ms.append(ev[0].elapsed_time(ev[1]))
        t = float(np.mean(ms))
        nbytes = 4.0 * m * (m + 1)
        rec = {"natoms": nat, "ntrain": ntr, "M": m, "ms": t, "bytes_lower_triangle": nbytes, "gb_per_s": nbytes / t * 1e-6,
               "frac_of_hbm_copy_peak": nbytes / t * 1e-6 / hbm_peak}
        if not grads:
            fp64_ms = ntr * (ntr + 1) / 2 * (3 * d + 3) / (fp64_peak * 1e12) * 1e3
            rec.update({"bound": "fp64 issue (SURVEY.md App. D)", "fp64_bound_ms": fp64_ms, "frac_of_fp64_bound": fp64_ms / t})
        else:
            rec.update({"bound": "hbm write"})
        out[name] = rec
        del k
        torch.cuda.empty_cache()
    return out


def md_steps_per_s(engine, model, dev):
    """Batched NVE (mlatom_b200.md.BatchedNVE: fused velocity-Verlet kernels + kreg_predict, one CUDA graph per step) on
    this run's Ntrain = 10k model: steps/s for 1, 64 and 4096 trajectories (md.py:158-205 does one trajectory per call)."""
    import torch

    from mlatom_b200 import md, synthetic as S

    eq = S.equilibrium(NAT)
    masses = torch.from_numpy(np.where(S.atomic_numbers(NAT) == 1, 1.008, 12.011)).to(dev)
    out = {}
    for b in (1, 64, 4096):
        x0 = torch.from_numpy(S.geometries(eq, b, 5, spread=0.02)).to(dev)
        v0 = torch.zeros_like(x0)
        sim = md.BatchedNVE(model, x0, v0, masses, dt=0.1, use_graph=True, units="mlatom")
        sim.run(50)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 500 if b < 4096 else 200
        sim.run(n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[f"trajectories_{b}"] = {"steps_per_s": n / dt, "trajectory_steps_per_s": n * b / dt, "us_per_step": dt / n * 1e6}
    return out


def kbuild_sharded(engine, dev, world, rank, barrier, max_over_ranks):
    """BASELINE configs[2]: the 50k x 50k values K (10 GB lower triangle) assembled as row slabs over the ranks and gathered
    onto the factorising GPU (rank 0) with the pipelined trapezoid gather of mlatom_b200.dist; potrf timed apart; the
    gathered matrix is checked on sampled rows against the independent rectangular-block kernel."""
    import torch
    import torch.distributed as dist

    from mlatom_b200 import dist as kd, synthetic as S

    ntr = 50000
    eq = S.equilibrium(NAT)
    xyz_h = S.geometries(eq, ntr, seed=1)
    e_h, _ = S.pair_potential(xyz_h, eq)
    xyz = torch.from_numpy(xyz_h).to(dev)
    xeq = engine.xeq(torch.from_numpy(eq).to(dev))
    m = ntr

    def build_rows(r0, r1, out, row_offset):
        engine.build_kernel_matrix(xyz, xeq, SIGMA, LAMBDA, LAMBDA, rows=(r0, r1), out=out, row_offset=row_offset)

    slabs = kd.triangle_row_slabs(m, world)
    r0, r1 = slabs[rank]
    # (1) assembly alone: this rank's slab into a trapezoid buffer, CUDA events, max over ranks
    buf = torch.empty((r1 - r0, kd.piece_width(r1, m)), dtype=torch.float64, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for it in range(4):
        barrier()
        ev[0].record()
        build_rows(r0, r1, buf, r0)
        ev[1].record()
        torch.cuda.synchronize()
        if it:
            ms.append(ev[0].elapsed_time(ev[1]))
    build_ms = max_over_ranks(float(np.mean(ms)))
    del buf
    if world > 1:  # open the NCCL point-to-point channels before timing (lazy connection setup is not the gather)
        tiny = torch.zeros(8, dtype=torch.float64, device=dev)
        if rank == 0:
            for r in range(1, world):
                dist.recv(tiny, r)
        else:
            dist.send(tiny, 0)
    # (2) assembly + gather, pipelined; device time on every rank, max over ranks
    total_ms, stats, k = [], {}, None
    for it in range(2):
        del k
        k = None
        torch.cuda.empty_cache()
        barrier()
        ev[0].record()
        k = kd.build_kernel_matrix_sharded(build_rows, m, dev, root=0, pieces=4, stats=stats)
        ev[1].record()
        torch.cuda.synchronize()
        total_ms.append(ev[0].elapsed_time(ev[1]))
    total = max_over_ranks(total_ms[-1])
    gbytes = stats.get("gather_bytes", 0)
    if world > 1:
        t = torch.tensor([float(gbytes if rank else 0)], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        gbytes = t.item()
    out = {"workload": f"cfg3: K 50k x 50k values (lower triangle, 4*M*(M+1) = {4.0 * m * (m + 1) / 1e9:.1f} GB), "
                       f"row slabs over {world} GPU(s), gathered onto rank 0",
           "assembly_ms_max_over_ranks": build_ms, "assembly_gb_per_s_aggregate": 4.0 * m * (m + 1) / build_ms * 1e-6,
           "assembly_plus_gather_ms": total if world > 1 else None, "gather_bytes_into_root": gbytes,
           "gather_ms_not_hidden": max(total - build_ms, 0.0) if world > 1 else None,
           "inbound_gb_per_s": (gbytes / (max(total - build_ms, 1e-3)) * 1e-6) if world > 1 else None,
           "nvlink_reference_gb_per_s": 770.0,
           "note": "slabs travel as trapezoids (columns j <= i only), in 4 equal-area pieces per rank, each sent while the "
                   "next is assembled; the root scatters them into its M x M matrix"}
    if rank == 0:
        rng = np.random.default_rng(11)
        rows = np.sort(rng.choice(ntr, size=32, replace=False))
        idx = torch.as_tensor(rows).to(dev)
        blk = engine.kernel_block(xyz[idx].contiguous(), xyz, xeq, SIGMA)
        worst = 0.0
        for n, i in enumerate(rows):
            if i:
                worst = max(worst, (k[i, :i] - blk[n, :i]).abs().max().item())
        out["sampled_rows_max_abs_difference_vs_kernel_block"] = worst
        y = torch.from_numpy(e_h - e_h.mean()).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        alpha = engine.solve(k, y)
        torch.cuda.synchronize()
        out["potrf_potrs_ms_timed_separately"] = 1e3 * (time.perf_counter() - t0)
        res = blk @ alpha + LAMBDA * alpha[idx] - y[idx]
        out["residual_rel_sampled_rows"] = (res.norm() / y[idx].norm()).item()
    del k
    torch.cuda.empty_cache()
    return out


def cfg4_strong(engine, dev, world, rank, barrier, max_over_ranks):
    """BASELINE configs[3]: 10 M aspirin-sized (21-atom) geometries IN TOTAL, split over the ranks (strong scaling),
    streamed from pinned host memory through PackedModel.predict_host (H2D, kernel, D2H overlapped on three streams);
    Ntrain = 10 000 (assumed), alpha from a real solve on rank 0."""
    import torch
    import torch.distributed as dist

    from mlatom_b200 import synthetic as S

    nat, ntr, total, sigma = 21, 10000, 10_000_000, 2.0
    d = nat * (nat - 1) // 2
    eq = S.equilibrium(nat)
    xtr_h = S.geometries(eq, ntr, seed=1)
    e_h, _ = S.pair_potential(xtr_h, eq)
    xtr = torch.from_numpy(xtr_h).to(dev)
    xeq = engine.xeq(torch.from_numpy(eq).to(dev))
    alpha = torch.empty(ntr, dtype=torch.float64, device=dev)
    if rank == 0:
        k = engine.build_kernel_matrix(xtr, xeq, sigma, LAMBDA, LAMBDA)
        alpha.copy_(engine.solve(k, torch.from_numpy(e_h - e_h.mean()).to(dev)))
        del k
    if world > 1:
        dist.broadcast(alpha, src=0)
    model = engine.PackedModel.from_training_set(xtr, xeq, alpha, sigma, float(e_h.mean()), ntr, 0)
    lo, hi = rank * total // world, (rank + 1) * total // world
    n = hi - lo
    # this rank's shard, generated on the device in slices and parked in pinned host memory (the workload's input)
    x_host = torch.empty((n, nat, 3), dtype=torch.float64, pin_memory=True)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    d_eq = torch.from_numpy(eq).to(dev)
    for a in range(0, n, 1 << 20):
        b = min(n, a + (1 << 20))
        x_host[a:b].copy_(d_eq + 0.05 * torch.randn((b - a, nat, 3), dtype=torch.float64, device=dev, generator=gen))
    out_e = torch.empty(n, dtype=torch.float64, pin_memory=True)
    out_g = torch.empty((n, nat, 3), dtype=torch.float64, pin_memory=True)
    torch.cuda.synchronize()
    host = engine.HostPredictor(model)
    warm = min(n, 2 * host.chunk)
    host.run(x_host[:warm], True, True, out_e[:warm], out_g[:warm])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    t0 = time.perf_counter()
    ev[0].record()
    host.run(x_host, True, True, out_e, out_g)
    ev[1].record()
    barrier()
    ms = max_over_ranks(max(ev[0].elapsed_time(ev[1]), 1e3 * (time.perf_counter() - t0)))
    flop = float(total) * ntr * (5 * d + 4) + float(total) * 30 * d
    return {"workload": "cfg4: KREG E+F screening of 10M 21-atom geometries in total (strong scaling), Ntrain=10000, "
                        "pinned host xyz -> pinned host E, dE/dxyz",
            "geometries_total": total, "geometries_per_gpu": n, "ms": ms, "value": total / ms * 1e3, "unit": UNIT,
            "scaling": "strong", "h2d_bytes_per_gpu": n * nat * 24, "d2h_bytes_per_gpu": n * (8 + nat * 24),
            "algorithmic_tflops_aggregate": flop / ms * 1e-9, "checksum": float(out_e[:1000].sum()),
            "finite": bool(torch.isfinite(out_e).all().item())}


def cfg1_side_by_side(engine, dev):
    """BASELINE configs[0]: energies only, Ntrain=1000, Ntest=10k -- train + predict, CPU port in full beside the GPU."""
    import torch

    from mlatom_b200 import synthetic as S
    from oracle import kreg_oracle as O

    ntr, nte = 1000, 10000
    eq = S.equilibrium(NAT)
    xtr, xte = S.geometries(eq, ntr, 1), S.geometries(eq, nte, 2)
    e_tr, _ = S.pair_potential(xtr, eq)
    prior = float(e_tr.mean())
    # CPU: the oracle's literal restatement (K assembly, LAPACK Cholesky, MLatomF prediction loop), all threads
    O.set_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    xeq = O.xeq(eq)
    X, Xte = O.descriptors(xtr, xeq), O.descriptors(xte, xeq)
    k = O.kernel_matrix(xtr, X, SIGMA, ntr, 0)
    alpha_cpu = O.solve(k, e_tr - prior, ntr, 0, LAMBDA, LAMBDA)
    t1 = time.perf_counter()
    e_cpu, _ = O.predict_mlatomf(xtr, X, alpha_cpu, SIGMA, prior, ntr, 0, xte, Xte, True, False)
    t2 = time.perf_counter()
    # GPU: host numpy in, host numpy out (copies included)
    to_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    res, trials = {}, []
    for it in range(6):
        torch.cuda.synchronize()
        g0 = time.perf_counter()
        d_xeq = engine.xeq(to_dev(eq))
        d_xtr = to_dev(xtr)
        kk = engine.build_kernel_matrix(d_xtr, d_xeq, SIGMA, LAMBDA, LAMBDA)
        a = engine.solve(kk, to_dev(e_tr - prior))
        model = engine.PackedModel.from_training_set(d_xtr, d_xeq, a, SIGMA, prior, ntr, 0)
        torch.cuda.synchronize()
        g1 = time.perf_counter()
        e_gpu, _ = model.predict(to_dev(xte), True, False)
        e_host = e_gpu.cpu().numpy()
        g2 = time.perf_counter()
        if it:  # the first pass creates the cuSOLVER handle and the allocator's blocks
            trials.append((1e3 * (g1 - g0), 1e3 * (g2 - g1)))
    res = {"train_ms": float(np.median([t[0] for t in trials])), "predict_ms": float(np.median([t[1] for t in trials])),
           "trials_train_ms": [round(t[0], 3) for t in trials], "timing": "host wall clock, median of 5 passes after one warm-up"}
    return {"workload": "cfg1: KREG energies only, 9 atoms, Ntrain=1000, Ntest=10000, train + predict, host arrays in/out",
            "gpu": res, "cpu_port": {"train_ms": 1e3 * (t1 - t0), "predict_ms": 1e3 * (t2 - t1), "cores": os.cpu_count()},
            "max_abs_energy_difference": float(np.abs(e_host - e_cpu).max()),
            "max_rel_energy_difference": float(np.abs(e_host - e_cpu).max() / np.abs(e_cpu).max())}


def kreg_so_baseline(eq, xyz_tr, alpha, seconds=6.0):
    """The reference AS SHIPPED: its prebuilt KREG.so (the default 'KREG_API' backend; kreg_mp_predict_,
    KREG.f90:534-596) through the oracle/_ref shims, all host threads, bounded sample.  None if oracle/_ref is absent."""
    from mlatom_b200 import synthetic as S
    from oracle import ref_kreg as R

    if not R.available():
        return None
    threads = os.cpu_count() or 1
    R.set_threads(threads)
    xeq = R.get_xeq(eq)
    x_tr = R.calc_re_descriptors(xyz_tr, xeq)

    def run(n):
        xte = S.geometries(eq, n, seed=2)
        t0 = time.perf_counter()
        x_te = R.calc_re_descriptors(xte, xeq)
        R.predict(xyz_tr, x_tr, xte, x_te, alpha, SIGMA, NTRAIN, 0, True, True)
        return time.perf_counter() - t0

    n0 = threads
    t = run(n0)
    n = int(max(n0, min(NTEST, n0 * seconds / max(t, 1e-6)))) // threads * threads
    t = run(max(n, threads))
    return {"value": max(n, threads) / t, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{max(n, threads)} of {NTEST} test geometries, full Ntrain={NTRAIN}, {t:.1f} s wall",
            "what": "reference's prebuilt mlatom/fortran/np2/KREG.so (ifort), E and gradient passes as shipped"}


def run_reference(args):
    """--impl reference: the reference algorithm on the host CPU (MLatomF itself cannot be built here:
    no Fortran compiler, binary absent from the checkout)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import kreg_oracle as O

    eq, xyz_tr, e_tr = training_set()
    prior = float(e_tr.mean())
    alpha = np.random.default_rng(0).standard_normal(NTRAIN) * 1e-3  # throughput does not depend on alpha
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    from mlatom_b200 import synthetic as S

    _, n, _ = cpu_baseline(eq, xyz_tr, alpha, prior, seconds=per_step, threads=threads)  # calibrates the sample
    xeq = O.xeq(eq)
    x_tr = O.descriptors(xyz_tr, xeq)
    xte = S.geometries(eq, n, seed=2)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        x_te = O.descriptors(xte, xeq)
        O.predict_mlatomf(xyz_tr, x_tr, alpha, SIGMA, prior, NTRAIN, 0, xte, x_te, True, True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    sample = f"{n} of {NTEST} test geometries per step, full Ntrain={NTRAIN}"
    shipped = kreg_so_baseline(eq, xyz_tr, alpha)  # the reference's own binary beside the port (slower: literal loops)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(),
        "notes": {"arm": "CPU only: MLatomF's optimised prediction loop restated in C/OpenMP (oracle/kreg_oracle.c), "
                         "all host threads; geometries/s is independent of the sample size.  The port is the arm "
                         "because it is the FASTER of the two reference programs (BASELINE metric: 'vs MLatomF CPU'); "
                         "the reference's shipped KREG.so, run through oracle/_ref, is reported beside it"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "cpu_baseline_reference_as_shipped": shipped,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mlatom_b200 import engine, synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback exists)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG=VERSION (this image's default) makes NCCL print its version on STDOUT, next to the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # whatever NCCL still prints must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    engine.device_check()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    to_dev = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

    # ---- model: assemble K and solve on rank 0's GPU, broadcast alpha (the only collective) ----
    eq, xyz_tr, e_tr = training_set()
    prior = float(e_tr.mean())
    xeq = engine.xeq(to_dev(eq))
    d_xyz_tr = to_dev(xyz_tr)
    alpha = torch.empty(NTRAIN, dtype=torch.float64, device=dev)
    extra = {}
    if rank == 0:
        k = torch.empty((NTRAIN, NTRAIN), dtype=torch.float64, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        kb = []
        for it in range(6):
            torch.cuda.synchronize()
            ev[0].record()
            engine.build_kernel_matrix(d_xyz_tr, xeq, SIGMA, LAMBDA, LAMBDA, out=k)
            ev[1].record()
            torch.cuda.synchronize()
            if it >= 2:
                kb.append(ev[0].elapsed_time(ev[1]))
        tri_bytes = 4.0 * NTRAIN * (NTRAIN + 1)
        kb_ms = float(np.mean(kb))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        alpha.copy_(engine.solve(k, to_dev(e_tr - prior)))
        torch.cuda.synchronize()
        extra["solve_ms"] = 1e3 * (time.perf_counter() - t0)
        del k
        extra["kbuild"] = {"ms": kb_ms, "bytes": tri_bytes, "layout": "lower triangle, 4*M*(M+1) bytes",
                           "gb_per_s": tri_bytes / kb_ms * 1e-6,
                           "flop": NTRAIN * (NTRAIN + 1) / 2 * (3 * D + 3)}
    if world > 1:
        dist.broadcast(alpha, src=0)
    model = engine.PackedModel.from_training_set(d_xyz_tr, xeq, alpha, SIGMA, prior, NTRAIN, 0)

    # ---- FP64 pipe peak, measured live (roofline denominator) ----
    peak_dmma = engine.measure_fp64_peak("dmma", reps=5, ms_target=30.0)
    peak_dfma = engine.measure_fp64_peak("dfma", reps=3, ms_target=30.0)

    # ---- K assembly at the BASELINE sizes: before the long prediction runs, on a GPU in the same state as a training job
    # would find it (the gradient-block kernels above 10 atoms are SM-bound and read 5-10 % lower after seconds of
    # full-power FP64 tensor work); clocks sampled over this region too ----
    kb_cfgs = None
    if world == 1 and args.workload == "cfg2":
        hbm_peak0 = 6650.0
        try:
            hbm_peak0 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        ksampler = ClockSampler(local_rank)
        ksampler.start()
        kb_cfgs = kbuild_configs(engine, dev, hbm_peak0, peak_dmma)
        kb_cfgs["clocks"] = ksampler.stop()

    # ---- this rank's test batch, resident in HBM ----
    xte_host = torch.from_numpy(S.geometries(eq, NTEST, seed=2 + rank)).pin_memory()
    d_xte = xte_host.to(dev, non_blocking=True)
    d_e = torch.empty(NTEST, dtype=torch.float64, device=dev)
    d_g = torch.empty((NTEST, NAT, 3), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def step():
        model.predict(d_xte, True, True, out_e=d_e, out_g=d_g)

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    value = world * NTEST * args.steps / (elapsed_ms * 1e-3)
    kernel_ms = elapsed_ms / args.steps  # one kernel launch per step, back to back on one stream

    # ---- end to end through the host-buffer API (copies inside the timed region) ----
    host = engine.HostPredictor(model)
    out_e = torch.empty(NTEST, dtype=torch.float64).pin_memory()
    out_g = torch.empty((NTEST, NAT, 3), dtype=torch.float64).pin_memory()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):
        host.run(xte_host, True, True, out_e, out_g)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(e2e_steps):
        host.run(xte_host, True, True, out_e, out_g)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)))
    e2e_value = world * NTEST * e2e_steps / (e2e_ms * 1e-3)
    checksum = float(out_e[:1000].sum())  # the device->host read of the step's result
    d_e_keep = None
    parity_n = 200000  # the CPU baseline's sample never exceeds this many geometries (12 s at ~9 k geometries/s)
    e_keep_host, g_keep_host = d_e[:parity_n].cpu().numpy(), d_g[:parity_n].cpu().numpy()
    del out_e, out_g, host

    # ---- the two multi-GPU configurations north_star names beside the headline (every rank takes part) ----
    sharded_blocks = {}
    if args.workload == "cfg2" and not args.no_extra:
        del d_xte, d_g, d_e_keep
        torch.cuda.empty_cache()
        sharded_blocks["kbuild_sharded"] = kbuild_sharded(engine, dev, world, rank, barrier, max_over_ranks)
        sharded_blocks["cfg4_strong"] = cfg4_strong(engine, dev, world, rank, barrier, max_over_ranks)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    achieved = FLOP_PER_STEP / (kernel_ms * 1e-3) * 1e-12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": shared_config(),
        "notes": {"arm": "alpha solved by cuSOLVER potrf on rank 0 and broadcast once; "
                         f"test batches sharded over {world} GPU(s)",
                  "l2": f"inputs+outputs {NTEST * (NAT * 48 + 8) / 1e6:.0f} MB per step per GPU exceed the 126 MB L2; no flush needed"},
        "clocks": clocks,
        "gpu_launches": args.steps,
        "roofline": {"bound": "tensor", "pipe": "FP64 tensor core (DMMA.8x8x4); shares the FP64 pipe with DFMA",
                     "achieved": achieved, "peak": peak_dmma, "unit": "TFLOP/s", "frac": achieved / peak_dmma,
                     "traffic": TRAFFIC,
                     "traffic_source": TRAFFIC_SOURCE,
                     "algorithmic_io_bytes_per_launch": NTEST * (NAT * 3 * 8 * 2 + 8),
                     "peak_source": "measured live by kreg_measure_fp64_peak(DMMA) (MEASURED_PEAKS.json has no FP64 figure)",
                     "peak_dfma": peak_dfma, "algorithmic_flop_per_launch": FLOP_PER_STEP,
                     "kernel": KERNEL, "kernel_ms": kernel_ms,
                     "note": "achieved counts SURVEY.md 8d's algorithmic flops, (5D+4) per pair; the two-GEMM formulation "
                             "executes (4D+2) + exp per pair, so at large D the fraction can exceed the executed-flop share"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": NTEST * NAT * 3 * 8,
                "d2h_bytes_per_step": NTEST * 8 + NTEST * NAT * 3 * 8, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "api": "PackedModel.predict_host (pinned host xyz -> pinned host E, dE/dxyz; 2-wave chunks, 3 streams)",
                "checksum": checksum},
    }
    if "kbuild" in extra:
        kb = extra["kbuild"]
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        fp64_ms = kb["flop"] / (peak_dmma * 1e12) * 1e3
        line["kbuild"] = {"metric": "K-matrix build GB/s (values block, Ntrain=10k, lower triangle)", "value": kb["gb_per_s"],
                          "unit": "GB/s", "ms": kb["ms"], "bytes": kb["bytes"],
                          "hbm_peak_gbs": hbm_peak if hbm_peak else 6650.0,
                          "hbm_peak_source": "MEASURED_PEAKS.json" if hbm_peak else "fallback",
                          "frac_of_hbm": kb["gb_per_s"] / (hbm_peak if hbm_peak else 6650.0),
                          "fp64_bound_ms": fp64_ms, "frac_of_fp64_bound": fp64_ms / kb["ms"],
                          "note": f"values-only K at D={D} is FP64-issue-bound, not HBM-bound (SURVEY.md App. D); this block is the "
                                  "model of the headline run (Ntrain=10k: 2.7 waves of tiles + three preparation kernels inside "
                                  "0.22 ms) -- the BASELINE sizes (cfg3 50k values, cfg5 56k gradient-augmented) are in kbuild_configs",
                          "potrf_potrs_ms_timed_separately": extra.get("solve_ms")}
    if world == 1 and args.workload == "cfg2":
        if kb_cfgs is not None:
            line["kbuild_configs"] = kb_cfgs
        line["step_latency"] = step_latency(engine, model, dev)
        line["md_steps_per_s"] = md_steps_per_s(engine, model, dev)
    if world == 1 and not args.no_cpu:
        base, n_cpu, _ = cpu_baseline(eq, xyz_tr, alpha.cpu().numpy(), prior, seconds=12.0)
        e_cpu, g_cpu = base.pop("results")
        line["cpu_baseline"] = base
        # parity of the timed run itself: the CPU port predicted the first n_cpu geometries of this rank's test batch
        # with the same alpha -- diff them against what the timed kernel left in d_e / d_g
        n_cpu = min(n_cpu, parity_n)
        e_cpu, g_cpu = e_cpu[:n_cpu], g_cpu[:n_cpu]
        e_gpu, g_gpu = e_keep_host[:n_cpu], g_keep_host[:n_cpu]
        line["parity"] = {
            "against": "cpu_baseline (oracle port of MLatomF's loop, A_KRR_kernel.f90:327-412), same alpha, same geometries",
            "geometries": int(n_cpu), "ntrain": NTRAIN,
            "max_rel_energy": float(np.abs(e_gpu - e_cpu).max() / np.abs(e_cpu).max()),
            "max_abs_gradient": float(np.abs(g_gpu - g_cpu).max()),
            "max_abs_gradient_reference": float(np.abs(g_cpu).max()),
            "tolerance": {"energy_rel": 1e-10, "gradient_abs": 1e-8},
        }
        line["parity"]["ok"] = bool(line["parity"]["max_rel_energy"] <= 1e-10 and line["parity"]["max_abs_gradient"] <= 1e-8)
        shipped = kreg_so_baseline(eq, xyz_tr, alpha.cpu().numpy())
        if shipped:
            line["cpu_baseline_reference_as_shipped"] = shipped
        if args.workload == "cfg2":
            line["cfg1"] = cfg1_side_by_side(engine, dev)
    line.update(sharded_blocks)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the kbuild_sharded / cfg4_strong blocks (profiling runs)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"],
                    help="cfg2 (default, the configuration the metric is quoted on) or cfg4 (10M 21-atom geometries)")
    args = ap.parse_args()
    set_workload(args.workload, int(os.environ.get("WORLD_SIZE", "1")))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
