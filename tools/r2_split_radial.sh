#!/bin/bash
# split launches of the radial-kernel models: the parity tests of that file, then the latency table
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_other_kernels.py -m gpu -x -q > gpurun_out/r2_split_tests.log 2>&1; echo "other_kernels rc=$?" >> gpurun_out/r2_split_tests.log
timeout 120 python tools/bench_radial_latency.py > gpurun_out/r2_radial_latency.json 2> gpurun_out/r2_radial_latency.err
tail -8 gpurun_out/r2_split_tests.log; cat gpurun_out/r2_radial_latency.json; tail -3 gpurun_out/r2_radial_latency.err
