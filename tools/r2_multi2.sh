set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_baseline_parity.py::test_non_current_device" -m gpu -q -s > gpurun_out/r2_tests_multi2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_multi2.log
tail -15 gpurun_out/r2_tests_multi2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
