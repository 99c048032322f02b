cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_parity.py -m gpu -q > gpurun_out/r2_tests18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests18.log
tail -4 gpurun_out/r2_tests18.log
for i in 1 2 3; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'lat', d.get('step_latency',{}).get('batch_1'))"; done
for n in 6 12 21 24; do timeout 300 python tools/bench_predict.py --natoms $n --ntrain 10000 --ntest 200000 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['natoms'], d['ms'], d['algorithmic_tflops'])"; done
