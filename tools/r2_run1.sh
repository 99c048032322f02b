set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests1.log
tail -30 gpurun_out/r2_tests1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
for nat in 9 12 21; do
  ntr=$((54000 / (3*nat))); timeout 300 python tools/bench_kbuild.py --natoms $nat --ntrain $ntr --grads --no-solve --reps 3 > gpurun_out/r2_kb_base_$nat.json 2>&1
done
tail -c 1500 gpurun_out/r2_bench1.json
cat gpurun_out/r2_kb_base_*.json | tail -c 2000
