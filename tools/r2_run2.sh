set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_large_molecules.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests2.log
tail -25 gpurun_out/r2_tests2.log
for nat in 9 12 21; do
  ntr=$((54000 / (3*nat)))
  for rt in 0 1; do
    KREG_EMIT_RT=$rt timeout 300 python tools/bench_kbuild.py --natoms $nat --ntrain $ntr --grads --no-solve --reps 3 > gpurun_out/r2_kb_rt${rt}_$nat.json 2>&1
    echo "nat=$nat emit_rt=$rt: $(grep -o '"build_ms_max_over_ranks": [0-9.]*' gpurun_out/r2_kb_rt${rt}_$nat.json) $(grep -o '"build_gb_per_s": [0-9.]*' gpurun_out/r2_kb_rt${rt}_$nat.json)"
  done
done
KREG_EMIT_RT=1 KREG_PAIR_RT=1 timeout 300 python tools/bench_kbuild.py --natoms 21 --ntrain 857 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'
timeout 300 python tools/bench_kbuild.py --natoms 32 --ntrain 560 --grads --no-solve --reps 3 2>&1 | tail -1
timeout 300 python tools/bench_kbuild.py --natoms 9 --ntrain 50000 --no-solve --reps 3 2>&1 | tail -1
timeout 300 python tools/bench_kbuild.py --natoms 21 --ntrain 20000 --no-solve --reps 3 2>&1 | tail -1
timeout 600 python tools/bench_predict.py --natoms 32 --ntrain 10000 --ntest 75776 2>&1 | tail -3
timeout 600 python tools/bench_predict.py --natoms 32 --ntrain 10000 --ntest 75776 --grads 2>&1 | tail -3
