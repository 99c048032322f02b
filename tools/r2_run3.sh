set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_large_molecules.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests3.log
tail -8 gpurun_out/r2_tests3.log
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
for cfg in "2 4 1 8" "2 4 2 8" "2 2 2 8" "1 4 1 8" "1 4 2 8" "2 6 1 8" "2 8 1 8" "2 4 1 32" "2 4 2 32" "1 2 2 16"; do
  set -- $cfg
  echo "nat=21 NV=$1 TJ=$2 NBUF=$3 NJB=$4: $(KREG_EMIT_RT=1 KREG_RT_NV=$1 KREG_RT_TJ=$2 KREG_RT_NBUF=$3 KREG_RT_NJB=$4 kb 21 857)"
done
for cfg in "2 14 2 8" "2 8 2 8" "2 6 2 8" "2 8 1 8" "1 6 2 8" "2 14 1 8" "2 8 2 32" "2 4 2 16"; do
  set -- $cfg
  echo "nat=12 NV=$1 TJ=$2 NBUF=$3 NJB=$4: $(KREG_EMIT_RT=1 KREG_RT_NV=$1 KREG_RT_TJ=$2 KREG_RT_NBUF=$3 KREG_RT_NJB=$4 kb 12 1500)"
done
for cfg in "2 2 1 8" "2 2 2 8" "1 2 1 8" "2 4 1 8"; do
  set -- $cfg
  echo "nat=32 NV=$1 TJ=$2 NBUF=$3 NJB=$4: $(KREG_RT_NV=$1 KREG_RT_TJ=$2 KREG_RT_NBUF=$3 KREG_RT_NJB=$4 kb 32 560)"
done
export KREG_EMIT_RT=1
timeout 200 python tools/bench_kbuild.py --natoms 21 --ntrain 857 --grads --no-solve --reps 1 > gpurun_out/kb_plain21.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_terms|grad_emit' -s 2 -c 2 -o gpurun_out/r2_prof_gradk21_rt -f python tools/bench_kbuild.py --natoms 21 --ntrain 857 --grads --no-solve --reps 1 > gpurun_out/ncu_gradk21.log 2>&1
tail -3 gpurun_out/ncu_gradk21.log
