set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_large_molecules.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests4.log
tail -8 gpurun_out/r2_tests4.log
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
export KREG_EMIT_RT=1
echo "nat=9 default: $(kb 9 2000)"
echo "nat=10 default: $(kb 10 1800)"
echo "nat=11 default: $(kb 11 1636)"
echo "nat=12 default: $(kb 12 1500)"
echo "nat=16 default: $(kb 16 1125)"
echo "nat=21 default: $(kb 21 857)"
echo "nat=24 default: $(kb 24 750)"
echo "nat=32 default: $(kb 32 560)"
echo "nat=48 default: $(kb 48 375)"
for cfg in "4 2 8" "4 1 8" "2 2 8" "3 2 8" "4 2 32" "4 1 32" "2 1 8"; do
  set -- $cfg
  echo "nat=21 TJ=$1 NBUF=$2 NJB=$3: $(KREG_RT_TJ=$1 KREG_RT_NBUF=$2 KREG_RT_NJB=$3 kb 21 857)"
done
for cfg in "7 2 8" "7 1 8" "4 2 8" "6 2 8" "7 2 32" "5 2 8"; do
  set -- $cfg
  echo "nat=12 TJ=$1 NBUF=$2 NJB=$3: $(KREG_RT_TJ=$1 KREG_RT_NBUF=$2 KREG_RT_NJB=$3 kb 12 1500)"
done
for cfg in "9 2 8" "9 1 8" "6 2 8" "8 2 8"; do
  set -- $cfg
  echo "nat=9 TJ=$1 NBUF=$2 NJB=$3: $(KREG_RT_TJ=$1 KREG_RT_NBUF=$2 KREG_RT_NJB=$3 kb 9 2000)"
done
echo "nat=21 pair_rt: $(KREG_PAIR_RT=1 kb 21 857)"
timeout 200 python tools/bench_kbuild.py --natoms 21 --ntrain 857 --grads --no-solve --reps 1 > gpurun_out/kb_plain21.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'grad_emit' -s 1 -c 1 -o gpurun_out/r2_prof_gradk21_rt2 -f python tools/bench_kbuild.py --natoms 21 --ntrain 857 --grads --no-solve --reps 1 > gpurun_out/ncu_gradk21.log 2>&1
tail -3 gpurun_out/ncu_gradk21.log
