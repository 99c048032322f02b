set -x
cd $GRAFT_REPO_ROOT
python tools/debug_kvals.py 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_large_molecules.py tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests6.log
tail -8 gpurun_out/r2_tests6.log
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
export KREG_EMIT_RT=1
for n in "9 2000" "10 1800" "11 1636" "12 1500" "16 1125" "21 857" "24 750" "32 560"; do set -- $n; echo "nat=$1 templated: $(kb $1 $2)   generic: $(KREG_RT_GENERIC=1 kb $1 $2)"; done
for cfg in "4 16" "4 8" "3 16" "2 16" "4 4"; do
  set -- $cfg
  echo "nat=21 TJ=$1 NIB=$2: $(KREG_RT_TJ=$1 KREG_RT_NIB=$2 kb 21 857)"
done
for cfg in "7 16" "5 16" "4 16" "6 8" "3 16"; do
  set -- $cfg
  echo "nat=12 TJ=$1 NIB=$2: $(KREG_RT_TJ=$1 KREG_RT_NIB=$2 kb 12 1500)"
done
for cfg in "9 16" "8 16" "6 16" "4 16"; do
  set -- $cfg
  echo "nat=9 TJ=$1 NIB=$2: $(KREG_RT_TJ=$1 KREG_RT_NIB=$2 kb 9 2000)"
done
timeout 200 python tools/bench_kbuild.py --natoms 21 --ntrain 857 --grads --no-solve --reps 1 > gpurun_out/kb_plain21.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'grad_emit' -s 1 -c 1 -o gpurun_out/r2_prof_gradk21_rt4 -f python tools/bench_kbuild.py --natoms 21 --ntrain 857 --grads --no-solve --reps 1 > gpurun_out/ncu_gradk21.log 2>&1
tail -3 gpurun_out/ncu_gradk21.log
