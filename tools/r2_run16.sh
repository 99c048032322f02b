set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_other_kernels.py tests/test_gpu_api.py tests/test_gpu_large_molecules.py -m gpu -q > gpurun_out/r2_tests16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests16.log
tail -12 gpurun_out/r2_tests16.log
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
for n in "11 1636" "12 1500" "16 1125" "21 857" "24 750" "28 643" "29 620" "32 560" "48 375" "64 280"; do set -- $n; echo "nat=$1 final: $(kb $1 $2)"; done
