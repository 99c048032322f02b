set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_tests14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests14.log
tail -12 gpurun_out/r2_tests14.log
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
for n in "6 3000" "9 2000" "10 1800" "11 1636" "12 1500" "16 1125" "21 857" "22 818" "23 782" "24 750" "32 560" "48 375" "64 280"; do set -- $n; echo "nat=$1 final: $(kb $1 $2)"; done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench14.json
tail -5 gpurun_out/r2_bench14.err
