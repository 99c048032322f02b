cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_molecules.py -m gpu -q > gpurun_out/r2_tests17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests17.log
tail -5 gpurun_out/r2_tests17.log
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 4 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
for n in "6 3000" "9 2000" "12 1500" "16 1125" "21 857" "24 750" "32 560" "48 375"; do set -- $n
  echo "nat=$1 chunks=1: $(KREG_GRAD_CHUNKS=1 kb $1 $2)  chunks=4: $(KREG_GRAD_CHUNKS=4 kb $1 $2)  chunks=8: $(KREG_GRAD_CHUNKS=8 kb $1 $2)  chunks=16: $(KREG_GRAD_CHUNKS=16 kb $1 $2)  chunks=32: $(KREG_GRAD_CHUNKS=32 kb $1 $2)"
done
