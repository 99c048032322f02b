# Quick check of a gradient-block kernel change: parity of the gradient-augmented assembly, then the size sweep.
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_molecules.py -m gpu -x -q -k "grad or kernel_matrix or K or size or slab or molecule" 2>&1 | tail -4
run() { timeout 300 python tools/bench_kbuild.py "$@" --grads --no-solve --reps 5 > gpurun_out/tmp.json 2>&1; echo "$*: $(grep -o '"build_ms_max_over_ranks": [0-9.]*' gpurun_out/tmp.json) $(grep -o '"build_gb_per_s": [0-9.]*' gpurun_out/tmp.json)"; }
for nat in ${SWEEP_NATS:-11 12 14 16 18 21 24 28 32 48 64}; do
  ntr=$((54000 / (3*nat)))
  run --natoms $nat --ntrain $ntr
done 2>&1 | tee gpurun_out/r02_gradk_sweep_quick.txt
