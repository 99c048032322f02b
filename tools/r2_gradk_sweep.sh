# Round-2 gradient-augmented K: parity suite, write-only HBM ceiling, assembly sweep over molecule sizes
# with the engine's K layout and with a dense torch.empty((M, M)), ncu --set full of the two gradient-block kernels.
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_fill.py | tee gpurun_out/r02_fill.json
run() { timeout 300 python tools/bench_kbuild.py "$@" --grads --no-solve --reps 5 > gpurun_out/tmp.json 2>&1; echo "$*: $(grep -o '"build_ms_max_over_ranks": [0-9.]*' gpurun_out/tmp.json) $(grep -o '"build_gb_per_s": [0-9.]*' gpurun_out/tmp.json)"; }
for nat in 6 9 10 11 12 14 16 18 21 24 28 32 48 64; do
  ntr=$((54000 / (3*nat)))
  run --natoms $nat --ntrain $ntr --plain
  run --natoms $nat --ntrain $ntr
done 2>&1 | tee gpurun_out/r02_gradk_sweep.txt
for nat in 21 9; do
  ntr=$((54000 / (3*nat)))
  timeout 600 ncu --set full --clock-control none -k regex:'grad_fused|grad_emit|pair_terms' -s $((nat > 10 ? 1 : 2)) -c 2 -o gpurun_out/r02_prof_gradk${nat}_final -f python tools/bench_kbuild.py --natoms $nat --ntrain $ntr --grads --no-solve --reps 1 > gpurun_out/ncu_gradk$nat.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
