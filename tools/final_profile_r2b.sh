# Round-2 measurement pass, part B (one GPU): ncu --set full of the gradient-block kernels (9 and 21 atoms) and of the two
# GEMM launches of the 32-atom prediction path; prediction sweep over molecule sizes.
set -x
cd $GRAFT_REPO_ROOT
for nat in 9 21; do
  ntr=$((54000 / (3*nat)))
  timeout 200 python tools/bench_kbuild.py --natoms $nat --ntrain $ntr --grads --no-solve --reps 1 > gpurun_out/kb_plain$nat.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none -k regex:'pair_terms|grad_emit' -s 2 -c 2 -o gpurun_out/r02_prof_gradk$nat -f python tools/bench_kbuild.py --natoms $nat --ntrain $ntr --grads --no-solve --reps 1 > gpurun_out/ncu_gradk$nat.log 2>&1
done
timeout 300 python tools/bench_predict.py --natoms 32 --ntrain 10000 --ntest 18944 --reps 1 > gpurun_out/bp32_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:'values_ring' -s 4 -c 2 -o gpurun_out/r02_prof_predict32 -f python tools/bench_predict.py --natoms 32 --ntrain 10000 --ntest 18944 --reps 1 > gpurun_out/ncu_bp32.log 2>&1
rm -f gpurun_out/r02_predict_sweep.txt
for n in 6 12 16 21 24 28 32 40 48 64; do nt=$((n<=24 ? 200000 : 75776)); timeout 300 python tools/bench_predict.py --natoms $n --ntrain 10000 --ntest $nt >> gpurun_out/r02_predict_sweep.txt 2>&1; timeout 300 python tools/bench_predict.py --natoms $n --ntrain 10000 --ntest $nt --grads >> gpurun_out/r02_predict_sweep.txt 2>&1; done
ls -la gpurun_out/
cat gpurun_out/r02_predict_sweep.txt
