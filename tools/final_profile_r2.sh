# Round-2 measurement pass (one GPU): bench line, launch list of the same command, ncu --set full of the dominant kernels.
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || exit 1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'predict_kernel|predict_finish|values_|prep_|chunk_rows|pair_terms|grad_emit|jacobian_rows|descriptor|xeq_kernel|pack_model|precompute|probe' -c 800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'predict_kernel' -s 3 -c 1 -o gpurun_out/r02_prof_predict -f python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_predict.log 2>&1
timeout 200 python tools/bench_kbuild.py --ntrain 50000 --no-solve --reps 1 > gpurun_out/kv_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'values_ring' -s 1 -c 1 -o gpurun_out/r02_prof_kvalues50k -f python tools/bench_kbuild.py --ntrain 50000 --no-solve --reps 1 > gpurun_out/ncu_kv.log 2>&1
for nat in 9 21; do
  ntr=$((54000 / (3*nat)))
  timeout 200 python tools/bench_kbuild.py --natoms $nat --ntrain $ntr --grads --no-solve --reps 1 > gpurun_out/kb_plain$nat.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_terms|grad_emit' -s 2 -c 2 -o gpurun_out/r02_prof_gradk$nat -f python tools/bench_kbuild.py --natoms $nat --ntrain $ntr --grads --no-solve --reps 1 > gpurun_out/ncu_gradk$nat.log 2>&1
done
timeout 300 python tools/bench_predict.py --natoms 32 --ntrain 10000 --ntest 18944 --reps 1 > gpurun_out/bp32_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'values_ring' -s 4 -c 2 -o gpurun_out/r02_prof_predict32 -f python tools/bench_predict.py --natoms 32 --ntrain 10000 --ntest 18944 --reps 1 > gpurun_out/ncu_bp32.log 2>&1
for n in 6 12 16 21 24 28 32 40 48; do timeout 300 python tools/bench_predict.py --natoms $n --ntrain 10000 --ntest $((n<=24 ? 200000 : 75776)) >> gpurun_out/r02_predict_sweep.txt 2>&1; timeout 300 python tools/bench_predict.py --natoms $n --ntrain 10000 --ntest $((n<=24 ? 200000 : 75776)) --grads >> gpurun_out/r02_predict_sweep.txt 2>&1; done
tail -c 600 gpurun_out/r02_bench_final.json
