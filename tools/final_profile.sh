# Round-end measurement pass (one GPU): bench line, launch list of the same command, ncu --set full of the K kernels.
set -x
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'predict_kernel|predict_finish|values_|prep_|chunk_rows|pair_terms|grad_emit|jacobian_rows|descriptor|xeq_kernel|pack_model|precompute|probe' -c 800 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 200 python tools/bench_kbuild.py --ntrain 2000 --grads --no-solve --reps 1 > gpurun_out/kb_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_terms|grad_emit' -s 2 -c 2 -o gpurun_out/prof_gradk_final -f python tools/bench_kbuild.py --ntrain 2000 --grads --no-solve --reps 1 > gpurun_out/ncu_gradk.log 2>&1
tail -c 600 gpurun_out/bench_final.json
