# Round-2 measurement pass, part A (one GPU): bench line, launch list of the same command, ncu --set full of the prediction
# kernel and the values kernel.  Part B: tools/final_profile_r2b.sh (gradient blocks, 32-atom prediction, size sweep).
# (Two calls: what a call writes under gpurun_out/ must stay below 64 MiB.)
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || exit 1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'predict_kernel|predict_finish|values_|prep_|chunk_rows|pair_terms|grad_emit|jacobian_rows|descriptor|xeq_kernel|pack_model|precompute|probe|md_' -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'predict_kernel' -s 3 -c 1 -o gpurun_out/r02_prof_predict -f python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_predict.log 2>&1
timeout 200 python tools/bench_kbuild.py --ntrain 50000 --no-solve --reps 1 > gpurun_out/kv_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:'values_ring' -s 1 -c 1 -o gpurun_out/r02_prof_kvalues50k -f python tools/bench_kbuild.py --ntrain 50000 --no-solve --reps 1 > gpurun_out/ncu_kv.log 2>&1
ls -la gpurun_out/
tail -c 600 gpurun_out/r02_bench_final.json
