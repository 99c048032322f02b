# Round-2 final measurement pass (one GPU): smoke, bench line, launch list of the same command, throughput of the
# radial-kernel and permutationally invariant models, ncu --set full of the radial GEMM1 (EPI_WR) launch.
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -6 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err || exit 1
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'predict_kernel|predict_finish|values_|prep_|chunk_rows|pair_terms|grad_emit|grad_fused|jacobian_rows|descriptor|xeq_kernel|pack_model|precompute|probe|md_' -c 900 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_launch.log 2>&1
timeout 300 python tools/bench_radial.py > gpurun_out/r02_radial.json 2> gpurun_out/r02_radial.err; cat gpurun_out/r02_radial.json
timeout 300 python tools/bench_perminv.py > gpurun_out/r02_perminv.json 2>&1; cat gpurun_out/r02_perminv.json
timeout 600 ncu --set full --clock-control none -k regex:'values_ring_kernel' -s 0 -c 1 -o gpurun_out/r02_prof_radial_gemm1 -f python tools/bench_radial.py > gpurun_out/ncu_radial.log 2>&1
ls -la gpurun_out/ | tail -12
tail -c 400 gpurun_out/r02_bench_final.json
