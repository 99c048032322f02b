# Round-2 closing pass (one GPU) over the final library: whole GPU suite, smoke, bench line.
set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02_gputests_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputests_final.log
tail -3 gpurun_out/r02_gputests_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log; tail -4 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r02_bench_final.json
