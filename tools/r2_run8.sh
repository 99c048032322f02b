set -x
cd $GRAFT_REPO_ROOT
for v in njt2 clamp zeroinit; do echo "== variant $v"; KREG_LIB=$PWD/tools/bin/lib_$v.so python tools/debug_kvals.py 5000 2>&1 | tail -10; done
echo "== default"; python tools/debug_kvals.py 5000 2>&1 | tail -10
echo "== sanitizer racecheck"
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/debug_kvals.py 5000 > gpurun_out/r2_racecheck.log 2>&1; tail -30 gpurun_out/r2_racecheck.log
