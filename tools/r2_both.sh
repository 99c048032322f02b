cd $GRAFT_REPO_ROOT
bash tools/r2_gradk_sweep.sh
bash tools/r2_debug_bounds.sh 2>&1 | tail -8
