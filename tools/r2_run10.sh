cd $GRAFT_REPO_ROOT
KREG_LIB=$PWD/tools/bin/lib_biasdbg.so python tools/debug_bias.py 2>&1 | tail -5
