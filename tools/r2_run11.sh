cd $GRAFT_REPO_ROOT
echo "== gbias"; KREG_LIB=$PWD/tools/bin/lib_gbias.so python tools/debug_kvals.py 5000 2>&1 | grep -E "bad|tj%4"
echo "== biasdbg K check"; KREG_LIB=$PWD/tools/bin/lib_biasdbg.so python tools/debug_kvals.py 5000 2>&1 | grep -E "bad|tj%4"
