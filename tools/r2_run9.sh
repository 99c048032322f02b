cd $GRAFT_REPO_ROOT
for v in swap sleep; do echo "== variant $v"; KREG_LIB=$PWD/tools/bin/lib_$v.so python tools/debug_kvals.py 5000 2>&1 | grep -E "bad|tj%4"; done
