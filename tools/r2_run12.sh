cd $GRAFT_REPO_ROOT
python tools/debug_kvals.py 5000 2>&1 | grep -vE "^\+" | cut -c1-900
