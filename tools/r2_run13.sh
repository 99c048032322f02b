cd $GRAFT_REPO_ROOT
python tools/debug_kvals.py 5000 20000 2>&1 | grep -E "bad|tj%4"
for i in 1 2; do python tools/bench_kbuild.py --natoms 9 --ntrain 50000 --no-solve --reps 5 2>&1 | grep -o '"build_ms_max_over_ranks": [0-9.]*'; done
python tools/bench_kbuild.py --natoms 21 --ntrain 20000 --no-solve --reps 5 2>&1 | grep -o '"build_ms_max_over_ranks": [0-9.]*'
