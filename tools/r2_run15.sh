cd $GRAFT_REPO_ROOT
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
for n in "12 1500" "16 1125" "21 857" "24 750" "32 560" "48 375"; do set -- $n
  echo "nat=$1 v4: $(KREG_EMIT_VARIANT=4 kb $1 $2)  v3 inorder: $(KREG_EMIT_VARIANT=3 KREG_RT_INORDER=1 kb $1 $2)  v3 holes: $(KREG_EMIT_VARIANT=3 KREG_RT_INORDER=0 kb $1 $2)"
done
for cfg in "4 8" "4 4" "3 8" "2 8"; do set -- $cfg; echo "nat=21 v3 inorder TJ=$1 NIB=$2: $(KREG_EMIT_VARIANT=3 KREG_RT_TJ=$1 KREG_RT_NIB=$2 kb 21 857)"; done
for cfg in "7 8" "5 8" "4 8"; do set -- $cfg; echo "nat=12 v3 inorder TJ=$1 NIB=$2: $(KREG_EMIT_VARIANT=3 KREG_RT_TJ=$1 KREG_RT_NIB=$2 kb 12 1500)"; done
