set -x
cd $GRAFT_REPO_ROOT
python tools/debug_kvals.py 2>&1 | tail -24
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 3 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
export KREG_EMIT_RT=1
export KREG_EMIT_VARIANT=4
for n in "9 2000" "12 1500" "16 1125" "21 857" "24 750" "32 560"; do set -- $n; echo "v4 nat=$1 templated: $(kb $1 $2)   generic: $(KREG_RT_GENERIC=1 kb $1 $2)"; done
for cfg in "4 8" "4 32" "3 8" "2 8" "2 32"; do
  set -- $cfg
  echo "v4 nat=21 TJ=$1 NJB=$2: $(KREG_RT_TJ=$1 KREG_RT_NJB=$2 kb 21 857)  generic: $(KREG_RT_GENERIC=1 KREG_RT_TJ=$1 KREG_RT_NJB=$2 kb 21 857)"
done
for cfg in "7 8" "6 8" "5 8" "4 8" "6 32"; do
  set -- $cfg
  echo "v4 nat=12 TJ=$1 NJB=$2: $(KREG_RT_TJ=$1 KREG_RT_NJB=$2 kb 12 1500)  generic: $(KREG_RT_GENERIC=1 KREG_RT_TJ=$1 KREG_RT_NJB=$2 kb 12 1500)"
done
for cfg in "9 8" "8 8" "6 8"; do
  set -- $cfg
  echo "v4 nat=9 TJ=$1 NJB=$2: $(KREG_RT_TJ=$1 KREG_RT_NJB=$2 kb 9 2000)  generic: $(KREG_RT_GENERIC=1 KREG_RT_TJ=$1 KREG_RT_NJB=$2 kb 9 2000)"
done
