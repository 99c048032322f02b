# A/B of library variants on the values K build: NTRAIN=<n> bash tools/ab_kvalues.sh lib_a.so lib_b.so ...   (default 50000 = cfg3)
cp mlatom_b200/libkreg_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp "$v" mlatom_b200/libkreg_b200.so
  echo "== $v"
  for i in 1 2; do timeout 200 python tools/bench_kbuild.py --ntrain ${NTRAIN:-50000} --no-solve --reps 8 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['build_ms_max_over_ranks'], d['build_gb_per_s'])"; done
done
cp /tmp/lib_orig.so mlatom_b200/libkreg_b200.so
