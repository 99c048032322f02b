# A/B of library variants on the values K build for larger molecules: bash tools/ab_kvalues_big.sh lib_a.so lib_b.so ...
cp mlatom_b200/libkreg_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp "$v" mlatom_b200/libkreg_b200.so
  echo "== $v"
  for a in "--natoms 21 --ntrain 20000" "--natoms 15 --ntrain 20000" "--natoms 12 --ntrain 30000"; do timeout 200 python tools/bench_kbuild.py $a --no-solve --reps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['workload'][:40], round(d['build_ms_max_over_ranks'],3))"; done
done
cp /tmp/lib_orig.so mlatom_b200/libkreg_b200.so
