# A/B of library variants on small-molecule prediction: bash tools/ab_predict_small.sh lib_a.so lib_b.so ...
cp mlatom_b200/libkreg_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp "$v" mlatom_b200/libkreg_b200.so
  echo "== $v"
  for n in 3 5 6 7; do timeout 120 python tools/bench_predict.py --natoms $n --ntrain 10000 --ntest 1000000 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['natoms'], d['D'], 'ms', round(d['ms'],2), 'ps/pair', round(d['ps_per_pair'],3))"; done
  timeout 120 python tools/bench_predict.py --natoms 5 --ntrain 10000 --ntest 500000 --grads | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('grads', d['natoms'], d['D'], 'ms', round(d['ms'],2), 'ps/pair', round(d['ps_per_pair'],3))"
done
cp /tmp/lib_orig.so mlatom_b200/libkreg_b200.so
