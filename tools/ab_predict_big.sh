# A/B of library variants on 21-atom prediction: bash tools/ab_predict_big.sh lib_a.so lib_b.so ...
cp mlatom_b200/libkreg_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp "$v" mlatom_b200/libkreg_b200.so
  echo "== $v"
  timeout 120 python tools/bench_predict.py --natoms 21 --ntrain 10000 --ntest 200000 --reps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('values', round(d['ms'],3), round(d['algorithmic_tflops'],2))"
  timeout 120 python tools/bench_predict.py --natoms 21 --ntrain 10000 --ntest 100000 --grads --reps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('grads ', round(d['ms'],3), round(d['algorithmic_tflops'],2))"
done
cp /tmp/lib_orig.so mlatom_b200/libkreg_b200.so
