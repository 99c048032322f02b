set -e
for args in "--natoms 9 --ntrain 10000 --ntest 1000000" "--natoms 9 --ntrain 2000 --ntest 1000000 --grads" "--natoms 9 --ntrain 10000 --ntest 500000 --grads" "--natoms 6 --ntrain 10000 --ntest 1000000" "--natoms 12 --ntrain 10000 --ntest 500000" "--natoms 12 --ntrain 10000 --ntest 200000 --grads" "--natoms 13 --ntrain 10000 --ntest 300000" "--natoms 15 --ntrain 10000 --ntest 300000" "--natoms 18 --ntrain 10000 --ntest 200000" "--natoms 21 --ntrain 10000 --ntest 200000" "--natoms 21 --ntrain 10000 --ntest 100000 --grads" "--natoms 21 --ntrain 10000 --ntest 200000 --energy-only"; do
  timeout 120 python tools/bench_predict.py $args
done
