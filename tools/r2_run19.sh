cd $GRAFT_REPO_ROOT
kb() { python tools/bench_kbuild.py --natoms $1 --ntrain $2 --grads --no-solve --reps 4 2>&1 | grep -o '"build_gb_per_s": [0-9.]*'; }
for n in "6 3000" "8 2250" "9 2000" "10 1800"; do set -- $n; echo "nat=$1 template: $(kb $1 $2)   rt emit: $(KREG_EMIT_RT=1 kb $1 $2)  rt emit TJ-1: $(KREG_EMIT_RT=1 KREG_RT_TJ=$((256/(3*$1)-1)) kb $1 $2)"; done
for n in 6 9 12 16 21 24; do timeout 300 python tools/bench_predict.py --natoms $n --ntrain 10000 --ntest 200000 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['natoms'], round(d['ms'],2), round(d['algorithmic_tflops'],1))"; done
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"; done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
