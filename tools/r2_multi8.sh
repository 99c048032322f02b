set -x
cd $GRAFT_REPO_ROOT
N=${NGPU:-8}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n$N.json").read())
print(d["n_gpus"], d["value"], d["e2e"]["value"])
print(json.dumps(d["kbuild_sharded"])[:1500])
print(json.dumps(d["cfg4_strong"])[:300])
PY
grep -v "unbatched P2P" gpurun_out/r02_bench_n$N.err | tail -3 | cut -c1-300
