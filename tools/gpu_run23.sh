cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_final_n2.json 2> gpurun_out/bench_final_n2.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_final_n2.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','verified')}), d['e2e']['value'])
c=d['ops']['cfg5_70b_sharded']; print(json.dumps({k:c[k] for k in c if k not in ('workload','note','timing','per_rank_dense_GB')}, indent=1))
PY
grep -E "Error|error" gpurun_out/bench_final_n2.err | head -5
python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -2
