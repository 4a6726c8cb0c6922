cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2f_n8.json 2> gpurun_out/bench_r2f_n8.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2f_n8.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','verified')})); print(json.dumps(d['e2e']))
c=d['ops']['cfg5_70b_sharded']; print(json.dumps({k:c[k] for k in c if k not in ('workload','note','timing','per_rank_dense_GB')}, indent=1))
PY
grep -E "Error|error" gpurun_out/bench_r2f_n8.err | head -5
