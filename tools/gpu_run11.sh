# 8-GPU box: final N = 8 bench line (all-gather recouple + mirror plan) and the socket A/B that names the e2e limiter:
# 4 ranks on ONE socket's GPUs (0-3) against 4 ranks spread 2 + 2 over both sockets (0,1,4,5), e2e leg only.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2f_n8.json 2> gpurun_out/bench_r2f_n8.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2f_n8.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','verified')})); print(json.dumps(d['e2e']))
c=d['ops']['cfg5_70b_sharded']; print(json.dumps({k:c[k] for k in c if k not in ('workload','note','timing','per_rank_dense_GB')}, indent=1))
PY
tail -3 gpurun_out/bench_r2f_n8.err
for vis in 0,1,2,3 0,1,4,5; do
  CUDA_VISIBLE_DEVICES=$vis python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 4 --steps 5 --warmup 3 --no-extra --no-cfg5 > gpurun_out/bench_r2f_n4_$vis.json 2> gpurun_out/bench_r2f_n4_$vis.err; echo rc=$?
  python - $vis <<'PY'
import json, sys
d=json.loads([l for l in open('gpurun_out/bench_r2f_n4_%s.json' % sys.argv[1]) if l.startswith('{')][-1])
print("N=4 on GPUs", sys.argv[1], "e2e:", json.dumps(d['e2e']), "numa:", json.dumps(d.get('numa')))
PY
done
