set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" > gpurun_out/lscpu_n8.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n8.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','verified')})); print(json.dumps(d['e2e']))
print(json.dumps(d['ops']['cfg5_70b_sharded'], indent=1))
PY
tail -3 gpurun_out/bench_r2_n8.err
CT_BENCH_NUMA=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra --no-cfg5 > gpurun_out/bench_r2_n8_nonuma.json 2> gpurun_out/bench_r2_n8_nonuma.err; echo bench-nonuma rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n8_nonuma.json') if l.startswith('{')][-1])
print("NUMA off:", json.dumps(d['e2e']))
PY
cat gpurun_out/lscpu_n8.txt; head -14 gpurun_out/topo_n8.txt
