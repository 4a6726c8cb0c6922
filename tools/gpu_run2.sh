set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_distributed.py tests/test_gpu_observe.py tests/test_gpu_sparse24q.py tests/test_gpu_compressors.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_r2b.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/pytest_r2b.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2b_n2.json 2> gpurun_out/bench_r2b_n2.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2b_n2.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','verified')}))
print(json.dumps(d['ops']['cfg5_70b_sharded'], indent=1))
PY
tail -5 gpurun_out/bench_r2b_n2.err
