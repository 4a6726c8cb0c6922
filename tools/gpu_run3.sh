set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2c.log 2>&1; echo pytest rc=$?; tail -12 gpurun_out/pytest_r2c.log
python tools/sparse_bench.py > gpurun_out/sparse_r2c.jsonl 2> gpurun_out/sparse_r2c.err; echo sparse rc=$?; cat gpurun_out/sparse_r2c.jsonl; tail -3 gpurun_out/sparse_r2c.err
python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2c.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','verified')}), json.dumps(d['roofline']))
print(json.dumps({k:(v.get('frac_of_peak'), v.get('ms')) for k,v in d['ops'].items()}))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/bench_ncu_r2.log 2>&1; echo ncu-list rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stream_tma_kernel' -c 2 -o gpurun_out/r2_quantpack python tools/profile_ops.py --layers 4 --reps 1 --ops quantpack > gpurun_out/ncu_quantpack_r2.log 2>&1; echo ncu-qp rc=$?
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ring_kernel|lookback_kernel|Sparse24|minmax_qparams|QuantizeOp' --launch-skip 0 -c 30 -o gpurun_out/r2_sparse python tools/profile_sparse.py > gpurun_out/ncu_sparse_r2.log 2>&1; echo ncu-sparse rc=$?
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_sweep.py > gpurun_out/memcheck_r2.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/memcheck_r2.log
ls -la gpurun_out/*.ncu-rep
