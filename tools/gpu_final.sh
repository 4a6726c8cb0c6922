# final 1-GPU validation + profiling of the shipping build (round 2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/smi_final.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke_final.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','verified','gpu_launches')}), json.dumps(d['roofline']))
print(json.dumps(d['e2e'])); print(json.dumps(d['cpu_baseline'])); print(json.dumps(d.get('clocks')))
print(json.dumps({k:(v.get('frac_of_peak'), v.get('ms')) for k,v in d['ops'].items() if isinstance(v, dict)}))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo ref rc=$?; tail -1 gpurun_out/bench_final_ref.json | cut -c1-600
python tools/sparse_bench.py > gpurun_out/sparse_final.jsonl 2>/dev/null; cat gpurun_out/sparse_final.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/bench_ncu_final.log 2>&1; echo ncu-list rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stream_tma_kernel' -c 2 -f -o gpurun_out/final_quantpack python tools/profile_ops.py --layers 4 --reps 1 --ops quantpack > gpurun_out/ncu_quantpack_final.log 2>&1; echo ncu-qp rc=$?
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'late_kernel|expand_rows|expand_tile|Sparse24|minmax_qparams' -c 16 -f -o gpurun_out/final_sparse python tools/profile_sparse.py > gpurun_out/ncu_sparse_final.log 2>&1; echo ncu-sparse rc=$?
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_sweep.py > gpurun_out/memcheck_final.log 2>&1; echo memcheck rc=$?; tail -3 gpurun_out/memcheck_final.log
ls -la gpurun_out/final_*.ncu-rep
