set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2e.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke_r2e.log
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2e.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/pytest_r2e.log
python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2e.json') if l.startswith('{')][-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','verified')}), json.dumps(d['roofline']))
print(json.dumps({k:(v.get('frac_of_peak'), v.get('ms')) for k,v in d['ops'].items()}))
print(json.dumps(d['e2e'])); print(json.dumps(d.get('cpu_baseline',{}).get('value')))
PY
tail -3 gpurun_out/bench_r2e.err
python tools/jitter.py --layers 8 --launches 20 --pipes 1 2>&1 | grep -E '"op"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], d['GBps_median'])
"
