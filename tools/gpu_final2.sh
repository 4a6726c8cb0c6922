# last check of the committed build: smoke, the whole -m gpu suite, a short bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final2.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_final2.log
timeout 200 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_final2.json 2>gpurun_out/bench_final2.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_final2.json') if l.startswith('{')][-1])
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','verified')}), d['roofline']['frac'], d['roofline']['traffic_source'][:120])
print(json.dumps({k:(v.get('frac_of_peak'), v.get('ms')) for k,v in d['ops'].items() if isinstance(v, dict)}))
PY
