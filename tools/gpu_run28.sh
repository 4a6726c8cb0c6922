cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>gpurun_out/ref_final3.err | tail -1 > gpurun_out/ref_final3.json; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ref_final3.json').read())
print({k:d[k] for k in ('impl','value','steps','warmup','ms_per_step')}); print(d['cpu_baseline']['sample'][-330:])
PY
