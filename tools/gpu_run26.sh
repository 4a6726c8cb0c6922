cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_sparse24q.py tests/test_gpu_compressors.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/sparse_bench.py 2>/dev/null | grep -E "sparse24"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print({k:(v.get('frac_of_peak'), v.get('ms')) for k,v in d['ops'].items() if k.startswith('cfg4')})"
