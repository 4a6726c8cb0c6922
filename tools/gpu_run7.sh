cd $GRAFT_REPO_ROOT
for K in 1 2 4; do for S in 0 20 100; do
echo "== K=$K sleep=$S"; CT_B200_BITMASK_LB_K=$K CT_B200_BITMASK_LB_SLEEP=$S python tools/sparse_bench.py 2>/dev/null | grep -E "onepass|expand_lookback" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ', d['op'], d['density'], d['us'])"
done; done
