set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2f.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/pytest_r2f.log
python tools/sparse_bench.py > gpurun_out/sparse_r2f.jsonl 2> gpurun_out/sparse_r2f.err; echo sparse rc=$?; cat gpurun_out/sparse_r2f.jsonl; tail -3 gpurun_out/sparse_r2f.err
CT_B200_BITMASK_V3=1 python tools/sparse_bench.py 2>/dev/null | grep -E "onepass" | sed 's/^/V3 /'
