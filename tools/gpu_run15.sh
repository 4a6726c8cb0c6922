cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_compressors.py tests/test_gpu_large.py -m gpu -q -x -k "bitmask or Bitmask or sparse" 2>&1 | tail -5
echo "== pipelined"; python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_expand_row_offsets"
echo "== stages 2"; CT_B200_BITMASK_ROWS_STAGES=2 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_expand_row_offsets"
echo "== stages 4"; CT_B200_BITMASK_ROWS_STAGES=4 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_expand_row_offsets"
echo "== per-row v1"; CT_B200_BITMASK_ROWS_V1=1 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_expand_row_offsets"
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "expand_rows_pipeline" 2>&1 | tail -4
