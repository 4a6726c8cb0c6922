cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_compressors.py tests/test_gpu_large.py -m gpu -q -k "bitmask or Bitmask or sparse" 2>&1 | tail -3
python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask"
