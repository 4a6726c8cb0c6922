cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compressors.py tests/test_gpu_large.py -m gpu -q -x -k "bitmask or Bitmask or sparse" 2>&1 | tail -3
timeout 300 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_compress_onepass"
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "onepass_lookback and not 14336" 2>&1 | tail -3
