cd $GRAFT_REPO_ROOT
echo "== v5 default (256 threads)"; timeout 300 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_compress_onepass"
echo "== v5 128 threads"; CT_B200_LIB=$PWD/compressed_tensors_b200/libct_b200_bm128.so timeout 300 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_compress_onepass|bitmask_expand_lookback"
CT_B200_LIB=$PWD/compressed_tensors_b200/libct_b200_bm128.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bitmask" 2>&1 | tail -1
