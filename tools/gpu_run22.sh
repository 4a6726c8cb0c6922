cd $GRAFT_REPO_ROOT
for t in 64 96; do
echo "== v5 $t threads"; CT_B200_LIB=$PWD/compressed_tensors_b200/libct_b200_bm$t.so timeout 300 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_compress_onepass|bitmask_expand_lookback"
CT_B200_LIB=$PWD/compressed_tensors_b200/libct_b200_bm$t.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bitmask" 2>&1 | tail -1
done
