cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_compressors.py tests/test_gpu_large.py -m gpu -q -k "bitmask or Bitmask or sparse" 2>&1 | tail -3
echo "== v5 (default)"; python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_compress_onepass"
echo "== v4"; CT_B200_BITMASK_V4=1 python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_compress_onepass"
for t in 192 224; do
  echo "== v5 threads $t"; CT_B200_LIB=$PWD/compressed_tensors_b200/libct_b200_bm$t.so python tools/sparse_bench.py 2>/dev/null | grep -E "bitmask_compress_onepass"
  CT_B200_LIB=$PWD/compressed_tensors_b200/libct_b200_bm$t.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bitmask" 2>&1 | tail -1
done
