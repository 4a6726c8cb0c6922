set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_compressors.py tests/test_gpu_robust.py tests/test_gpu_sparse24q.py -m gpu -q > gpurun_out/pytest_r2d.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/pytest_r2d.log
python tools/sparse_bench.py > gpurun_out/sparse_r2d.jsonl 2> gpurun_out/sparse_r2d.err; echo sparse rc=$?; cat gpurun_out/sparse_r2d.jsonl; tail -3 gpurun_out/sparse_r2d.err
CT_B200_BITMASK_V1=1 python tools/sparse_bench.py 2>/dev/null | grep -E "onepass|lookback" | sed 's/^/V1 /'
echo "== tile A/B: 1024 chunks (shipped) vs 2048"
python tools/jitter.py --layers 8 --launches 30 --tune 4:3,3:3,6:2 --ops quantpack,unpackdeq,fp8_q,fp8_dq,fake_w4 2>&1 | tail -4
CT_B200_LIB=$GRAFT_REPO_ROOT/compressed_tensors_b200/libct_b200_t2048.so python tools/jitter.py --layers 8 --launches 30 --tune 2:3,3:2,2:2 --ops quantpack,unpackdeq,fp8_q,fp8_dq,fake_w4 2>&1 | tail -4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-extra > gpurun_out/bench_ncu_r2.log 2>&1; echo ncu-list rc=$?
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'tile_kernel|Sparse24' -c 8 -o gpurun_out/r2_sparse2 python tools/profile_sparse.py > gpurun_out/ncu_sparse2_r2.log 2>&1; echo ncu-sparse rc=$?
ls -la gpurun_out/*.ncu-rep
