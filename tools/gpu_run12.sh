cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_observe.py tests/test_gpu_fp4.py -m gpu -q 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gather_kernel' -c 2 -o gpurun_out/r2_gather2 -f python tools/profile_sparse.py > gpurun_out/ncu_gather2.log 2>&1; echo rc=$?
ls -la gpurun_out/r2_gather2.ncu-rep
