cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel' -c 2 -o gpurun_out/r2_gather python tools/profile_sparse.py > gpurun_out/ncu_gather.log 2>&1; echo rc=$?
ls -la gpurun_out/r2_gather.ncu-rep
