cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'expand_rows|move_vec16' -c 4 -f -o gpurun_out/r2_expand python tools/profile_expand.py > gpurun_out/ncu_expand.log 2>&1; echo rc=$?
tail -2 gpurun_out/ncu_expand.log; ls -la gpurun_out/r2_expand.ncu-rep
