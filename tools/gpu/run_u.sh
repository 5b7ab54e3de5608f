cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tools/gpu/diag_families2.py > gpurun_out/r2u_diag.log 2>&1
tail -80 gpurun_out/r2u_diag.log
