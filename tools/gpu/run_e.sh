cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_svat_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -15 gpurun_out/r2e_tests.log
python tools/gpu/probe_steps.py 4096 120 gpurun_out/r2e_probe_order.json 0
ROGER_B200_CELL_ORDER=0 python tools/gpu/probe_steps.py 4096 120 gpurun_out/r2e_probe_noorder.json 0
