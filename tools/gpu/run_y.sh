cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_svat_gpu.py tests/test_restart.py tests/test_abi.py -m gpu -q > gpurun_out/r2y_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2y_tests.log
grep -v "^  \|^$" gpurun_out/r2y_tests.log | tail -15
