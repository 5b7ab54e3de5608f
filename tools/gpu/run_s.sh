cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gamma.py tests/test_sas_gpu.py tests/test_nitrate_gpu.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2s_tests.log
tail -15 gpurun_out/r2s_tests.log
