cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/sas_timing.py 262144 1000 3 2>&1 | tail -n 2
python tools/sas_timing.py 1048576 1500 2 2>&1 | tail -n 1
timeout 1200 python -m pytest tests/test_sas_gpu.py tests/test_nitrate_gpu.py tests/test_gamma.py -m gpu -q > gpurun_out/r2af_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2af_tests.log; grep -v "^  \|^$" gpurun_out/r2af_tests.log | tail -6
python tools/euler_timing.py 262144 1000 2 2>&1 | tail -n 2
