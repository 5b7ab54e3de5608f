cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tools/gpu/diag_families2.py > gpurun_out/r2v_diag.log 2>&1
grep -c "equal" gpurun_out/r2v_diag.log
timeout 1500 python -m pytest tests/test_gamma.py tests/test_sas_gpu.py tests/test_nitrate_gpu.py -m gpu -q > gpurun_out/r2v_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2v_tests.log
grep -v "^  \|^$" gpurun_out/r2v_tests.log | tail -30
