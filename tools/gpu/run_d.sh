cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_svat_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -4 gpurun_out/r2d_tests.log
python tools/gpu/probe_steps.py 4096 120 gpurun_out/r2d_probe_default.json 0
ROGER_B200_LIB=$PWD/roger_b200/lib/variants/mb3.so python tools/gpu/probe_steps.py 4096 120 gpurun_out/r2d_probe_mb3.json 0
