cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/sas_timing.py 262144 1000 3 2>&1 | tail -n 1
python tools/sas_timing.py 1048576 1500 2 2>&1 | tail -n 1
python bench.py --workload sas --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2am_sas.json 2> gpurun_out/r2am_sas.err
python -c "
import json
l=json.load(open('gpurun_out/r2am_sas.json'))
print('bench sas value', l['value'], 'kernel ms', l['roofline']['kernel_ms_per_step'])"
timeout 1500 python -m pytest tests/test_sas_gpu.py tests/test_nitrate_gpu.py tests/test_gamma.py tests/test_svat_gpu.py -m gpu -q > gpurun_out/r2am_tests.log 2>&1; echo "exit $?" >> gpurun_out/r2am_tests.log; grep -v "^  \|^$" gpurun_out/r2am_tests.log | tail -5
