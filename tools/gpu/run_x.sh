cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_svat_gpu.py tests/test_abi.py -m gpu -q > gpurun_out/r2x_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2x_tests.log
grep -v "^  \|^$" gpurun_out/r2x_tests.log | tail -15
python bench.py --steps 20 --warmup 3 --no-sas --no-nitrate --no-dropin --no-sweep --no-scenarios --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python -c "
import json
l=json.load(open('gpurun_out/r2x_bench.json'))
print('value', l['value'], 'ms', l['ms_per_step'], 'frac', l['roofline']['frac'], 'e2e', l['e2e']['value'])"
