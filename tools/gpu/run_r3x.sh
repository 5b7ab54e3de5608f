cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# the shipped build again (inline cold paths): SAS / nitrate parity tests and the SAS bench line on the box that measured r3w
timeout 300 python -m pytest tests/test_sas_gpu.py tests/test_nitrate_gpu.py -m gpu -q > gpurun_out/r3x_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r3x_tests.log; tail -3 gpurun_out/r3x_tests.log
python bench.py --workload sas --no-cpu-baseline --no-sweep --no-nitrate --no-dropin --no-scenarios > gpurun_out/r3x_bench_sas.json 2> gpurun_out/r3x_bench_sas.err; echo "bench exit $?"
python -c "
import json
l=json.load(open('gpurun_out/r3x_bench_sas.json'))
print('value', l['value'], l['unit'], 'ms', l['ms_per_step'], 'frac', l['roofline']['frac'], l['clocks'])"
