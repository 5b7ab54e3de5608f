cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# after the cold-path change of the SAS sub-step pass: every SAS / nitrate GPU parity test, the full-size SAS replica test, timing, SAS bench line
timeout 600 python -m pytest tests/test_sas_gpu.py tests/test_nitrate_gpu.py tests/test_full_size_gpu.py -m gpu -q -k "not oned" > gpurun_out/r3w_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r3w_tests.log; tail -3 gpurun_out/r3w_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3w_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r3w_smoke.log; tail -4 gpurun_out/r3w_smoke.log
python tools/sas_timing.py 262144 1000 3 > gpurun_out/r3w_sas_timing.log 2>&1; cat gpurun_out/r3w_sas_timing.log
python bench.py --workload sas --no-cpu-baseline --no-sweep --no-nitrate --no-dropin --no-scenarios > gpurun_out/r3w_bench_sas.json 2> gpurun_out/r3w_bench_sas.err; echo "bench exit $?"
python -c "
import json
l=json.load(open('gpurun_out/r3w_bench_sas.json'))
print('value', l['value'], l['unit'], 'ms', l['ms_per_step'], 'frac', l['roofline']['frac'])"
