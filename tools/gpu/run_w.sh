cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2w_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2w_tests.log
grep -v "^  \|^$" gpurun_out/r2w_tests.log | tail -25
python bench.py --workload sas --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2w_sas.json 2> gpurun_out/r2w_sas.err
python -c "
import json
l=json.load(open('gpurun_out/r2w_sas.json'))
print('sas value', l['value'], 'kernel ms', l['roofline'].get('kernel_ms_per_step'))"
