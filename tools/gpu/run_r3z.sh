cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# full GPU test suite, smoke, default bench of the final build
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r3z_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r3z_tests.log
grep -v "^  \|^$" gpurun_out/r3z_tests.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3z_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r3z_smoke.log; tail -4 gpurun_out/r3z_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3z_bench_reference.json 2> gpurun_out/r3z_bench_reference.err; echo "reference arm exit $?"
python -c "
import json
l=json.load(open('gpurun_out/r3z_bench.json'))
print('value', l['value'], 'frac', l['roofline']['frac'], 'e2e', l['e2e']['value'], 'sas', l.get('sas',{}).get('value'), 'nitrate', l.get('nitrate',{}).get('value'))
r=json.load(open('gpurun_out/r3z_bench_reference.json'))
print('reference arm', r.get('value'), r.get('unit'))"
