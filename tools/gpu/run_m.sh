cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -4 gpurun_out/r2m_tests.log
( time python bench.py --steps 20 --warmup 3 > gpurun_out/r2m_bench20.json ) 2> gpurun_out/r2m_bench20.err
tail -4 gpurun_out/r2m_bench20.err
python -c "
import json
l=json.load(open('gpurun_out/r2m_bench20.json'))
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}); print(l['roofline']['frac'], l['roofline']['traffic']); print(l['e2e']['value']); print(l.get('dropin')); print(l.get('nitrate')); print(l['sas']['value'], l['sas']['roofline']['frac'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
