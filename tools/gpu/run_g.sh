cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -4 gpurun_out/r2g_tests.log
( time python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench20.json ) 2> gpurun_out/r2g_bench20.err
tail -4 gpurun_out/r2g_bench20.err
python -c "
import json
l=json.load(open('gpurun_out/r2g_bench20.json'))
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(l['config']); print(l['roofline']['frac'], l['roofline']['kernel_ms_per_step_by_dt_secs']); print(l['e2e']); print(l.get('scenarios'))"
