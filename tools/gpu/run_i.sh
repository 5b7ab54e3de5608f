cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_svat_gpu.py -m gpu -x -q -k "lazy or cell_order or drop_in" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -4 gpurun_out/r2i_tests.log
( time python bench.py --steps 20 --warmup 3 --no-sas > gpurun_out/r2i_bench20.json ) 2> gpurun_out/r2i_bench20.err
tail -4 gpurun_out/r2i_bench20.err
python -c "
import json
l=json.load(open('gpurun_out/r2i_bench20.json'))
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}); print(l['config']['window'], l['config']['step_mix_dt_secs']); print(l['roofline']['frac']); print(l['e2e']); print(l.get('dropin')); print(l.get('strong'))"
