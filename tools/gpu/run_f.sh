cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python bench.py --side 512 --scenario-side 256 --steps 6 --warmup 3 --sas-side 64 --ages 200 --ref-side 64 --ref-cells 256 > gpurun_out/r2f_small.json ) 2> gpurun_out/r2f_small.err
tail -5 gpurun_out/r2f_small.err
( time python bench.py > gpurun_out/r2f_bench.json ) 2> gpurun_out/r2f_bench.err
tail -4 gpurun_out/r2f_bench.err
python -c "
import json
l=json.load(open('gpurun_out/r2f_bench.json'))
print({k:l[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(l['roofline']); print(l['e2e']); print(l.get('sweep')); print(l.get('strong')); print(l.get('scenarios')); print(l['sas']['value'], l['sas']['roofline']['frac']); print(l.get('cpu_baseline'))"
