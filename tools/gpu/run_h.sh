cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2h_bench_2gpu.json ) 2> gpurun_out/r2h_bench_2gpu.err
tail -6 gpurun_out/r2h_bench_2gpu.err
python -c "
import json
l=json.load(open('gpurun_out/r2h_bench_2gpu.json'))
print({k:l[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}); print(l['roofline']['frac']); print(l['e2e']); print(l.get('strong')); print([ (r['cells_per_gpu'], round(r['value']/1e9,2)) for r in l['sweep']]); print(l.get('scenarios')); print(l['sas']['value'], l['sas']['e2e'])"
