cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2al_bench_2gpu.json 2> gpurun_out/r2al_bench_2gpu.err
echo "exit $?"
python -c "
import json
l=json.load(open('gpurun_out/r2al_bench_2gpu.json'))
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', l['e2e']['value'], l['e2e']['value']/l['value']); print('sas', l.get('sas',{}).get('value')); print('scen', {k:v.get('value') for k,v in l.items() if isinstance(v,dict) and k.startswith('scen')})"
