cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 --no-sas --no-nitrate --no-dropin --no-sweep > gpurun_out/r2n_bench_8gpu.json ) 2> gpurun_out/r2n_bench_8gpu.err
tail -5 gpurun_out/r2n_bench_8gpu.err
python -c "
import json
l=json.load(open('gpurun_out/r2n_bench_8gpu.json'))
print({k:l[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}); print(l['roofline']['frac']); print(l['e2e']); print(l.get('scenarios'))"
nvidia-smi topo -m | head -14
