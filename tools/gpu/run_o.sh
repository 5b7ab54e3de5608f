cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 --no-sas --no-nitrate --no-dropin --no-sweep --no-scenarios > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench_2gpu.err
python -c "
import json
l=json.load(open('gpurun_out/r2o_bench_2gpu.json'))
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}); print(l['e2e']['value'], l['e2e']['value']/l['value'])"
