cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for out in 1 0; do
ROGER_BENCH_E2E_OUTPUT=$out python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$out bench.py --gpus 2 --steps 20 --warmup 3 --no-sas --no-nitrate --no-dropin --no-sweep --no-scenarios > gpurun_out/r2p_$out.json 2> gpurun_out/r2p_$out.err
python -c "
import json
l=json.load(open('gpurun_out/r2p_$out.json'))
print('output=$out', l['value'], l['e2e']['value'], l['e2e']['value']/l['value'])"
done
