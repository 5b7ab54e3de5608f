cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for side in 275 550 1024; do
  python bench.py --workload sas --sas-side $side --ages 1000 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2r_sas_$side.json 2> gpurun_out/r2r_sas_$side.err
  python -c "
import json
l=json.load(open('gpurun_out/r2r_sas_$side.json'))
c=$side*$side
print('side $side cells', c, 'blocks', (c+255)//256, 'value', l['value'], 'kernel ms', l['roofline']['kernel_ms_per_step'], 'ns per cell-day', 1e6*l['roofline']['kernel_ms_per_step']/c)"
done
