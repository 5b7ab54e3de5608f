cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default nohints; do
  if [ $v = default ]; then unset ROGER_B200_LIB; else export ROGER_B200_LIB=$PWD/roger_b200/lib/variants/$v.so; fi
  echo "== $v"
  python bench.py --workload sas --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2ak_$v.json 2> gpurun_out/r2ak_$v.err
  python -c "
import json
l=json.load(open('gpurun_out/r2ak_$v.json'))
print('value', l['value'], 'kernel ms', l['roofline']['kernel_ms_per_step'], 'clocks', l.get('clocks'))"
  python tools/sas_timing.py 1048576 1500 2 2>&1 | tail -n 1
done
