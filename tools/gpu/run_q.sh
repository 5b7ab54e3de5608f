cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default split2 split3 split4; do
  if [ $v = default ]; then unset ROGER_B200_LIB; else export ROGER_B200_LIB=$PWD/roger_b200/lib/variants/$v.so; fi
  python bench.py --workload sas --sas-side 512 --ages 1000 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2q_sas_$v.json 2> gpurun_out/r2q_sas_$v.err
  python -c "
import json
l=json.load(open('gpurun_out/r2q_sas_$v.json'))
print('$v', l['value'], l['ms_per_step'], l['roofline']['kernel_ms_per_step'])"
done
ROGER_B200_LIB=$PWD/roger_b200/lib/variants/split4.so python -m pytest tests/test_sas_gpu.py -m gpu -x -q -k "bit_for_bit_on_fixture or seeded or edge" 2>&1 | tail -3
