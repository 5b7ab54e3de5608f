cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default sasB sasC sasD sasE; do
  if [ $v = default ]; then unset ROGER_B200_LIB; else export ROGER_B200_LIB=$PWD/roger_b200/lib/variants/$v.so; fi
  python bench.py --workload sas --sas-side 512 --ages 1000 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2j_sas_$v.json 2> gpurun_out/r2j_sas_$v.err
  python -c "
import json
l=json.load(open('gpurun_out/r2j_sas_$v.json'))
print('$v', l['value'], l['ms_per_step'], l['roofline']['kernel_ms_per_step'])"
done
