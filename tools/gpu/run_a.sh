cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
for v in default nopf pf148 pf296; do
  if [ $v = default ]; then unset ROGER_B200_LIB; else export ROGER_B200_LIB=$PWD/roger_b200/lib/variants/$v.so; fi
  python bench.py --no-sas --no-cpu-baseline --steps 24 --warmup 4 > gpurun_out/r2a_bench_$v.json 2> gpurun_out/r2a_bench_$v.err
done
unset ROGER_B200_LIB
python bench.py --no-sas --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/r2a_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_svat_step -s 5 -c 2 -o gpurun_out/r2a_svat python bench.py --no-sas --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/r2a_ncu.log 2>&1
tail -3 gpurun_out/r2a_tests.log; for v in default nopf pf148 pf296; do python -c "
import json,sys
l=json.load(open('gpurun_out/r2a_bench_$v.json'))
print('$v', l['value'], l['ms_per_step'], l['roofline']['frac'], l['roofline']['kernel_ms_per_step'], l['e2e']['value'], l['clocks'])"; done
