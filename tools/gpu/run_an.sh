cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in default b128x5 b128x6; do
  if [ $v = default ]; then unset ROGER_B200_LIB; else export ROGER_B200_LIB=$PWD/roger_b200/lib/variants/$v.so; fi
  echo "== $v"
  python tools/sas_timing.py 262144 1000 3 2>&1 | tail -n 1
  python tools/sas_timing.py 1048576 1500 2 2>&1 | tail -n 1
done
