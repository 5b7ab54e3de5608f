cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in va vb; do
  export ROGER_B200_LIB=$PWD/roger_b200/lib/variants/$v.so
  echo "== $v"
  python tools/euler_timing.py 262144 1000 3 2>&1 | tail -n 1
done
