cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# SAS day kernel: launch-shape and cold-path variants against the shipped build (tools/sas_timing.py, 262144 cells x 1000 bins), each with its parity check
for v in default t192b3 t160b4 t320b2 cold; do
  if [ $v = default ]; then unset ROGER_B200_LIB; else export ROGER_B200_LIB=$PWD/roger_b200/lib/variants/$v.so; [ -f $ROGER_B200_LIB ] || continue; fi
  echo "== $v" >> gpurun_out/r3v_sas.log
  timeout 120 python tools/sas_timing.py 262144 1000 4 >> gpurun_out/r3v_sas.log 2>&1
  timeout 120 python -m pytest tests/test_sas_gpu.py -m gpu -q -x -k "test_cuda_equals_oracle_bit_for_bit_on_seeded_inputs or test_edge_shapes" 2>&1 | tail -1 >> gpurun_out/r3v_sas.log
done
cat gpurun_out/r3v_sas.log
