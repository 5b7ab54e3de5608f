#!/bin/bash
# builds variants of the SAS kernel (unroll : min blocks : block size : prefetch stages) into temp libs and times them
# usage: VARIANTS="2:2:256:4 1:4:128:8" tools/sas_variants.sh
set -e
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-2:2:256:4 2:2:256:8 2:4:128:4 1:2:256:8 4:2:256:2 2:3:256:4}; do
  IFS=: read U B T D <<< "$v"
  export ROGER_B200_LIB=/tmp/libroger_u${U}_b${B}_t${T}_d${D}.so
  ROGER_NVCC_EXTRA="-DROGER_SAS_UNROLL=$U -DROGER_SAS_MIN_BLOCKS=$B -DROGER_SAS_BLOCK=$T " python -c "from roger_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A2 "k_sas_dayILb0" | grep -E "Used|spill" | tr '\n' ' '
  echo; echo "== unroll $U min_blocks $B block $T stages $D"
  python tools/sas_timing.py ${CELLS:-262144} ${AGES:-1000} 2 | tail -1
done
