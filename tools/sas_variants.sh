#!/bin/bash
# builds variants of the SAS kernel (unroll / occupancy / block size) into temp libs and times them
# usage: VARIANTS="2:4:128 1:6:128" tools/sas_variants.sh
set -e
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-2:4:128 2:3:128 1:4:128 1:6:128 2:6:128 4:3:128 2:2:256}; do
  IFS=: read U B T <<< "$v"
  export ROGER_B200_LIB=/tmp/libroger_u${U}_b${B}_t${T}.so
  ROGER_NVCC_EXTRA="-DROGER_SAS_UNROLL=$U -DROGER_SAS_MIN_BLOCKS=$B -DROGER_SAS_BLOCK=$T" python -c "from roger_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A2 "k_sas_day" | grep -E "Used|spill" | tr '\n' ' '
  echo; echo "== unroll $U min_blocks $B block $T"
  python tools/sas_timing.py ${CELLS:-262144} ${AGES:-1000} 2 | tail -1
done
