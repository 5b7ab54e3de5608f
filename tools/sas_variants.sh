#!/bin/bash
# builds variants of the SAS kernel (unroll / occupancy / block size) into temp libs and times them
set -e
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-"2 4 128" "2 5 128" "2 6 128" "1 6 128" "1 8 128" "2 8 64" "2 2 256" "1 4 256"}; do
  set -- $v
  export ROGER_B200_LIB=/tmp/libroger_u$1_b$2_t$3.so
  ROGER_NVCC_EXTRA="-DROGER_SAS_UNROLL=$1 -DROGER_SAS_MIN_BLOCKS=$2 -DROGER_SAS_BLOCK=$3" python -c "from roger_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A2 "k_sas_day" | grep -E "Used|spill" | tr '\n' ' '
  echo; echo "== unroll $1 min_blocks $2 block $3"
  python tools/sas_timing.py ${CELLS:-262144} ${AGES:-1000} 2 | tail -1
done
