#!/bin/bash
# times variants of the day-blocked SVAT kernel: VARIANTS="minblocks:block ..." tools/svat_variants.sh
set -e
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-2:128 1:128 3:128 4:128 1:256 2:64}; do
  IFS=: read B T <<< "$v"
  export ROGER_B200_LIB=/tmp/libroger_sb${B}_t${T}.so
  ROGER_NVCC_EXTRA="-DROGER_STEPS_MIN_BLOCKS=$B -DROGER_STEPS_BLOCK=$T" python -c "from roger_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A3 "k_svat_stepsILb1" | grep -E "Used|spill" | tr '\n' ' '
  echo; echo "== steps kernel min_blocks $B block $T"
  python bench.py --no-sas --no-cpu-baseline --side ${SIDE:-4096} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.3f frac %.3f launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['gpu_launches']))"
done
