#!/bin/bash
# times occupancy variants of the single-step SVAT kernel: VARIANTS="minblocks:block ..." tools/svat_step_variants.sh
set -e
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-4:128 5:128 6:128 8:64 10:64 3:192 2:256}; do
  IFS=: read B T <<< "$v"
  export ROGER_B200_LIB=/tmp/libroger_kb${B}_t${T}.so
  ROGER_NVCC_EXTRA="-DROGER_MIN_BLOCKS=$B -DROGER_BLOCK=$T" python -c "from roger_b200 import build; build.build(force=True, verbose=True)" 2>&1 | grep -A3 "k_svat_stepILb1ELb0" | grep -E "Used|spill" | tr '\n' ' '
  echo; echo "== step kernel min_blocks $B block $T"
  python bench.py --no-sas --no-cpu-baseline --side ${SIDE:-4096} --steps 24 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.3e ms/step %.3f frac %.3f mix %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['step_mix_dt_secs']))"
done
