#!/bin/bash
# FUSE_LAST : LOADGROUP variants of the SAS kernel
set -e
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-0:8 0:4 1:4 1:8 0:2}; do
  IFS=: read F G <<< "$v"
  export ROGER_B200_LIB=/tmp/libroger_f${F}_g${G}.so
  ROGER_NVCC_EXTRA="-DROGER_SAS_FUSE_LAST=$F -DROGER_SAS_LOADGROUP=$G" python -c "from roger_b200 import build; build.build(force=True)"
  echo "== fuse_last $F loadgroup $G"
  python tools/sas_timing.py 262144 1000 2 | tail -1
  python tools/sas_timing.py 524288 1500 2 | tail -1
done
