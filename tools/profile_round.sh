#!/bin/bash
# The profiling recipe of B200_PROFILING.md for this repo (run under gpurun on ONE GPU):
#   1. plain bench run (must exit 0), 2. launch list of the same command, 3. one `--set full` capture of each hot kernel.
# usage: tools/profile_round.sh <tag>
set -e
cd "$(dirname "$0")/.."
TAG=${1:-rX}
OUT=gpurun_out
BENCH="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --sas-side 256 --sas-steps 2"
$BENCH > $OUT/plain_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
SV="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sas --side 2048"
$SV > $OUT/plain_svat_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_svat_step -s 4 -c 2 -o $OUT/prof_svat_$TAG $SV > $OUT/ncu_svat_$TAG.log 2>&1
SA="python tools/sas_timing.py 131072 1500 2"
$SA > $OUT/plain_sas_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sas_day -s 1 -c 1 -o $OUT/prof_sas_$TAG $SA > $OUT/ncu_sas_$TAG.log 2>&1
tail -2 $OUT/plain_sas_$TAG.log
