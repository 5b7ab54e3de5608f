#!/bin/bash
# One profiling step per gpurun call (a plain run that must exit 0, then ONE ncu invocation of the same command):
#   tools/profile_step.sh <tag> launches | svat | sas | euler
set -e
cd "$(dirname "$0")/.."
TAG=$1; WHAT=$2; OUT=gpurun_out
case $WHAT in
  launches) CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --sas-side 256 --sas-steps 2"
            $CMD > $OUT/plain_$TAG.log 2>&1
            ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1 ;;
  svat)     CMD="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sas --side 2048"
            $CMD > $OUT/plain_svat_$TAG.log 2>&1
            ncu --set full --clock-control none --import-source on -k regex:k_svat_step -s 4 -c 2 -o $OUT/prof_svat_$TAG $CMD > $OUT/ncu_svat_$TAG.log 2>&1 ;;
  sas)      CMD="python tools/sas_timing.py 131072 1500 2"
            $CMD > $OUT/plain_sas_$TAG.log 2>&1
            ncu --set full --clock-control none --import-source on -k regex:k_sas_day -s 1 -c 1 -o $OUT/prof_sas_$TAG $CMD > $OUT/ncu_sas_$TAG.log 2>&1
            tail -2 $OUT/plain_sas_$TAG.log ;;
  euler)    CMD="python tools/euler_timing.py 131072 1000 2"
            $CMD > $OUT/plain_euler_$TAG.log 2>&1
            ncu --set full --clock-control none --import-source on -k regex:k_sas_day_euler -s 1 -c 1 -o $OUT/prof_euler_$TAG $CMD > $OUT/ncu_euler_$TAG.log 2>&1
            tail -1 $OUT/plain_euler_$TAG.log ;;
esac
echo "profile step $WHAT done"
