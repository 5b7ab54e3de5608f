cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
B="python bench.py --steps 12 --warmup 3 --no-sas --no-sweep --no-scenarios --no-dropin --no-nitrate --no-cpu-baseline"
# launch list of the timed part of the bench (skip the setup launches), then one full capture of the SVAT step at the bench size
$B > gpurun_out/r2ab_plain.json 2> gpurun_out/r2ab_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 440 -c 200 --csv --log-file gpurun_out/r2ab_launches.csv $B > gpurun_out/r2ab_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_svat_step -s 12 -c 1 -o gpurun_out/r2ab_svat $B > gpurun_out/r2ab_ncu.log 2>&1
# the SAS day kernel at the BASELINE size (1024^2 cells x 1500 bins)
S="python bench.py --workload sas --steps 1 --warmup 1 --no-cpu-baseline"
$S > gpurun_out/r2ab_sas_plain.json 2> gpurun_out/r2ab_sas_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_sas_day -s 1 -c 1 -o gpurun_out/r2ab_sas $S > gpurun_out/r2ab_sas_ncu.log 2>&1
ls -la gpurun_out/ | grep r2ab
