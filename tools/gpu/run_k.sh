cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# launch list of the default bench command (shares, not absolutes), then one full capture of the SVAT step at the bench size
python bench.py --steps 6 --warmup 3 --no-sas --no-sweep --no-scenarios --no-dropin --no-cpu-baseline > gpurun_out/r2k_plain.json 2> gpurun_out/r2k_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 6 --warmup 3 --no-sas --no-sweep --no-scenarios --no-dropin --no-cpu-baseline > gpurun_out/r2k_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_svat_step -s 12 -c 1 -o gpurun_out/r2k_svat python bench.py --steps 6 --warmup 3 --no-sas --no-sweep --no-scenarios --no-dropin --no-cpu-baseline > gpurun_out/r2k_ncu.log 2>&1
# the SAS day kernel at the BASELINE size (1024^2 cells x 1500 bins)
python bench.py --workload sas --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_sas_plain.json 2> gpurun_out/r2k_sas_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_sas_day -s 1 -c 1 -o gpurun_out/r2k_sas python bench.py --workload sas --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2k_sas_ncu.log 2>&1
ls -la gpurun_out/
