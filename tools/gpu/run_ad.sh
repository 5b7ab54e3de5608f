cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
B="python bench.py --steps 12 --warmup 3 --no-sas --no-sweep --no-scenarios --no-dropin --no-nitrate --no-cpu-baseline"
# launch list of the timed part of the bench: the setup (plane uploads, cell ordering) issues ~635 launches first
$B > gpurun_out/r2ad_plain.json 2> gpurun_out/r2ad_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 300 --csv --log-file gpurun_out/r2ad_launches.csv $B > gpurun_out/r2ad_launches.log 2>&1
tail -2 gpurun_out/r2ad_launches.csv | cut -c1-200
