cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -m pytest tests/test_nitrate_gpu.py tests/test_sas_gpu.py -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
tail -15 gpurun_out/r2l_tests.log
# launch list of the timed part of the bench (skip the setup launches: plane uploads, cell ordering)
python bench.py --steps 12 --warmup 3 --no-sas --no-sweep --no-scenarios --no-dropin --no-cpu-baseline > gpurun_out/r2l_plain.json 2> gpurun_out/r2l_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 440 -c 200 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 12 --warmup 3 --no-sas --no-sweep --no-scenarios --no-dropin --no-cpu-baseline > gpurun_out/r2l_launches.log 2>&1
tail -3 gpurun_out/r2l_launches.csv | cut -c1-250
