cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/gpu/probe_steps.py 4096 42 gpurun_out/r2c_probe.json 0 > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_svat_step -s 20 -c 1 -o gpurun_out/r2c_wet python tools/gpu/probe_steps.py 4096 22 gpurun_out/r2c_probe_ncu.json 0 > gpurun_out/r2c_ncu_wet.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_svat_step -s 40 -c 1 -o gpurun_out/r2c_post python tools/gpu/probe_steps.py 4096 42 gpurun_out/r2c_probe_ncu2.json 0 > gpurun_out/r2c_ncu_post.log 2>&1
tail -2 gpurun_out/r2c_plain.log
