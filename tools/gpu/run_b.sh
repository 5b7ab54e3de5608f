cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python tools/gpu/probe_steps.py 4096 120 gpurun_out/r2b_probe_4096.json 0
python tools/gpu/probe_steps.py 4096 120 gpurun_out/r2b_probe_4096_batch.json 2
python tools/gpu/probe_steps.py 2048 120 gpurun_out/r2b_probe_2048.json 0
