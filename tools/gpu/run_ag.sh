cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "== default"; python tools/euler_timing.py 262144 1000 3 2>&1 | tail -n 2
echo "== code of commit 6a7dce3 (before the RK4 refactoring) with the round's final pow"
(cd tools/gpu/_old_wt && python tools/euler_timing.py 262144 1000 3 2>&1 | tail -n 2)
