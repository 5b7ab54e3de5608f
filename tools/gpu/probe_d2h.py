"""GPU probe: bandwidth of the asynchronous plane hand-over (snapshot gather + D2H on the copy stream) with the GPU idle and
while the step kernel runs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from roger_b200.engine import CudaSvatEngine

m = bench.build_model(CudaSvatEngine, (4096, 4096), "oned", seed=43, engine_kwargs={"device": 0})
eng = m.engine
C = m.C
pin = [torch.empty(C, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
eng.reserve_snapshots(2)
for _ in range(8):
    m.run_steps(1)
eng.synchronize()
cases = [("under load, no copy", 24, -1), ("idle, copy engine", 0, 0), ("under load, copy engine", 24, 0)]
for nb in (4, 8, 16, 32, 64, 592):
    cases += [(f"idle, SM copy kernel x{nb} blocks", 0, nb), (f"under load, SM copy kernel x{nb} blocks", 24, nb)]
cases.append(("under load, no copy", 24, -1))
for label, steps, mode in cases:
    if mode >= 0:
        eng.set_d2h_mode(mode)
    t0 = time.perf_counter()
    if mode >= 0:
        eng.download_async("theta_rz", pin[0]); eng.download_async("S", pin[1])
    t1 = time.perf_counter()
    left = steps
    while left > 0:
        left -= m.run_steps(left)
    eng.synchronize()
    t2 = time.perf_counter()
    eng.download_wait()
    t3 = time.perf_counter()
    print(f"{label}: issue {1e3*(t1-t0):.2f} ms, steps {1e3*(t2-t1):.2f} ms, wait after steps {1e3*(t3-t2):.2f} ms, total {1e3*(t3-t0):.2f} ms; "
          f"268 MB -> {0.268/(t3-t0):.1f} GB/s lower bound")
# plain torch copy for reference
x = torch.empty(C * 2, dtype=torch.float64, device="cuda")
h = torch.empty(C * 2, dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(x, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"torch pinned D2H 268 MB idle: {1e3*(t1-t0):.2f} ms = {0.268/(t1-t0):.1f} GB/s")
