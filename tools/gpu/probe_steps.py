"""GPU probe: per-step kernel time of the oneD step over a long stretch, with clocks / power sampled every 20 ms.
usage: python tools/gpu/probe_steps.py SIDE NSTEPS OUT.json [batch]"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from roger_b200.engine import CudaSvatEngine  # noqa: E402

side, nsteps, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
samples = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active",
                         "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)


def reader():
    for line in proc.stdout:
        samples.append((time.time(), line.strip()))


threading.Thread(target=reader, daemon=True).start()
m = bench.build_model(CudaSvatEngine, (side, side), "oned", seed=43, engine_kwargs={"device": 0})
eng = m.engine
eng.set_block_stepping(batch, 0)
series = []
t_start = time.time()
done = 0
while done < nsteps:
    ms0 = eng.kernel_ms
    t0 = time.time()
    n = m.run_steps(1 if not batch else min(24, nsteps - done))
    eng.synchronize()
    series.append({"t": t0 - t_start, "n": n, "ms_per_step": (eng.kernel_ms - ms0) / n, "dts": eng.last_step_dts()[:2]})
    done += n
proc.terminate()
json.dump({"side": side, "batch": batch, "series": series, "smi": [(t - t_start, s) for t, s in samples if t >= t_start - 0.2]}, open(out, "w"))
ms = [s["ms_per_step"] for s in series]
print(side, "batch", batch, "first", [round(x, 3) for x in ms[:6]], "last", [round(x, 3) for x in ms[-6:]], "mean", sum(ms) / len(ms))
