#!/usr/bin/env python
"""Quick device timing of the StorAge-selection day kernel: python tools/sas_timing.py CELLS AGES [DAYS] [EMPTY_FRACTION]

EMPTY_FRACTION > 0 starts from an age distribution in which that fraction of the age bins is empty (the same bins in every
cell): the long-run state under spatially uniform forcing, where every dry day leaves an empty bin behind."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roger_b200 import sas  # noqa: E402
from roger_b200.sas_engine import CudaSasEngine  # noqa: E402

C_, A = int(sys.argv[1]), int(sys.argv[2])
days = int(sys.argv[3]) if len(sys.argv) > 3 else 4
empty = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
rng = np.random.default_rng(0)
eng = CudaSasEngine(C_, sas.default_settings(A))
spec = dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))
eng.set_parameters(np.ones(C_), sas.sas_parameter_block(C_, spec))
eng.init_linear(rng.uniform(60, 160, C_), rng.uniform(200, 700, C_), float(sas.delta_to_conc(-10.0)))
if empty > 0:
    keep = (np.random.default_rng(1).random(A) >= empty).astype(np.float64)
    keep[0] = 0
    for name in ("sa_rz", "sa_ss"):
        eng.upload_state(name, eng.download_state(name) * keep[None, :] / max(keep.mean(), 1e-9))
f = np.zeros((9, C_))
for d in range(days):
    f[0] = rng.uniform(0, 8) * (d % 2)
    f[3], f[4], f[5], f[6], f[7] = 0.6, 1.8, 1.1, 0.9, 0.05
    f[8] = float(sas.delta_to_conc(-7.0))
    eng.kernel_ms = 0
    t0 = time.perf_counter()
    eng.step(f)
    wall = time.perf_counter() - t0
    print(f"day {d}: kernel {eng.kernel_ms:.2f} ms  wall {wall * 1e3:.2f} ms  -> {C_ / eng.kernel_ms * 1e3:.3e} cell-days/s, "
          f"{(64 * A + 344) * C_ / eng.kernel_ms / 1e6:.1f} GB/s algorithmic", flush=True)
