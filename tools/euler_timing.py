#!/usr/bin/env python
"""Device timing of the explicit Euler SAS day kernel: python tools/euler_timing.py CELLS AGES [DAYS]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from roger_b200 import sas  # noqa: E402
from roger_b200.sas_engine import CudaSasEngine  # noqa: E402
from tests.test_sas_gpu import _random_case  # noqa: E402

C_, A = int(sys.argv[1]), int(sys.argv[2])
days = int(sys.argv[3]) if len(sys.argv) > 3 else 3
daily, spec = _random_case(4096, A, days + 1, seed=3)
daily = {k: np.tile(v, (C_ // 4096, 1)) for k, v in daily.items()}
m = sas.SasModel(CudaSasEngine, C_, A, daily, spec, solver=sas.EULER)
m.setup()
for d in range(days):
    m.engine.kernel_ms = 0
    m.step()
    print(f"day {d}: kernel {m.engine.kernel_ms:.2f} ms -> {C_ / m.engine.kernel_ms * 1e3:.3e} cell-days/s", flush=True)
