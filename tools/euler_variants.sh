#!/bin/bash
# times prebuilt variants of the Euler kernel (roger_b200/lib/variants/*.so): block size : min blocks per SM
cd "$(dirname "$0")/.."
for lib in roger_b200/lib/libroger_b200.so roger_b200/lib/variants/*.so; do
  echo "== $lib"
  ROGER_B200_LIB=$PWD/$lib python - <<'PY' 2>&1 | tail -2
import numpy as np, sys
sys.path.insert(0, '.')
from roger_b200 import sas
from roger_b200.sas_engine import CudaSasEngine
from tests.test_sas_gpu import _random_case
for (C_, A) in ((262144, 1000),):
    daily, spec = _random_case(4096, A, 4, seed=3)
    daily = {k: np.tile(v, (C_ // 4096, 1)) for k, v in daily.items()}
    m = sas.SasModel(CudaSasEngine, C_, A, daily, spec, solver=1)
    m.setup()
    ts = []
    for d in range(3):
        m.engine.kernel_ms = 0; m.step(); ts.append(round(m.engine.kernel_ms, 1))
    print(C_, A, "ms per day", ts, "cell-days/s %.3e" % (C_ / (ts[-1] / 1e3)))
PY
done
