set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_sas_gpu.py -x -q -k "euler" 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -8
import time, numpy as np
from roger_b200 import sas
from roger_b200.sas_engine import CudaSasEngine
import sys; sys.path.insert(0,'.')
from tests.test_sas_gpu import _random_case
for solver in (0, 1):
    C_, A = 262144, 1000
    daily, spec = _random_case(4096, A, 5, seed=3)
    daily = {k: np.tile(v, (C_ // 4096, 1)) for k, v in daily.items()}
    m = sas.SasModel(CudaSasEngine, C_, A, daily, spec, solver=solver)
    m.setup()
    ts = []
    for d in range(4):
        m.engine.kernel_ms = 0; m.step(); ts.append(round(m.engine.kernel_ms, 2))
    print("solver", solver, "kernel ms per day", ts, "cell-days/s %.3e" % (C_ / (ts[-1] / 1e3)))
PY
