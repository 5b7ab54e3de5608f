"""GPU diagnostic: which SAS family / solver combination separates the device from the oracle, and where (first differing cell)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle.engine import OracleSasEngine, lib  # noqa: E402
from roger_b200 import sas  # noqa: E402
from roger_b200.sas_engine import CudaSasEngine  # noqa: E402
from tests import util_sas as u  # noqa: E402
from tests.test_sas_gpu import _random_case  # noqa: E402

lib().sas_oracle_set_pow(1)
PW = dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))
for solver in (1, 2):
    for name, over in [("power only", {}), ("gamma transp", dict(transp=(4, 1.5, 4.0))), ("gamma q_ss", dict(q_ss=(4, 0.8, 2.5))),
                       ("52 evap", dict(evap_soil=(52, 2.0))), ("52 cpr", dict(cpr_rz=(52, 2.0))), ("52 q_ss + gamma q_rz", dict(q_ss=(52, 1.0), q_rz=(4, 2.0, 3.0))),
                       ("kum transp", dict(transp=(3, 0.7, 1.6))), ("uniform evap", dict(evap_soil=(1,)))]:
        C_, A, ndays = 64, 40, 4
        daily, _ = _random_case(C_, A, ndays, seed=77)
        spec = dict(PW, **over)
        ms = [sas.SasModel(e, C_, A, daily, spec, store_tt=True, substeps=3, solver=solver) for e in (CudaSasEngine, OracleSasEngine)]
        for m in ms:
            m.setup()
        msg = "equal"
        for day in range(ndays):
            for m in ms:
                m.step()
            sa, sb = u.snapshot(ms[0]), u.snapshot(ms[1])
            bad = [k for k in sa if not u.same(sa[k], sb[k])]
            if bad:
                k = bad[0]
                x, y = np.asarray(sa[k]), np.asarray(sb[k])
                neq = ~((x == y) | (np.isnan(x) & np.isnan(y)))
                idx = np.argwhere(neq)[0]
                msg = f"day {day}: {len(bad)} quantities differ, first {k} at {tuple(idx)}: cuda {x[tuple(idx)]!r} oracle {y[tuple(idx)]!r}; all: {bad[:8]}"
                break
        print(f"solver {solver} {name:24s} {msg}", flush=True)
