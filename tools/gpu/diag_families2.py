"""GPU diagnostic: the failing seeded all-families cases, every differing quantity with its cells and family ids."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle.engine import OracleSasEngine, lib  # noqa: E402
from roger_b200 import sas  # noqa: E402
from roger_b200.sas_engine import CudaSasEngine  # noqa: E402
from tests import util_sas as u  # noqa: E402
from tests.test_sas_gpu import _random_case  # noqa: E402

lib().sas_oracle_set_pow(1)
np.set_printoptions(linewidth=200, precision=6)
for (solver, C_, A, ndays, substeps, masked) in [(2, 2, 9, 4, 1, True), (1, 3, 130, 3, 2, True), (1, 3, 130, 3, 2, False), (1, 3, 20, 3, 1, False)]:
    daily, spec = _random_case(C_, A, ndays, seed=4200 + (130 if A == 20 else A) + solver, families="all")
    mask = np.ones(C_)
    if masked:
        mask[4::13] = 0.0
    ms = [sas.SasModel(e, C_, A, daily, spec, store_tt=True, mask=mask, substeps=substeps, solver=solver) for e in (CudaSasEngine, OracleSasEngine)]
    for m in ms:
        m.setup()
    print(f"== solver {solver} C {C_} A {A} substeps {substeps} masked {masked}")
    for day in range(ndays):
        for m in ms:
            m.step()
        sa, sb = u.snapshot(ms[0]), u.snapshot(ms[1])
        bad = [k for k in sa if not u.same(sa[k], sb[k])]
        if bad:
            for k in bad:
                x, y = np.asarray(sa[k]), np.asarray(sb[k])
                neq = ~((x == y) | (np.isnan(x) & np.isnan(y)))
                cells = sorted(set(int(i[0]) for i in np.argwhere(neq)))
                print(f"day {day} {k}: cells {cells}")
                if x.ndim == 2 and k.startswith("tt_"):
                    c = cells[0]
                    print("   ids of the cell:", {p: float(np.asarray(spec[p])[c, 0]) for p in spec})
                    print("   cuda  ", x[c][:12])
                    print("   oracle", y[c][:12])
            break
    else:
        print("equal")
