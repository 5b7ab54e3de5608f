"""Child process of tests/test_cuda_backend_patch.py (TEST INFRASTRUCTURE): runs the reference's SHIPPED model
`roger.models.svat.svat.SVATSetup` -- imported from a patched copy of the reference package that is first on sys.path -- with
the backend named by ROGER_BACKEND, through the reference's own `RogerSetup.setup()` / `RogerSetup.step()`, and dumps what it
computed.  File IO is stubbed: the three helper methods of the shipped setup that open `forcing.nc` are overridden to serve an
in-memory synthetic series.  With ROGER_B200_TEST_ENGINE=oracle the cuda backend's engine class is replaced by the CPU oracle
(the build container has no GPU); without it the CUDA engine is used and construction fails loudly when there is none."""
import os
import sys
import types

import numpy as np

for name in ("h5netcdf", "h5py", "cftime"):  # absent offline; never called (diskless mode, stubbed helpers)
    m = types.ModuleType(name)
    m.File = None
    sys.modules[name] = m

out_path, nx, ny, ndays = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])

import roger  # noqa: E402  (the patched copy)
from roger import runtime_settings as rs  # noqa: E402

assert rs.backend == os.environ["ROGER_BACKEND"], rs.backend  # read from the environment by roger/runtime.py
rs.update(float_type="float64", int_type="int64", diskless_mode=True, loglevel="error")

from roger import roger_routine  # noqa: E402
from roger.models.svat.svat import SVATSetup  # noqa: E402  (the shipped model)

from roger_b200 import synth  # noqa: E402

if rs.backend == "cuda" and os.environ.get("ROGER_B200_TEST_ENGINE") == "oracle":
    from oracle.engine import OracleSvatEngine
    from roger_b200 import integration

    integration.DEFAULT_ENGINE = OracleSvatEngine

F = synth.mixed_forcing(ndays, seed=42, start_doy=100)
n = len(F["PREC"])
DATA = {"PREC": F["PREC"][None, None, :], "TA": F["TA"][None, None, :], "PET": F["PET"][None, None, :], "Time": np.arange(n),
        "YEAR": F["YEAR"], "MONTH": F["MONTH"], "DOY": F["DOY"]}


class StubbedIO(SVATSetup):
    def _read_var_from_nc(self, var, path_dir, file):
        return np.asarray(DATA[var])

    def _get_runlen(self, path_dir, file):
        return ndays * 86400

    def _get_time_origin(self, path_dir, file):
        return "2022-12-31 00:00:00"

    @roger_routine
    def set_settings(self, state):
        super().set_settings(state)  # the shipped model's settings; only the grid size is changed
        s = state.settings
        s.nx, s.ny = nx, ny

    @roger_routine
    def set_parameters_setup(self, state):
        from roger.core.operators import update, at

        vs = state.variables
        prim, _, _ = synth.heterogeneous_parameters((nx, ny), seed=5)
        for k, v in prim.items():
            cur = getattr(vs, k)
            setattr(vs, k, update(cur, at[2:-2, 2:-2], np.asarray(v).astype(cur.dtype)))
        vs.sealing = update(vs.sealing, at[2:-2, 2:-2], 0)
        vs.S_dep_tot = update(vs.S_dep_tot, at[2:-2, 2:-2], 0)

    @roger_routine
    def set_initial_conditions(self, state):
        from roger.core.operators import update, at

        vs = state.variables
        _, th_rz, th_ss = synth.heterogeneous_parameters((nx, ny), seed=5)
        vs.theta_rz = update(vs.theta_rz, at[2:-2, 2:-2, : vs.taup1], th_rz[:, :, None])
        vs.theta_ss = update(vs.theta_ss, at[2:-2, 2:-2, : vs.taup1], th_ss[:, :, None])


model = StubbedIO()
model.setup()
state = model.state
vs = state.variables
names = ["S_fp_rz", "S_lp_rz", "S_fp_ss", "S_lp_ss", "theta_rz", "theta_ss", "S", "S_snow", "swe", "z0", "z_sat", "k_rz", "h_ss",
         "q_ss", "q_hof", "aet", "inf_mat", "inf_mp", "prec", "ta", "S_int_top", "S_int_ground", "ground_cover"]
lazy = rs.backend == "cuda" and os.environ.get("ROGER_CUDA_MIRROR") == "lazy"
if lazy:  # read three state variables per step: that is all the lazy mirror should download
    names = ["S", "theta_rz", "swe"]
rec = {k: [] for k in names}
refused = 0
scal = []
while vs.time < state.settings.runlen:
    model.step(state)  # RogerSetup.step of the (patched) reference: with the cuda backend it routes to roger_b200
    scal.append([float(np.asarray(getattr(vs, k))) for k in ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt")]
                + [float(vs.event_id[0]), float(vs.event_id[1]), float(vs.month[1])])
    for k in names:
        rec[k].append(np.asarray(getattr(vs, k)).copy())
    if lazy and refused == 0:
        try:
            vs.q_ss  # a flux the step did not store: refused, not served stale
        except RuntimeError:
            refused = 1
drv = getattr(model, "_cuda_driver", None)
np.savez(out_path, scalars=np.asarray(scal), backend=rs.backend, driver=type(drv).__name__ if drv is not None else "",
         engine=type(drv.engine).__name__ if drv is not None else "", refused=refused, downloads=(drv.downloads if drv is not None else -1),
         **{k: np.stack(v) for k, v in rec.items()})
