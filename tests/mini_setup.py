"""A minimal stand-in for a RoGeR setup object (TEST INFRASTRUCTURE): the attributes and hooks
roger_b200.integration.CudaStepDriver touches, with the reference's conventions -- arrays with two ghost cells on each
side, `(x, y, timesteps)` variables with tau = 1 / taum1 = 0, 0-d scalars, `vs.unlock()`, `state.var_meta[name].active`,
and the SVAT model's hooks (roger/models/svat/svat.py:116-174, :187-384).  It lets the GPU tests drive the CUDA engine
through the drop-in driver on a box where the reference package itself is not installed."""
import contextlib
import types

import numpy as np

from roger_b200 import abi, params

TWO_LEVEL = {"ta", "z_root", "ground_cover", "S_sur", "S_int_top", "S_int_ground", "S_dep", "S_snow", "swe", "S_rz", "S_ss", "S_s", "S",
             "z_sat", "z_wf", "z_wf_t0", "z_wf_t1", "y_mp", "y_sc", "theta_rz", "theta_ss", "theta", "k_rz", "k_ss", "k", "h_rz", "h_ss", "h",
             "z0", "prec", "swe_top", "swe_ground"}


class _Vars:
    tau, taum1, taup1 = 1, 0, 2

    @contextlib.contextmanager
    def unlock(self):
        yield


class MiniSvatSetup:
    def __init__(self, shape, forcing, primary, theta_rz0, theta_ss0, lateral=False, weights=None):
        nx, ny = shape
        self.shape = shape
        self.C = nx * ny
        self.forcing = forcing
        self.weights = weights
        self.primary = {k: np.asarray(v).reshape(-1) for k, v in primary.items()}
        self._setup_done = False
        st = types.SimpleNamespace(nx=nx, ny=ny, enable_lateral_flow=lateral, enable_offline_transport=False, runlen=0,
                                   enable_macropore_lower_boundary_condition=False, **params.SETTINGS_DEFAULTS)
        vs = _Vars()
        self.state = types.SimpleNamespace(variables=vs, settings=st, var_meta={})
        self._th = (np.asarray(theta_rz0, dtype=np.float64).reshape(-1), np.asarray(theta_ss0, dtype=np.float64).reshape(-1))

    def _alloc(self, name, interior):
        nx, ny = self.shape
        two = name in TWO_LEVEL
        a = np.zeros((nx + 4, ny + 4, 2) if two else (nx + 4, ny + 4))
        v = np.asarray(interior, dtype=np.float64).reshape(nx, ny)
        if two:
            a[2:-2, 2:-2, :] = v[:, :, None]
        else:
            a[2:-2, 2:-2] = v
        setattr(self.state.variables, name, a)
        self.state.var_meta[name] = types.SimpleNamespace(active=True)

    def setup(self):
        vs = self.state.variables
        p = self.primary
        lateral = self.state.settings.enable_lateral_flow
        d = params.soil_parameters(p, lateral=lateral, slope=p.get("slope"))
        ic = params.initial_conditions(p, d, *self._th)
        zeros = np.zeros(self.C)
        sp = params.surface_parameters(p["lu_id"], 1, zeros, zeros, p.get("c_int"))
        planes = {n: zeros for n in abi.FIELD_NAMES}
        for src in (d, ic, sp):
            planes.update({k: v for k, v in src.items() if k in abi.FIELD_ID})
        for k in ("lu_id", "z_soil", "dmpv", "lmpv", "theta_ac", "theta_ufc", "theta_pwp", "ks", "kf", "sealing", "dmph", "slope"):
            if k in p:
                planes[k] = p[k].astype(np.float64)
        planes["z_gw"] = np.full(self.C, 1000.0)
        for n in abi.FIELD_NAMES:
            if n == "maskRiverLake":
                continue
            self._alloc(n, planes[n])
        nx, ny = self.shape
        vs.maskRiver = np.zeros((nx + 4, ny + 4), dtype=bool)
        vs.maskLake = np.zeros((nx + 4, ny + 4), dtype=bool)
        vs.maskRiver[2:-2, 2:-2] = planes["maskRiverLake"].reshape(nx, ny) != 0
        for n in ("prec_day", "ta_day", "pet_day"):
            setattr(vs, n, np.zeros((nx + 4, ny + 4, abi.SLOTS_PER_DAY)))
        for k, v in dict(time=0, itt=0, dt_secs=3600, itt_day=0, event_id_counter=1, time_event0=0, itt_forc=0).items():
            setattr(vs, k, np.asarray(v, dtype=np.int64))
        vs.dt = np.asarray(1.0)
        vs.event_id = np.zeros(2, dtype=np.int64)
        vs.month = np.ones(2, dtype=np.int64)
        self._setup_done = True

    def _ensure_setup_done(self):
        assert self._setup_done

    # ---- hooks -----------------------------------------------------------------------------------------------
    def read_data(self, state):
        pass

    def set_boundary_conditions(self, state):
        pass

    def set_forcing(self, state):  # models/svat/svat.py:150-174
        vs = state.variables
        if int(vs.time) % 86400 != 0:
            return
        vs.itt_day = np.asarray(0, dtype=np.int64)
        i0 = int(vs.itt_forc)
        f = self.forcing
        month = np.array(vs.month)
        month[vs.tau] = int(f["MONTH"][i0])
        vs.month = month
        sl = slice(i0, i0 + abi.SLOTS_PER_DAY)
        nx, ny = self.shape
        for name, key, op in (("prec_day", "PREC", "prec_weight"), ("ta_day", "TA", "ta_offset"), ("pet_day", "PET", "pet_weight")):
            row = np.asarray(f[key][sl], dtype=np.float64)
            block = np.broadcast_to(row, (self.C, abi.SLOTS_PER_DAY)).copy()
            if self.weights is not None:
                w = np.asarray(self.weights[op], dtype=np.float64).reshape(-1, 1)
                block = block + w if op == "ta_offset" else block * w
            a = np.array(getattr(vs, name))
            a[2:-2, 2:-2, :] = block.reshape(nx, ny, abi.SLOTS_PER_DAY)
            setattr(vs, name, a)
        vs.itt_forc = np.asarray(i0 + abi.SLOTS_PER_DAY, dtype=np.int64)

    def set_parameters(self, state):  # models/svat/svat.py:116-120: monthly land-use parameters
        vs = state.variables
        if (vs.month[vs.tau] != vs.month[vs.taum1]) and (int(vs.itt) > 1):
            nx, ny = self.shape
            ta = vs.ta[2:-2, 2:-2, vs.tau].reshape(-1)
            swe_top_tot = vs.swe_top_tot[2:-2, 2:-2].reshape(-1)
            sp = params.surface_parameters(self.primary["lu_id"], int(vs.month[vs.tau]), ta, swe_top_tot, self.primary.get("c_int"))
            for k, v in sp.items():
                if k in abi.FIELD_ID and k != "swe_top_tot":
                    a = np.array(getattr(vs, k))  # a NEW array, as the reference's update() returns
                    if a.ndim == 3:
                        a[2:-2, 2:-2, vs.tau] = np.asarray(v, dtype=np.float64).reshape(nx, ny)
                    else:
                        a[2:-2, 2:-2] = np.asarray(v, dtype=np.float64).reshape(nx, ny)
                    setattr(vs, k, a)

    def after_timestep(self, state):  # models/svat/svat.py:187-384
        vs = state.variables
        for name in TWO_LEVEL:
            if not hasattr(vs, name):
                continue
            a = np.array(getattr(vs, name))
            a[2:-2, 2:-2, vs.taum1] = a[2:-2, 2:-2, vs.tau]
            setattr(vs, name, a)
        for name in ("event_id", "month"):
            a = np.array(getattr(vs, name))
            a[vs.taum1] = a[vs.tau]
            setattr(vs, name, a)
