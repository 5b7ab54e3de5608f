"""Host-side driver of the SVAT / oneD hot path: the part of `RogerSetup.setup()/step()/run()`
(roger/roger.py:258-580) and of the model hooks `set_forcing` / `set_parameters`
(roger/models/svat/svat.py:116-174) that surrounds the fused per-cell step.

The per-step compute is delegated to an *engine* (roger_b200.engine.CudaSvatEngine in the
product; the tests can plug the CPU oracle in through the same interface).  Everything here is
orchestration: once-per-run parameter derivation, once-per-day forcing hand-over, once-per-month
land-use parameter refresh.
"""
import os

import numpy as np

from . import abi, params

DAY = 24 * 60 * 60


class SvatModel:
    def __init__(self, engine_cls, shape, forcing, primary, theta_rz0, theta_ss0, lateral=False, weights=None,
                 settings=None, engine_kwargs=None, cell_order=None, z_gw_daily=None):
        self.shape = tuple(shape)
        self.C = int(np.prod(self.shape))
        self.lateral = bool(lateral)
        self.forcing = {k: np.asarray(v) for k, v in forcing.items()}
        self.primary = {k: np.asarray(v).reshape(-1) for k, v in primary.items()}
        self.theta_rz0 = np.asarray(theta_rz0, dtype=np.float64).reshape(-1)
        self.theta_ss0 = np.asarray(theta_ss0, dtype=np.float64).reshape(-1)
        self.weights = None if weights is None else {k: np.asarray(v, dtype=np.float64).reshape(-1) for k, v in weights.items()}
        # groundwater boundary (settings.enable_groundwater_boundary): z_gw_daily holds the depth of the water table (m) per cell and
        # day, (n_cells, ndays + 1) with column 0 the initial state -- the boundary condition the setups read from their forcing file
        # (examples/hillslope_scale/oneD_with_groundwater_boundary_condition_distributed_tutorial/oneD.py:317-361)
        self.z_gw_daily = None if z_gw_daily is None else np.asarray(z_gw_daily, dtype=np.float64).reshape(self.C, -1)
        settings = dict(settings or {})
        if self.z_gw_daily is not None:
            settings.setdefault("enable_groundwater_boundary", 1)
        self.settings = abi.default_settings(lateral=self.lateral, forcing_uniform=self.weights is None, **settings)
        self.engine = engine_cls(self.C, self.settings, **(engine_kwargs or {}))
        # device-side cell ordering by infiltration capacity (engine.optimize_cell_order); ROGER_B200_CELL_ORDER=0 switches it off
        self.cell_order = (os.environ.get("ROGER_B200_CELL_ORDER", "1") != "0") if cell_order is None else bool(cell_order)
        self.itt_forc = 0
        self.month = [1, 1]  # vs.month initial (roger/variables.py:318-326)
        self.n_steps = 0
        self._setup_done = False

    # ------------------------------------------------------------------ setup
    def setup(self, override_planes=None):
        """RogerSetup.setup (roger/roger.py:258-353) for this path: derive parameters, initial conditions."""
        p = self.primary
        d = params.soil_parameters(p, lateral=self.lateral, slope=p.get("slope"))
        ic = params.initial_conditions(p, d, self.theta_rz0, self.theta_ss0)
        zeros = np.zeros(self.C)
        sp = params.surface_parameters(p["lu_id"], self.month[1], zeros, zeros, p.get("c_int"))
        planes = {n: zeros for n in abi.FIELD_NAMES}
        planes.update({k: v for k, v in d.items() if k in abi.FIELD_ID})
        planes.update({k: v for k, v in ic.items() if k in abi.FIELD_ID})
        planes.update({k: v for k, v in sp.items() if k in abi.FIELD_ID})
        for k in ("lu_id", "z_soil", "dmpv", "lmpv", "theta_ac", "theta_ufc", "theta_pwp", "ks", "kf", "sealing", "dmph", "slope", "n0"):
            if k in p:
                planes[k] = p[k].astype(np.float64)
        planes["z_gw"] = np.full(self.C, 1000.0)  # vs.z_gw initial value (roger/variables.py:3248)
        if self.z_gw_daily is not None:
            planes["z_gw"] = self.z_gw_daily[:, 0].copy()
        planes["z_gw_m1"] = planes["z_gw"]
        if override_planes:
            planes.update(override_planes)
        for n in abi.FIELD_NAMES:
            self.engine.upload(n, planes[n])
        if self.cell_order and hasattr(self.engine, "optimize_cell_order"):
            self.engine.optimize_cell_order()
        self.engine.set_scalars(abi.initial_scalars())
        self._setup_done = True

    # ------------------------------------------------------------------ hooks
    def _set_forcing(self, sc):
        """models/svat/svat.py:150-174: at midnight hand the next 144 slots to the engine."""
        if sc.time % DAY != 0:
            return
        sc.itt_day = 0
        f = self.forcing
        i0 = self.itt_forc
        self.month[1] = int(f["MONTH"][i0])
        sc.month = self.month[1]
        sl = slice(i0, i0 + abi.SLOTS_PER_DAY)
        prec, ta, pet = f["PREC"][sl], f["TA"][sl], f["PET"][sl]
        if self.weights is not None:
            w = self.weights
            prec = prec[None, :] * w["prec_weight"][:, None]
            ta = ta[None, :] + w["ta_offset"][:, None]
            pet = pet[None, :] * w["pet_weight"][:, None]
        self.engine.set_forcing_day(prec, ta, pet)
        if self.z_gw_daily is not None:  # vs.z_gw[tau] of the day that begins
            self.engine.upload("z_gw", self.z_gw_daily[:, min(i0 // abi.SLOTS_PER_DAY + 1, self.z_gw_daily.shape[1] - 1)])
        self.itt_forc += abi.SLOTS_PER_DAY

    def _set_parameters(self, sc):
        """models/svat/svat.py:116-120: refresh land-use parameters on month change (and itt > 1)."""
        if (self.month[1] != self.month[0]) and (sc.itt > 1):
            lu = self.primary["lu_id"]
            sp = params.surface_parameters(lu, self.month[1], self.engine.download("ta"),
                                           self.engine.download("swe_top_tot"), self.primary.get("c_int"))
            # swe_top_tot is skipped: its ta-dependent where-chain (surface.py:236-301) is re-applied
            # verbatim by every step before first use (interception.py:170-205), so the refresh is idempotent.
            for k, v in sp.items():
                if k in abi.FIELD_ID and k != "swe_top_tot":
                    self.engine.upload(k, v)

    # ------------------------------------------------------------------ stepping
    def step(self, write_fluxes=True):
        """RogerSetup.step (roger/roger.py:355-489), online-hydrology branch."""
        assert self._setup_done
        sc = self.engine.get_scalars()
        self._set_forcing(sc)
        self.engine.set_scalars(sc)
        self._set_parameters(sc)
        self.engine.step(write_fluxes)
        self.month[0] = self.month[1]  # after_timestep
        self.n_steps += 1
        return self.engine.get_scalars()

    def run_steps(self, max_steps, write_fluxes=False):
        """Up to max_steps adaptive steps but never past the next midnight (where the hooks have to run again).
        Returns the number of steps taken.  Engines with `run_steps` may execute them in one launch."""
        assert self._setup_done
        sc = self.engine.get_scalars()
        self._set_forcing(sc)
        self.engine.set_scalars(sc)
        self._set_parameters(sc)
        if hasattr(self.engine, "run_steps"):
            n = self.engine.run_steps(max_steps, write_fluxes)
        else:
            n = 0
            while True:
                self.engine.step(write_fluxes)
                n += 1
                if n >= max_steps or self.engine.get_scalars().time % DAY == 0:
                    break
        self.month[0] = self.month[1]
        self.n_steps += n
        return n

    def checkpoint(self):
        """the engine's in-memory checkpoint plus the host-side cursor of the hooks"""
        self.engine.checkpoint()
        self._saved = (self.itt_forc, list(self.month), self.n_steps)

    def rollback(self):
        self.engine.rollback()
        self.itt_forc, self.month, self.n_steps = self._saved[0], list(self._saved[1]), self._saved[2]

    def run(self, runlen_seconds, write_fluxes=False, callback=None):
        sc = self.engine.get_scalars()
        start = sc.time
        while sc.time - start < runlen_seconds:
            sc = self.step(write_fluxes)
            if callback is not None:
                callback(self, sc)
        return sc

    def get(self, name):
        return self.engine.download(name)
