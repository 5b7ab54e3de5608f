"""Drop-in step driver for an existing RoGeR setup: runs `RogerSetup.step` (roger/roger.py:355-489, online SVAT / oneD
branch) with the per-cell work on the CUDA engine while the setup's own hooks (`read_data`, `set_boundary_conditions`,
`set_forcing`, `set_parameters`, `after_timestep`), its `RogerState`, `RogerVariables` and tau/taum1 time levels stay
exactly as the reference defines them.

    model = MySvatSetup()          # any RogerSetup subclass of the online SVAT / oneD family, unchanged
    model.setup()
    driver = CudaStepDriver(model) # uploads the post-setup state
    while vs.time < runlen:
        driver.step()              # instead of model.step(model.state)

`INTEGRATION.md` shows the few lines that make `roger.runtime_settings.backend = "cuda"` route `RogerSetup.step` here.
This module imports nothing from the reference: it only touches the objects it is handed (duck typing on
`setup.state.variables / settings / var_meta`), so it is importable without RoGeR installed.

Coherence contract.  The engine owns the per-cell state between steps.  After every step the variables named in `pull`
(default: every state and flux plane that is an active reference variable) are written back into `vs.<name>[2:-2, 2:-2
(, tau)]` so that `after_timestep`, diagnostics and user code observe the values the NumPy backend would have produced;
host-side changes of parameter arrays (e.g. the monthly land-use refresh in `set_parameters`, models/svat/svat.py:116-120)
are detected by array identity -- the reference's `update()` returns new arrays -- and re-uploaded before the step.
State arrays modified by user code between steps must be announced with `driver.push(name)`.
"""
import os

import numpy as np

from . import abi

DAY = 24 * 60 * 60
_SCALARS = ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0")


# settings (roger/settings.py:86-113) whose branches of RogerSetup.step the CUDA engine does not implement: a setup that
# switches one of them on is refused instead of being run without it
UNSUPPORTED_SETTINGS = (
    "enable_groundwater", "enable_groundwater_boundary", "enable_film_flow", "enable_routing_1D", "enable_routing_2D",
    "enable_runon_infiltration", "enable_urban", "enable_soil_compaction", "enable_time_variant_macropores",
    "enable_irrigation", "enable_net_irrigation", "enable_crop_phenology", "enable_crop_rotation",
    "enable_crop_specific_irrigation_demand", "enable_crop_partitioning", "enable_crop_water_stress", "enable_crop_optimization",
    "enable_nitrate", "enable_virtualtracer",
)


class CudaStepDriver:
    def __init__(self, setup, engine_cls=None, engine_kwargs=None, pull="all", forcing_uniform=None,
                 clamp_small_negative_storage=None):
        if engine_cls is None:
            from .engine import CudaSvatEngine as engine_cls  # no fallback: raises without the CUDA library / a device
        self.setup = setup
        self.state = setup.state
        vs, st = self.state.variables, self.state.settings
        if getattr(st, "enable_offline_transport", False):
            raise ValueError("CudaStepDriver drives the online SVAT / oneD models; use roger_b200.sas for offline transport")
        for flag in UNSUPPORTED_SETTINGS:  # every switch of roger/settings.py:86-113 that changes what the online step computes
            if getattr(st, flag, False):
                raise NotImplementedError(f"settings.{flag} is outside the CUDA hot path (DESIGN.md section 7)")
        if not getattr(st, "enable_adaptive_time_stepping", True):
            raise NotImplementedError("settings.enable_adaptive_time_stepping = False: the CUDA step IS the adaptive time step "
                                      "(adaptive_time_stepping.py:34-381)")
        self.nx, self.ny = int(st.nx), int(st.ny)
        self.C = self.nx * self.ny
        self.lateral = bool(getattr(st, "enable_lateral_flow", False))
        self._uniform = forcing_uniform
        self._clamp = (not self.lateral) if clamp_small_negative_storage is None else bool(clamp_small_negative_storage)
        self._engine_cls = engine_cls
        self._engine_kwargs = dict(engine_kwargs or {})
        self.engine = None
        meta = self.state.var_meta
        self._active = {n for n, m in meta.items() if m.active}
        kinds = dict(zip(abi.FIELD_NAMES, abi.FIELD_KINDS))
        if pull == "all":
            pull = [n for n in abi.FIELD_NAMES if kinds[n] in "SF" and n in self._active]
            pull.append("ks_ss")  # the one parameter plane a step writes (subsoil.calc_ks, roger/core/subsoil.py:6-20)
        self.pull = list(pull)
        self._write_fluxes = any(kinds[n] == "F" for n in self.pull)
        self._param_ids = {}
        self._pushed = set()

    # ------------------------------------------------------------------ vs <-> planes
    def _interior(self, name):
        vs = self.state.variables
        if name == "maskRiverLake":
            a = np.asarray(vs.maskRiver) | np.asarray(vs.maskLake)
            return a[2:-2, 2:-2].reshape(-1).astype(np.float64)
        a = np.asarray(getattr(vs, name))
        if a.ndim == 3:
            a = a[:, :, int(vs.tau)]
        return a[2:-2, 2:-2].reshape(-1).astype(np.float64)

    def _store(self, name, values):
        vs = self.state.variables
        cur = np.array(getattr(vs, name))  # copy: reference arrays are read-only
        v = values.reshape(self.nx, self.ny).astype(cur.dtype)
        if cur.ndim == 3:
            cur[2:-2, 2:-2, int(vs.tau)] = v
        else:
            cur[2:-2, 2:-2] = v
        setattr(vs, name, cur)
        if name in self._param_ids:
            self._param_ids[name] = id(getattr(vs, name))

    def _settings_struct(self, uniform):
        st = self.state.settings
        over = {}
        for k in ("ta_fm", "hpi", "end_event", "sf", "rmax", "r_mp", "l_sc", "pi", "transp_water_stress", "dx", "atol", "rtol"):
            if hasattr(st, k):
                over[k] = float(getattr(st, k))
        s = abi.default_settings(lateral=self.lateral, forcing_uniform=uniform, **over)
        s.enable_macropore_lower_boundary_condition = int(bool(getattr(st, "enable_macropore_lower_boundary_condition", False)))
        s.clamp_small_negative_storage = int(self._clamp)
        return s

    def _create_engine(self, uniform):
        self.engine = self._engine_cls(self.C, self._settings_struct(uniform), **self._engine_kwargs)
        vs = self.state.variables
        zeros = np.zeros(self.C)
        for name in abi.FIELD_NAMES:
            if name == "maskRiverLake" or name in self._active:
                self.engine.upload(name, self._interior(name))
                if abi.FIELD_KINDS[abi.FIELD_ID[name]] == "P" and name != "maskRiverLake":
                    self._param_ids[name] = id(getattr(vs, name))
            else:
                self.engine.upload(name, zeros)
        if hasattr(self.engine, "optimize_cell_order") and os.environ.get("ROGER_B200_CELL_ORDER", "1") != "0":
            self.engine.optimize_cell_order()  # device-side only: every call keeps the setup's cell numbering

    def push(self, name):
        """Announce that user code changed `vs.<name>` on the host; it is uploaded before the next step."""
        self._pushed.add(name)

    # ------------------------------------------------------------------ the step
    def _scalars_to_engine(self):
        vs = self.state.variables
        sc = abi.SvatScalars()
        for k in _SCALARS:
            setattr(sc, k, int(np.asarray(getattr(vs, k))))
        sc.event_id[0], sc.event_id[1] = int(vs.event_id[0]), int(vs.event_id[1])
        sc.month = int(vs.month[int(vs.tau)])
        sc.dt = float(np.asarray(vs.dt))
        sc.sanity_ok = 1
        self.engine.set_scalars(sc)

    def _scalars_from_engine(self):
        vs = self.state.variables
        sc = self.engine.get_scalars()
        with vs.unlock():
            for k in _SCALARS:
                setattr(vs, k, getattr(sc, k))
            vs.dt = sc.dt
            ev = np.array(vs.event_id)
            # engine.step ends with the after_timestep roll (event_id[taum1] <- event_id[tau]); the reference's hook redoes it
            ev[int(vs.tau)] = sc.event_id[1]
            vs.event_id = ev
        return sc

    def _hand_forcing_day(self):
        vs = self.state.variables
        blocks = [np.asarray(getattr(vs, n))[2:-2, 2:-2, :].reshape(self.C, abi.SLOTS_PER_DAY) for n in ("prec_day", "ta_day", "pet_day")]
        uniform_today = all(np.array_equal(b, np.broadcast_to(b[0], b.shape), equal_nan=True) for b in blocks)
        if self.engine is None:
            uniform = uniform_today if self._uniform is None else bool(self._uniform)
            self._create_engine(uniform)
            self._uniform = uniform
        if self._uniform:
            if not uniform_today:
                raise RuntimeError("forcing differs between cells but the engine was created for spatially uniform forcing; "
                                   "construct CudaStepDriver(..., forcing_uniform=False)")
            self.engine.set_forcing_day(*[b[0].copy() for b in blocks])
        else:
            self.engine.set_forcing_day(*[np.ascontiguousarray(b, dtype=np.float64) for b in blocks])

    def step(self):
        setup, state = self.setup, self.state
        vs = state.variables
        setup._ensure_setup_done()
        setup.read_data(state)
        setup.set_boundary_conditions(state)
        setup.set_forcing(state)
        if int(np.asarray(vs.time)) % DAY == 0 or self.engine is None:
            self._hand_forcing_day()
        setup.set_parameters(state)
        for name, old in list(self._param_ids.items()):  # host-side parameter refresh (monthly land use)
            if id(getattr(vs, name)) != old:
                self.engine.upload(name, self._interior(name))
                self._param_ids[name] = id(getattr(vs, name))
        for name in self._pushed:
            self.engine.upload(name, self._interior(name))
        self._pushed.clear()
        pend = getattr(self, "_pending_output", None)
        if pend is not None:  # enable_output() was called before the engine existed: the accumulators start with this step
            self._pending_output = None
            self._open_output(*pend)
        self._scalars_to_engine()
        self.engine.step(self._write_fluxes)
        sc = self._scalars_from_engine()
        with vs.unlock():
            for name in self.pull:
                self._store(name, self.engine.download(name))
        if not sc.sanity_ok:
            import warnings

            warnings.warn(f"Solution diverged at iteration {sc.itt} (numerics.sanity_check)")
        self._maybe_output(sc.time)
        setup.after_timestep(state)
        return sc

    def run(self, runlen=None):
        vs = self.state.variables
        runlen = float(self.state.settings.runlen) if runlen is None else runlen
        start = int(np.asarray(vs.time))
        while int(np.asarray(vs.time)) - start < runlen:
            self.step()
        self.close_output()

    # ------------------------------------------------------------------ output files of the setup's own diagnostics
    def enable_output(self, base_path=".", force_overwrite=False):
        """Writes the files the setup's `set_diagnostics` asked for -- `state.diagnostics["rate" | "average" | "maximum" |
        "minimum" | "collect"]` with their `output_variables`, `output_frequency` and `output_path` -- from device accumulators
        through `roger_b200.output.AsyncDiagnosticOutput` (accumulation after every step, the reference's sampling; record at
        the times `diagnostics.output` writes one, roger/diagnostics/api.py:45-66).  One output frequency for all of them."""
        import os

        st = self.state.settings
        spec, files, freq = [], {}, None
        for name in ("rate", "average", "maximum", "minimum", "collect"):
            d = self.state.diagnostics.get(name)
            if d is None or not d.output_variables or not d.output_frequency:
                continue
            if freq is not None and d.output_frequency != freq:
                raise NotImplementedError("diagnostics with different output frequencies")
            freq = d.output_frequency
            unknown = [v for v in d.output_variables if v not in abi.FIELD_ID]
            if unknown:
                raise NotImplementedError(f"diagnostic {name!r}: {unknown} are not planes of the CUDA engine")
            if any((v, m) for v, m in spec if v in d.output_variables):
                raise NotImplementedError("a variable in more than one diagnostic")
            spec += [(v, name) for v in d.output_variables]
            fname = d.output_path.format(identifier=st.identifier)
            files[os.path.join(str(getattr(d, "base_output_path", None) or base_path), fname)] = list(d.output_variables)
        if not spec:
            return None
        if freq not in (DAY, 3600, 600):
            raise NotImplementedError("output_frequency must be 86400, 3600 or 600 seconds")
        if self.engine is None:  # the engine exists from the first forcing day on
            self._pending_output = (spec, files, freq, force_overwrite)
            return None
        self._open_output(spec, files, freq, force_overwrite)
        return self._out

    def _open_output(self, spec, files, freq, force_overwrite):
        from .output import AsyncDiagnosticOutput

        st, vs = self.state.settings, self.state.variables
        self._out = AsyncDiagnosticOutput(self.engine, None, (self.nx, self.ny), spec, files=files, identifier=st.identifier,
                                          time_origin=st.time_origin, x=np.asarray(vs.x)[2:-2], y=np.asarray(vs.y)[2:-2],
                                          mask=self._interior("maskCatch"), force_overwrite=force_overwrite)
        self._out_freq = freq
        if any(abi.FIELD_KINDS[abi.FIELD_ID[v]] == "F" for v, _ in spec):
            self._write_fluxes = True

    def _maybe_output(self, time_seconds):
        out = getattr(self, "_out", None)
        if out is None:
            return
        t, f = int(time_seconds), self._out_freq
        if t % f == 0 or t % 3600 == 0 and f < 3600 or t % DAY == 0:  # roger/diagnostics/api.py:49-60
            out.output(t)

    def close_output(self):
        out = getattr(self, "_out", None)
        if out is not None:
            out.close()
            self._out = None


class CudaTransportDriver:
    """Drop-in for the offline oxygen-18 transport branch of `RogerSetup.step` (roger/roger.py:467-483): the setup's own
    hooks feed `vs.inf_mat_rz, vs.transp, ... vs.C_in` for the day, the engine runs
    `transport.svat_transport_model_deterministic` (core/transport.py:948-991) -- or, with `settings.sas_solver = "Euler"`,
    the day's explicit Euler steps (`svat_transport_model_euler`, :2063-2413, driven by :3221-3258) -- for every cell, and the results are written
    back into `vs` with both time levels as `after_substep_iso` (:3421-3520) leaves them.

        model = MyOxygen18Setup(); model.setup(); model.warmup()
        driver = CudaTransportDriver(model)
        for day in range(ndays): driver.step()

    `pull_state=False` leaves the four age-ranked arrays on the device (they are (cells x ages) each); the per-cell
    concentrations and delta values are always written back.
    """

    _FORCING = ("inf_mat_rz", "inf_pf_rz", "inf_pf_ss", "evap_soil", "transp", "q_rz", "q_ss", "cpr_rz", "C_in")

    def __init__(self, setup, engine_cls=None, engine_kwargs=None, pull_state=True, store_tt=True):
        from . import sas

        if engine_cls is None:
            from .sas_engine import CudaSasEngine as engine_cls  # no fallback
        self.sas = sas
        self.setup = setup
        self.state = setup.state
        vs, st = self.state.variables, self.state.settings
        iso18, iso2 = bool(getattr(st, "enable_oxygen18", False)), bool(getattr(st, "enable_deuterium", False))
        self.anion = bool(getattr(st, "enable_bromide", False)) or bool(getattr(st, "enable_chloride", False))
        if not getattr(st, "enable_offline_transport", False) or not (iso18 or iso2 or self.anion):
            raise ValueError("CudaTransportDriver drives the offline oxygen-18 / deuterium / bromide / chloride transport models")
        solver = getattr(st, "sas_solver", None)
        if solver not in ("deterministic", "Euler"):
            raise NotImplementedError("sas_solver = 'deterministic' and 'Euler' run on the device (DESIGN.md section 7)")
        self.euler = solver == "Euler"
        if self.euler and self.anion and bool(getattr(st, "enable_age_statistics", False)):
            raise NotImplementedError("the anion branch of the explicit Euler solver runs without age statistics")
        self.nx, self.ny, self.A = int(st.nx), int(st.ny), int(st.ages)
        self.C = self.nx * self.ny
        self.pull_state = bool(pull_state)
        self.stats = bool(getattr(st, "enable_age_statistics", False))
        # deuterium runs the same kernels with its own standard ratio and plausibility window (transport.py:315-340)
        vsmow, dmin, dmax = (st.VSMOW_conc2H, st.d2H_min, st.d2H_max) if iso2 else (st.VSMOW_conc18O, st.d18O_min, st.d18O_max)
        s = sas.SasSettings(self.A, int(st.sas_solver_substeps), float(vsmow), float(dmin), float(dmax), int(bool(store_tt)), int(self.stats),
                            sas.ANION if self.anion else sas.ISOTOPE, sas.EULER if self.euler else sas.DETERMINISTIC)
        self.settings = s
        self.engine = engine_cls(self.C, s, **(engine_kwargs or {}))
        tau = int(vs.tau)
        mask = np.asarray(vs.maskCatch)[2:-2, 2:-2].reshape(-1).astype(np.float64)
        block = np.stack([np.asarray(getattr(vs, "sas_params_" + p))[2:-2, 2:-2, :].reshape(self.C, 8) for p in sas.PROCESSES])
        self.engine.set_parameters(mask, block.astype(np.float64))
        if self.anion:
            inner = lambda n: np.asarray(getattr(vs, n))[2:-2, 2:-2].reshape(-1).astype(np.float64)  # noqa: E731
            self.engine.set_anion_parameters(inner("alpha_transp"), inner("alpha_q"), inner("lu_id"), inner("S_sat_rz"))
        for name in sas.STATE:
            self.engine.upload_state(name, np.asarray(getattr(vs, name))[2:-2, 2:-2, tau, :].reshape(self.C, self.A))
        if self.stats:
            self.engine.set_balance_reference(np.asarray(vs.C_s)[2:-2, 2:-2, int(vs.taum1)].reshape(-1))

    def _put(self, name, values, level=None):
        vs = self.state.variables
        cur = np.array(getattr(vs, name))
        v = np.asarray(values, dtype=cur.dtype)
        if level is None:
            cur[2:-2, 2:-2, ...] = v.reshape(cur[2:-2, 2:-2, ...].shape)
        else:
            for lv in level:
                cur[2:-2, 2:-2, lv, ...] = v.reshape(cur[2:-2, 2:-2, lv, ...].shape)
        setattr(vs, name, cur)

    def step(self):
        setup, state, sas = self.setup, self.state, self.sas
        vs = state.variables
        setup._ensure_setup_done()
        with vs.unlock():
            vs.itt = vs.itt + 1
            if not self.euler:  # roger.py:470-471
                vs.time = vs.time + vs.dt_secs
        setup.read_data(state)
        setup.set_boundary_conditions(state)
        setup.set_forcing(state)
        setup.set_parameters(state)
        if self.euler:  # the clock advances with the Euler steps (transport.py:3234), ageing when it reaches midnight (:3247)
            n = int(self.settings.substeps)
            with vs.unlock():
                vs.time = vs.time + n * int(vs.dt_secs / n)
                vs.itt_substep = n - 1
            if int(vs.time) % (24 * 60 * 60) != 0:
                raise NotImplementedError("Euler steps that do not add up to whole days (the device ages the StorAge once per call)")
        f = np.stack([np.asarray(getattr(vs, n))[2:-2, 2:-2].reshape(-1).astype(np.float64) for n in self._FORCING])
        self.engine.step(f)
        out = self.engine.download_outputs()
        tau, taum1 = int(vs.tau), int(vs.taum1)
        active = {n for n, m in state.var_meta.items() if m.active}
        with vs.unlock():
            for k, name in enumerate(sas.OUTPUTS):
                if self.anion and name.startswith("C_iso_"):  # anion transport: these planes carry the solute masses M_*
                    name = "M_" + name[len("C_iso_"):]
                if name not in active:
                    continue
                if np.asarray(getattr(vs, name)).ndim == 3:  # C_rz, C_ss, C_s, C_iso_*: both levels after after_substep_iso
                    self._put(name, out[k], level=(tau, taum1))
                else:
                    self._put(name, out[k])
            if self.pull_state:
                for name in sas.STATE:
                    self._put(name, self.engine.download_state(name), level=(tau, taum1))
            if self.settings.store_tt:
                for p in sas.PROCESSES:
                    if "tt_" + p in active:
                        self._put("tt_" + p, self.engine.download_tt(p))
            if self.stats:
                st = self.engine.download_statistics()
                for k, name in enumerate(sas.STATISTICS):
                    if name in active:
                        self._put(name, st[k])
        out_w = getattr(self, "_out", None)
        if out_w is not None and int(vs.time) % self._out_freq == 0:  # roger/diagnostics/api.py:45-66
            tau = int(vs.tau)
            vals = {}
            for f in out_w.files:
                for v in f.variables:
                    a = np.asarray(getattr(vs, v))
                    vals[v] = (a[2:-2, 2:-2, tau] if a.ndim == 3 else a[2:-2, 2:-2]).reshape(-1)
            out_w.output(int(vs.time), vals)
        setup.after_timestep(state)

    def enable_output(self, base_path=".", force_overwrite=False):
        """Writes the files the setup's `set_diagnostics` asked for (`state.diagnostics[...]` with `output_variables`,
        `output_frequency`, `output_path`) for the per-cell results of the transport step -- concentrations, delta values / solute
        masses, age statistics -- on a writer thread (`roger_b200.output.ThreadedRecordWriter`).  With one step per day every
        accumulating diagnostic (rate, average, maximum, minimum) holds the day's value, as collect does."""
        import os

        from .output import ThreadedRecordWriter

        st, vs = self.state.settings, self.state.variables
        files, freq = {}, None
        for name in ("rate", "average", "maximum", "minimum", "collect"):
            d = self.state.diagnostics.get(name)
            if d is None or not d.output_variables or not d.output_frequency:
                continue
            if freq is not None and d.output_frequency != freq:
                raise NotImplementedError("diagnostics with different output frequencies")
            freq = d.output_frequency
            for v in d.output_variables:
                a = np.asarray(getattr(vs, v))
                if a.ndim > 3 or (a.ndim == 3 and a.shape[2] != 2):
                    raise NotImplementedError(f"diagnostic {name!r}: {v} is not a per-cell variable (age-ranked arrays are not written)")
            fname = d.output_path.format(identifier=st.identifier)
            files[os.path.join(str(getattr(d, "base_output_path", None) or base_path), fname)] = list(d.output_variables)
        if not files:
            return None
        if freq % int(vs.dt_secs) != 0:
            raise NotImplementedError("output_frequency must be a multiple of the daily step")
        self._out = ThreadedRecordWriter((self.nx, self.ny), files, identifier=st.identifier, time_origin=st.time_origin,
                                         x=np.asarray(vs.x)[2:-2], y=np.asarray(vs.y)[2:-2],
                                         mask=np.asarray(vs.maskCatch)[2:-2, 2:-2].reshape(-1), force_overwrite=force_overwrite)
        self._out_freq = int(freq)
        return self._out

    def close_output(self):
        out = getattr(self, "_out", None)
        if out is not None:
            out.close()
            self._out = None
