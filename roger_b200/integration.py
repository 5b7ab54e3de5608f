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

Coherence contract.  The engine owns the per-cell state between steps; `vs.<name>` are MIRRORS of its planes
(`RogerVariables.__getattr__ / __setattr__`, roger/state.py:213-251, extended on the instance by `_install_mirror`):

* host -> device: any assignment `vs.<name> = ...` of a variable the engine holds (the monthly land-use refresh of
  `set_parameters`, models/svat/svat.py:116-120; a user hook changing a state array) marks it dirty and it is uploaded
  before the next step.  Nothing has to be announced.
* device -> host, `mirror="eager"` (default): after every step the variables named in `pull` (default: every state and flux
  plane that is an active reference variable) are written back into `vs.<name>[2:-2, 2:-2(, tau)]`, then the setup's own
  `after_timestep` runs: hooks, diagnostics and user code observe exactly what the NumPy backend would have produced, both
  time levels included.  This is the mode the parity tests run.
* device -> host, `mirror="lazy"`: after a step the mirrors are only marked stale; a variable is downloaded when somebody
  reads `vs.<name>` (a diagnostic, a user hook, the monthly parameter refresh reading `vs.ta`).  The setup's `after_timestep`
  -- in every shipped model a pure tau -> taum1 roll that reads all thirty rolled variables -- is replaced by the engine's
  own roll (single time level on the device); a materialised mirror shows the step's result in BOTH time levels, i.e. the
  state `after_timestep` leaves behind.  Flux variables are stored by the step only when they are listed in `fluxes` or
  named by one of the setup's diagnostics; reading any other flux variable raises instead of returning stale numbers.
"""
import os

import numpy as np

from . import abi

DAY = 24 * 60 * 60
_SCALARS = ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0")


# settings (roger/settings.py:86-113) whose branches of RogerSetup.step the CUDA engine does not implement: a setup that
# switches one of them on is refused instead of being run without it
UNSUPPORTED_SETTINGS = (
    "enable_groundwater", "enable_film_flow", "enable_routing_1D", "enable_routing_2D",
    "enable_runon_infiltration", "enable_urban", "enable_soil_compaction", "enable_time_variant_macropores",
    "enable_irrigation", "enable_net_irrigation", "enable_crop_phenology", "enable_crop_rotation",
    "enable_crop_specific_irrigation_demand", "enable_crop_partitioning", "enable_crop_water_stress", "enable_crop_optimization",
    "enable_nitrate", "enable_virtualtracer",
)


class CudaStepDriver:
    def __init__(self, setup, engine_cls=None, engine_kwargs=None, pull="all", forcing_uniform=None,
                 clamp_small_negative_storage=None, mirror="eager", fluxes=None, call_after_timestep=None):
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
        if mirror not in ("eager", "lazy"):
            raise ValueError("mirror must be 'eager' or 'lazy'")
        self.mirror = mirror
        self.call_after_timestep = (mirror == "eager") if call_after_timestep is None else bool(call_after_timestep)
        if mirror == "lazy":
            want = set(fluxes or ())
            for d in (getattr(self.state, "diagnostics", None) or {}).values():  # what the setup's own diagnostics will read
                want |= set(getattr(d, "output_variables", None) or ())
            flux_names = [n for n in abi.FIELD_NAMES if kinds[n] == "F" and n in self._active]
            self._untracked_fluxes = set(flux_names) - want
            pull = [n for n in abi.FIELD_NAMES if (kinds[n] == "S" or (kinds[n] == "F" and n in want)) and n in self._active]
            pull.append("ks_ss")
        else:
            self._untracked_fluxes = set()
        self.pull = list(pull)
        self._write_fluxes = any(kinds[n] == "F" for n in self.pull)
        self._device_names = set(n for n in abi.FIELD_NAMES if n in self._active and n != "maskRiverLake")
        self._dirty = set()      # assigned on the host since the last upload
        self._stale = set()      # lazy mode: the device holds a newer value than the host array
        self._shadow = {}        # lazy mode: the host arrays of stale / untracked variables (ghost cells, both levels)
        self._writing_back = False
        self._uploaded_fp = {}   # fingerprint of the last upload of every plane
        # planes a step writes: the uploaded fingerprint says nothing about their current device content
        self._device_written = set(n for n in abi.FIELD_NAMES if kinds[n] in "SF") | {"ks_ss"}
        self.downloads = 0       # planes downloaded through the mirrors (bench / tests)
        self.uploads = 0

    # ------------------------------------------------------------------ vs <-> planes
    def _interior(self, name):
        vs = self.state.variables
        if name == "maskRiverLake":
            a = np.asarray(vs.maskRiver) | np.asarray(vs.maskLake)
            return a[2:-2, 2:-2].reshape(-1).astype(np.float64)
        if name == "z_gw_m1":  # the previous step's water table (groundwater boundary): vs.z_gw[taum1]
            a = np.asarray(vs.z_gw)
            a = a[:, :, int(vs.taum1)] if a.ndim == 3 else a
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
        self._writing_back = True
        try:
            setattr(vs, name, cur)
        finally:
            self._writing_back = False

    # ------------------------------------------------------------------ mirrors (roger/state.py:213-251 on this instance)
    def _install_mirror(self):
        """Gives `vs` (whatever class it is: the reference's RogerVariables or a stand-in) the two mirror hooks: an assignment
        of a device-held variable marks it dirty; a read of a stale one downloads it."""
        vs = self.state.variables
        base = vs.__class__
        if getattr(base, "_roger_b200_mirrored", False):
            object.__setattr__(vs, "_cuda_driver", self)
            return
        driver_of = lambda obj: obj.__dict__.get("_cuda_driver")  # noqa: E731

        def __setattr__(obj, name, val):
            base.__setattr__(obj, name, val)
            drv = driver_of(obj)
            if drv is not None and not drv._writing_back and name in drv._device_names:
                drv._dirty.add(name)
                drv._stale.discard(name)
                drv._shadow.pop(name, None)

        def __getattr__(obj, name):
            drv = obj.__dict__.get("_cuda_driver")
            if drv is not None:
                if name in drv._stale:
                    return drv._materialise(name)
                if name in drv._untracked_fluxes and name in drv._shadow:
                    raise RuntimeError(f"vs.{name} is a flux the lazy cuda mirror does not store (it is written by the step only on "
                                       f"request): construct CudaStepDriver(..., fluxes=[{name!r}, ...]) or name it in a diagnostic")
            parent = getattr(base, "__getattr__", None)
            if parent is not None:
                return parent(obj, name)
            raise AttributeError(name)

        cls = type("CudaMirrored" + base.__name__, (base,), {"__setattr__": __setattr__, "__getattr__": __getattr__,
                                                               "_roger_b200_mirrored": True})
        vs.__class__ = cls
        object.__setattr__(vs, "_cuda_driver", self)

    def _materialise(self, name):
        """a read of a stale mirror: download the plane into the host array (both time levels show the step's result, i.e.
        the state after `after_timestep`); works on a locked `vs` -- reading is not a modification"""
        vs = self.state.variables
        cur = np.array(self._shadow.pop(name))
        self._stale.discard(name)
        self.downloads += 1
        v = self.engine.download(name).reshape(self.nx, self.ny).astype(cur.dtype)
        if cur.ndim == 3:
            cur[2:-2, 2:-2, : int(vs.tau) + 1] = v[:, :, None]
        else:
            cur[2:-2, 2:-2] = v
        cur.flags.writeable = False  # as the reference's own arrays (roger/variables.py:6432)
        vs.__dict__[name] = cur
        return cur

    def _mark_stale(self):
        d = self.state.variables.__dict__
        for name in self.pull:
            if name in d:
                self._shadow[name] = d.pop(name)
            self._stale.add(name)
        for name in self._untracked_fluxes:
            if name in d:
                self._shadow[name] = d.pop(name)

    @staticmethod
    def _fingerprint(values):
        """64-bit fingerprint of a plane's bit patterns (cheap stand-in for keeping a host copy of every uploaded plane)"""
        bits = np.ascontiguousarray(values, dtype=np.float64).view(np.uint64)
        w = np.arange(1, bits.size + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            return int(np.bitwise_xor.reduce(bits)) ^ int((bits * w).sum(dtype=np.uint64))

    def _upload(self, name):
        v = self._interior(name)
        self.engine.upload(name, v)
        self._uploaded_fp[name] = self._fingerprint(v)

    def _upload_dirty(self):
        for name in sorted(self._dirty):
            v = self._interior(name)
            fp = self._fingerprint(v)
            if name not in self._stale and self._uploaded_fp.get(name) == fp and name not in self._device_written:
                continue  # re-assigned with the values the device already holds (e.g. a tau -> taum1 roll of a parameter array)
            self.engine.upload(name, v)
            self._uploaded_fp[name] = fp
            self.uploads += 1
        self._dirty.clear()

    def _settings_struct(self, uniform):
        st = self.state.settings
        over = {}
        for k in ("ta_fm", "hpi", "end_event", "sf", "rmax", "r_mp", "l_sc", "pi", "transp_water_stress", "dx", "atol", "rtol"):
            if hasattr(st, k):
                over[k] = float(getattr(st, k))
        s = abi.default_settings(lateral=self.lateral, forcing_uniform=uniform, **over)
        s.enable_macropore_lower_boundary_condition = int(bool(getattr(st, "enable_macropore_lower_boundary_condition", False)))
        s.clamp_small_negative_storage = int(self._clamp)
        s.enable_groundwater_boundary = int(bool(getattr(st, "enable_groundwater_boundary", False)))
        s.kf_max = float(getattr(st, "kf_max", 3600.0))
        return s

    def _create_engine(self, uniform):
        self.engine = self._engine_cls(self.C, self._settings_struct(uniform), **self._engine_kwargs)
        vs = self.state.variables
        zeros = np.zeros(self.C)
        for name in abi.FIELD_NAMES:
            if name == "maskRiverLake" or name in self._active or (name == "z_gw_m1" and "z_gw" in self._active):
                self._upload(name)
            else:
                self.engine.upload(name, zeros)
        if hasattr(self.engine, "optimize_cell_order") and os.environ.get("ROGER_B200_CELL_ORDER", "1") != "0":
            self.engine.optimize_cell_order()  # device-side only: every call keeps the setup's cell numbering
        self._dirty.clear()
        self._install_mirror()

    def push(self, name):
        """Kept for callers of the first version: assignments to `vs.<name>` are tracked by the mirror itself; in-place edits
        of a (writeable copy of a) state array can still be announced here."""
        self._dirty.add(name)

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

    @staticmethod
    def _is_uniform(a):
        """every cell of the (nx, ny, slots) day block holds the row of cell (0, 0) (NaN == NaN, -0.0 == 0.0).  One chunked pass over the
        bit patterns, no copy of the block: 0.3 s instead of 5 s per variable at 1024 x 1024 cells (the setups' `update(vs.prec_day, ...)`
        materialises the block even when all cells share one series)."""
        if a.dtype == np.float64 and a.strides[-1] == a.itemsize:
            ai = a.view(np.int64)
            ref = ai[0, 0, :]
            if all(bool((ai[i0:i0 + 64] == ref).all()) for i0 in range(0, ai.shape[0], 64)):
                return True
        ref = a[0, 0, :]
        return bool(np.all((a == ref) | (np.isnan(a) & np.isnan(ref))))

    def _hand_forcing_day(self):
        vs = self.state.variables
        views = [np.asarray(getattr(vs, n))[2:-2, 2:-2, :] for n in ("prec_day", "ta_day", "pet_day")]
        uniform_today = None
        if self.engine is None or self._uniform:  # (an engine built for per-cell forcing takes whatever comes)
            uniform_today = all(self._is_uniform(a) for a in views)
        if self.engine is None:
            uniform = uniform_today if self._uniform is None else bool(self._uniform)
            self._create_engine(uniform)
            self._uniform = uniform
        if self._uniform:
            if not uniform_today:
                raise RuntimeError("forcing differs between cells but the engine was created for spatially uniform forcing; "
                                   "construct CudaStepDriver(..., forcing_uniform=False)")
            self.engine.set_forcing_day(*[np.array(a[0, 0, :], dtype=np.float64) for a in views])
        else:
            self.engine.set_forcing_day(*[np.ascontiguousarray(a.reshape(self.C, abi.SLOTS_PER_DAY), dtype=np.float64) for a in views])

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
        self._upload_dirty()  # whatever the hooks assigned: the monthly land-use refresh, user edits of state arrays
        pend = getattr(self, "_pending_output", None)
        if pend is not None:  # enable_output() was called before the engine existed: the accumulators start with this step
            self._pending_output = None
            self._open_output(*pend)
        self._scalars_to_engine()
        self.engine.step(self._write_fluxes)
        sc = self._scalars_from_engine()
        pulled = {}
        if self.mirror == "eager":
            with vs.unlock():
                for name in self.pull:
                    pulled[name] = self.engine.download(name)
                    self.downloads += 1
                    self._store(name, pulled[name])
        else:
            self._mark_stale()
        if not sc.sanity_ok:
            import warnings

            warnings.warn(f"Solution diverged at iteration {sc.itt} (numerics.sanity_check)")
        self._maybe_output(sc.time)
        tail = getattr(self, "reference_tail", None)
        if tail is not None:  # the reference's own diagnostics on the mirrors (cuda_step below)
            tail(sc)
        if self.call_after_timestep:
            setup.after_timestep(state)
            # the hook's tau -> taum1 roll re-assigns the rolled arrays: only values that really differ from what the device
            # holds count as host-side changes
            for name in list(self._dirty):
                if name in pulled and np.array_equal(self._interior(name), pulled[name], equal_nan=True):
                    self._dirty.discard(name)
        else:
            with vs.unlock():  # the part of every after_timestep hook that is not a roll of device-held arrays
                for name in ("event_id", "year", "month", "doy"):
                    if hasattr(vs, name):
                        a = np.array(getattr(vs, name))
                        if a.ndim == 1 and a.shape[0] >= 2:
                            a[int(vs.taum1)] = a[int(vs.tau)]
                            setattr(vs, name, a)
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
            for v in d.output_variables:
                mode = name
                if name in ("maximum", "minimum") and np.asarray(getattr(self.state.variables, v)).ndim == 3:
                    mode = name + "_t"  # (x, y, timesteps) variables: the extreme of the last two time levels (maximum.py:60-75)
                spec.append((v, mode))
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


def cuda_step(setup, state=None, **driver_kwargs):
    """What the `cuda` branch of `RogerSetup.step` calls (integration/roger_cuda_backend.patch): one time step of `setup`
    with the per-cell work on the GPU.  The driver is created on first use and kept on the setup object; the reference's
    own diagnostics (`diagnostics.reset / diagnose / output`, roger/roger.py:453-466) and restart hook run unchanged on
    the mirrors when the reference package is importable."""
    drv = getattr(setup, "_cuda_driver", None)
    if drv is None:
        st = setup.state.settings
        kw = dict(driver_kwargs)
        if getattr(st, "enable_offline_transport", False):
            drv = CudaTransportDriver(setup, engine_cls=kw.get("engine_cls", DEFAULT_TRANSPORT_ENGINE))
        else:
            kw.setdefault("engine_cls", DEFAULT_ENGINE)
            kw.setdefault("mirror", os.environ.get("ROGER_CUDA_MIRROR", "eager"))
            drv = CudaStepDriver(setup, **kw)
            drv.reference_tail = _reference_tail(setup)
        setup._cuda_driver = drv
    return drv.step()


def _reference_tail(setup):
    """roger/roger.py:447-466 after the storages: the diagnostics block of the reference, on the mirrors"""
    try:
        from roger import diagnostics
    except Exception:  # stand-in setups without the reference package
        return None
    state = setup.state

    def tail(sc):
        vs, settings = state.variables, state.settings
        if getattr(settings, "warmup_done", False):
            with vs.unlock():
                if vs.time_for_diag >= settings.output_frequency:
                    vs.time_for_diag = 0
                diagnostics.reset(state)
                vs.time_for_diag = vs.time_for_diag + vs.dt_secs
            diagnostics.diagnose(state)
            diagnostics.output(state)

    return tail


DEFAULT_ENGINE = None            # None: roger_b200.engine.CudaSvatEngine (no fallback); tests may plug the CPU oracle in here
DEFAULT_TRANSPORT_ENGINE = None  # None: roger_b200.sas_engine.CudaSasEngine


class CudaTransportDriver:
    """Drop-in for the offline oxygen-18 transport branch of `RogerSetup.step` (roger/roger.py:467-483): the setup's own
    hooks feed `vs.inf_mat_rz, vs.transp, ... vs.C_in` for the day, the engine runs
    `transport.svat_transport_model_deterministic` (core/transport.py:948-991) -- or, with `settings.sas_solver = "Euler"` / `"RK4"`,
    the day's explicit steps (`svat_transport_model_euler`, :2063-2413, `svat_transport_model_rk4`, :1138-2047, driven by :3221-3302) -- for
    every cell, and the results are written back into `vs` with both time levels as `after_substep_iso` (:3421-3520) leaves them.
    Solutes: oxygen-18 / deuterium, bromide / chloride, nitrate (`settings.enable_nitrate`: + the soil nitrogen cycle) and the virtual
    tracer (`settings.enable_virtualtracer`: the anion model whose soil evaporation carries the tracer, evapotranspiration.py:723-791;
    such a setup writes the day's `vs.C_in` itself -- `C_IN` is not an active variable there, variables.py:5497-5512).
    Variants: `settings.enable_crop_phenology` (anion / nitrate / virtual tracer, with or without age statistics: redistribution after
    root growth / root loss opens the day), `settings.enable_groundwater_boundary` (same transport; the reference computes no balance
    errors then).  Not driven: the crop variant of the ISOTOPE models (its redistribution kernel ADDS the mixed concentration to the
    root zone's instead of replacing it, crop.py:2406-2425 `update_add` -- concentrations double per day) and of the Euler scheme (its
    anion branch reads `vs.q_re_rg`, a variable the reference does not declare, transport.py:3012).
    Deviations kept on purpose: one result record per day under the Euler / RK4 schemes (the reference calls `write_output` after
    every step of the day, transport.py:3240); the derived age-ranked arrays `SA_*`, `MSA_*`, `csa_*`, `TT_*` are not written back.

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
        self.nitrate = bool(getattr(st, "enable_nitrate", False))
        # the virtual tracer is the anion model whose soil evaporation carries the solute (evapotranspiration.py:1014-1016)
        self.tracer = bool(getattr(st, "enable_virtualtracer", False))
        self.anion = bool(getattr(st, "enable_bromide", False)) or bool(getattr(st, "enable_chloride", False)) or self.nitrate or self.tracer
        if not getattr(st, "enable_offline_transport", False) or not (iso18 or iso2 or self.anion):
            raise ValueError("CudaTransportDriver drives the offline oxygen-18 / deuterium / bromide / chloride / nitrate / virtual-tracer "
                             "transport models")
        solver = getattr(st, "sas_solver", None)
        if solver not in ("deterministic", "Euler", "RK4"):
            raise NotImplementedError(f"sas_solver = {solver!r}: 'deterministic', 'Euler' and 'RK4' run on the device")
        self.euler = solver in ("Euler", "RK4")  # the schemes that take `sas_solver_substeps` steps per day (transport.py:3221-3302)
        if solver == "RK4" and self.anion:
            raise NotImplementedError("the RK4 solver runs the isotope branch only (the reference's RK4 kernels have no anion branch that runs)")
        stats = bool(getattr(st, "enable_age_statistics", False))
        # variants of the offline model (calculate_storage_selection, transport.py:3135-3396).  crop: redistribution after root growth
        # / root loss opens the day (svat_crop_transport_model_deterministic :1042-1087); groundwater boundary: svat_lbc_transport_model_
        # deterministic (:995-1040) adds the capillary rise into the subsoil, whose transport kernels are empty stubs
        # (capillary_rise.py:590-611), and the reference then computes no balance errors (numerics.py:423-428, :1380-1387)
        self.crop = bool(getattr(st, "enable_crop_phenology", False))
        self.lbc = bool(getattr(st, "enable_groundwater_boundary", False))
        for flag in ("enable_groundwater", "enable_routing_1D", "enable_routing_2D"):
            if getattr(st, flag, False):
                raise NotImplementedError(f"offline transport with settings.{flag} is outside the CUDA hot path (DESIGN.md section 7)")
        if self.crop and (not self.anion or self.euler or self.lbc):
            raise NotImplementedError("settings.enable_crop_phenology: anion / nitrate transport with the deterministic solver, without "
                                      "the groundwater boundary")
        if self.lbc and self.euler:
            raise NotImplementedError("settings.enable_groundwater_boundary runs with the deterministic solver")
        if self.nitrate and self.euler:
            raise NotImplementedError("nitrate runs with the deterministic solver")

        self.nx, self.ny, self.A = int(st.nx), int(st.ny), int(st.ages)
        self.C = self.nx * self.ny
        self.pull_state = bool(pull_state)
        self.stats = bool(getattr(st, "enable_age_statistics", False))
        # deuterium runs the same kernels with its own standard ratio and plausibility window (transport.py:315-340)
        vsmow, dmin, dmax = (st.VSMOW_conc2H, st.d2H_min, st.d2H_max) if iso2 else (st.VSMOW_conc18O, st.d18O_min, st.d18O_max)
        s = sas.SasSettings(self.A, int(st.sas_solver_substeps), float(vsmow), float(dmin), float(dmax), int(bool(store_tt)), int(self.stats),
                            2 if self.nitrate else (sas.VIRTUALTRACER if self.tracer else (sas.ANION if self.anion else sas.ISOTOPE)),
                            {"deterministic": sas.DETERMINISTIC, "Euler": sas.EULER, "RK4": sas.RK4}[solver])
        self.settings = s
        self.engine = engine_cls(self.C, s, **(engine_kwargs or {}))
        tau = int(vs.tau)
        mask = np.asarray(vs.maskCatch)[2:-2, 2:-2].reshape(-1).astype(np.float64)
        block = np.stack([np.asarray(getattr(vs, "sas_params_" + p))[2:-2, 2:-2, :].reshape(self.C, 8) for p in sas.PROCESSES])
        self.engine.set_parameters(mask, block.astype(np.float64))
        if self.anion:
            inner = lambda n: np.asarray(getattr(vs, n))[2:-2, 2:-2].reshape(-1).astype(np.float64)  # noqa: E731
            self.engine.set_anion_parameters(inner("alpha_transp"), inner("alpha_q"), inner("lu_id"), inner("S_sat_rz"))
        if self.crop:
            rows = [np.asarray(getattr(vs, "sas_params_" + p))[2:-2, 2:-2, :].reshape(self.C, 8).astype(np.float64) for p in ("re_rg", "re_rl")]
            self.engine.set_crop_parameters(rows[0], rows[1])
        for name in sas.STATE:
            self.engine.upload_state(name, np.asarray(getattr(vs, name))[2:-2, 2:-2, tau, :].reshape(self.C, self.A))
        if self.nitrate:
            from . import nitrate as nit

            self.nit = nit
            inner = lambda n: np.asarray(getattr(vs, n))[2:-2, 2:-2].reshape(-1).astype(np.float64)  # noqa: E731
            self.engine.set_nitrate_parameters(np.stack([inner(k) for k in nit.PARAMS]), float(st.dx), float(st.dy),
                                               float(st.zroot_to_zsoil_max), float(np.asarray(vs.dt)))
            for k, name in enumerate(("Nmin_rz", "Nmin_ss")):
                self.engine.upload_state_index(4 + k, np.asarray(getattr(vs, name))[2:-2, 2:-2, tau, :].reshape(self.C, self.A))
        if self.stats:
            self.engine.set_balance_reference(np.asarray(vs.C_s)[2:-2, 2:-2, int(vs.taum1)].reshape(-1))
            # storage-bound part of numerics.sanity_check (numerics.py:1051-1122) and its tolerances
            inner = lambda n: np.asarray(getattr(vs, n))[2:-2, 2:-2].reshape(-1).astype(np.float64)  # noqa: E731
            if all(hasattr(vs, n) for n in ("S_sat_rz", "S_sat_ss", "S_pwp_rz", "S_pwp_ss")):
                self.engine.set_sanity_bounds(inner("S_sat_rz"), inner("S_sat_ss"), inner("S_pwp_rz"), inner("S_pwp_ss"),
                                              float(getattr(st, "atol", 1e-2)), float(getattr(st, "rtol", 1e-2)))

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
        if self.euler:  # validate before touching vs: the device ages the StorAge once per call, i.e. the Euler steps must fill whole days
            n = int(self.settings.substeps)
            if (n * int(int(np.asarray(vs.dt_secs)) / n)) % (24 * 60 * 60) != 0:
                raise NotImplementedError("Euler steps that do not add up to whole days (the device ages the StorAge once per call)")
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
        f = np.stack([np.asarray(getattr(vs, n))[2:-2, 2:-2].reshape(-1).astype(np.float64) for n in self._FORCING])
        if self.nitrate:  # nitrate.calculate_nitrogen_cycle, first half (nitrate.py:717-731): the soil temperature, on the host
            nit, st = self.nit, state.settings
            inner = lambda n: np.asarray(getattr(vs, n))[2:-2, 2:-2].reshape(-1).astype(np.float64)  # noqa: E731
            tau = int(vs.tau)
            S_s = np.asarray(vs.S_s)[2:-2, 2:-2, tau].reshape(-1).astype(np.float64)
            temp = nit.soil_temperature(np.asarray(vs.TA), int(vs.itt), int(st.nitt), float(vs.doy[1]), inner("phi_soil_temp"), inner("damp_soil_temp"),
                                        inner("z_soil"), S_s, inner("S_sat_rz"), inner("S_sat_ss"), inner("maskCatch"), float(st.pi))
            z_root = np.asarray(vs.z_root)[2:-2, 2:-2, tau].reshape(-1).astype(np.float64)
            lu = np.asarray(vs.LU_ID)[2:-2, 2:-2, int(vs.itt)].reshape(-1).astype(np.float64)
            self.engine.set_nitrate_forcing(np.stack([temp, z_root, lu, inner("nup")]))
        if self.crop:  # the setup's set_forcing has put the day's redistribution fluxes into vs.re_rg, vs.re_rl
            self.engine.set_redistribution(np.asarray(vs.re_rg)[2:-2, 2:-2].reshape(-1), np.asarray(vs.re_rl)[2:-2, 2:-2].reshape(-1))
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
            if self.crop:
                ro = self.engine.download_redistribution()
                for k, name in enumerate(sas.REDISTRIBUTION):
                    if name in active:
                        self._put(name, ro[k])
            if self.stats:
                st = self.engine.download_statistics()
                for k, name in enumerate(sas.STATISTICS):
                    if self.lbc and name in ("dS_num_error", "dC_num_error", "sanity_ok"):
                        continue  # not computed by the reference under the groundwater boundary: the planes keep their initial 0
                    if name in active:
                        self._put(name, st[k])
                ok = st[sas.STATISTICS.index("sanity_ok")]
                if np.any(ok == 0.0) and not self.lbc:  # the reference's scalar is the AND over all cells (roger.py:453-456: a warning, not an error)
                    import warnings

                    warnings.warn(f"Solution diverged at iteration {int(np.asarray(vs.itt))} (numerics.sanity_check, {int((ok == 0.0).sum())} cells)")
            if self.nitrate:
                no = self.engine.download_nitrate_outputs()
                for k, name in enumerate(self.nit.OUTPUTS):
                    if name == "Nmin_s":
                        self._put(name, no[k], level=(tau, taum1))
                    elif name in active:
                        self._put(name, no[k])
                self._put("temp_soil", temp, level=(tau, taum1))
                if self.pull_state:
                    for k, name in enumerate(("Nmin_rz", "Nmin_ss")):
                        self._put(name, self.engine.download_state_index(4 + k), level=(tau, taum1))
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
