"""ctypes wrapper of the CPU oracle (oracle/svat_oracle.c).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libroger_oracle.so")


def build(force=False):
    deps = ("svat_oracle.c", "sas_oracle.c", "Makefile", os.path.join("..", "roger_b200", "csrc", "roger_pow.h"),
            os.path.join("..", "roger_b200", "csrc", "roger_gamma.h"),
            os.path.join("..", "roger_b200", "csrc", "roger_pow_tables.h"), os.path.join("..", "include", "roger_b200_sas.h"),
            os.path.join("..", "include", "roger_b200_fields.h"))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB) for f in deps):
        subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


class OracleSvatEngine:
    """Same engine interface as roger_b200.engine.CudaSvatEngine, computed by the C oracle."""

    def __init__(self, n_cells, settings, flag_reducer=None):
        from roger_b200 import abi

        self.abi = abi
        self.n = int(n_cells)
        self.settings = settings
        self._reducer = None
        if flag_reducer is not None:  # same contract as CudaSvatEngine(flag_reducer=...): one engine per process
            fn_t = C.CFUNCTYPE(None, C.POINTER(C.c_uint32), C.c_int, C.c_void_p)

            def _cb(ptr, n, _user):
                vals = flag_reducer([int(ptr[i]) for i in range(n)])
                for i in range(n):
                    ptr[i] = vals[i]

            self._reducer = fn_t(_cb)
            lib().svat_oracle_set_flag_reducer(self._reducer, None)
        self.planes = np.zeros((abi.N_FIELDS, self.n), dtype=np.float64)
        self.scalars = abi.initial_scalars()
        self.forcing = None
        self._fn = lib().svat_oracle_step
        self._fn.restype = None

    def upload(self, name, arr):
        self.planes[self.abi.FIELD_ID[name]] = np.asarray(arr, dtype=np.float64).reshape(-1)

    def download(self, name):
        return self.planes[self.abi.FIELD_ID[name]].copy()

    def set_scalars(self, s):
        self.scalars = s

    def get_scalars(self):
        return self.scalars

    def set_forcing_day(self, prec, ta, pet):
        self.forcing = [np.ascontiguousarray(a, dtype=np.float64) for a in (prec, ta, pet)]

    def checkpoint(self):
        """the in-memory checkpoint of roger_b200.engine.CudaSvatEngine (roger_svat_checkpoint)"""
        sc = self.abi.SvatScalars()
        C.memmove(C.byref(sc), C.byref(self.scalars), C.sizeof(sc))
        self._ckpt = (self.planes.copy(), sc, None if self.forcing is None else [a.copy() for a in self.forcing])

    def rollback(self):
        planes, sc, forcing = self._ckpt
        self.planes[:] = planes
        C.memmove(C.byref(self.scalars), C.byref(sc), C.sizeof(sc))
        self.forcing = None if forcing is None else [a.copy() for a in forcing]

    def step(self, write_fluxes=True):
        p, t, e = self.forcing
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        self._fn(C.c_int64(self.n), dp(self.planes), dp(p), dp(t), dp(e), C.byref(self.scalars), C.byref(self.settings))
        if getattr(self, "_rate_names", None):
            self._diagnose()

    def set_diagnostics(self, spec):
        """NumPy restatement of roger/diagnostics/{rate,average,maximum,minimum}.py for 2-D variables"""
        self._rate_names = [n for n, _ in spec]
        self._rate_modes = [m for _, m in spec]
        self._rates = np.zeros((len(spec), self.n))
        self._rates_prev = np.stack([self.planes[self.abi.FIELD_ID[n]].copy() for n, _ in spec]) if spec else np.zeros((0, self.n))
        self._diag_count = 0

    def set_rate_fields(self, names):
        self.set_diagnostics([(n, "rate") for n in names])

    def _diagnose(self):
        self._diag_count += 1
        for k, (name, mode) in enumerate(zip(self._rate_names, self._rate_modes)):
            v = self.planes[self.abi.FIELD_ID[name]]
            a = self._rates[k]
            with np.errstate(all="ignore"):
                if mode == "rate":
                    self._rates[k] = a + v
                elif mode == "average":
                    self._rates[k] = np.where(np.isnan(a), 0, a) + v
                elif mode in ("maximum_t", "minimum_t"):  # maximum.py:60-75: the extreme of the last two time levels
                    p = self._rates_prev[k].copy()
                    self._rates_prev[k] = v
                    mx = mode == "maximum_t"
                    d0 = np.where((p == 0) & ((v < 0) if mx else (v > 0)), np.nan, p)
                    both = np.maximum(d0, v) if mx else np.minimum(d0, v)
                    self._rates[k] = np.where(np.isnan(d0), v, np.where(np.isnan(v), d0, both))
                elif mode == "maximum":
                    d0 = np.where((a == 0) & (v < 0), np.nan, a)
                    self._rates[k] = np.where(np.isnan(d0), v, np.where(np.isnan(v), d0, np.maximum(d0, v)))
                else:
                    d0 = np.where((a == 0) & (v > 0), np.nan, a)
                    self._rates[k] = np.where(np.isnan(d0), v, np.where(np.isnan(v), d0, np.minimum(d0, v)))

    def download_rates(self, reset=True):
        vals = self._rates.copy()
        for k, mode in enumerate(self._rate_modes):
            if mode == "average" and self._diag_count > 0:
                vals[k] = vals[k] / self._diag_count
        out = dict(zip(self._rate_names, vals))
        if reset:
            self._rates[:] = 0
            self._diag_count = 0
        return out

    # the asynchronous hand-over of roger_b200.engine.CudaSvatEngine, executed synchronously
    def download_async(self, name, host_pinned):
        host_pinned[:] = self.download(name).reshape(-1)

    def download_wait(self):
        pass

    def download_diagnostics_async(self, host_pinned, reset=True):
        n = len(self._rate_names)
        host_pinned[: n * self.n] = self._rates.reshape(-1)
        count = self._diag_count
        if reset:
            self._rates[:] = 0
            self._diag_count = 0
        return count

    def download_fence(self):
        return 0

    def fence_wait(self, token):
        pass

    def synchronize(self):
        pass


class OracleSasEngine:
    """Same interface as roger_b200.sas_engine.CudaSasEngine, computed by oracle/sas_oracle.c."""

    def __init__(self, n_cells, settings):
        from roger_b200 import sas

        self.sas = sas
        self.n = int(n_cells)
        self.settings = settings
        self.A = int(settings.ages)
        self.state = {k: np.zeros((self.n, self.A)) for k in sas.STATE}
        self.out = np.zeros((len(sas.OUTPUTS), self.n))
        self.ttbuf = np.zeros((len(sas.PROCESSES), self.n, self.A)) if settings.store_tt else None
        self.stats = np.zeros((len(sas.STATISTICS), self.n)) if settings.age_statistics else None
        self.balance_prev = np.zeros(self.n)
        if settings.solver == 2 and settings.solute != 0:
            raise NotImplementedError("RK4 solver: isotope branch only")
        self._fn = lib().sas_oracle_day_stats
        self._fn.restype = None

    def set_parameters(self, mask, sas_params):
        self.mask = np.ascontiguousarray(mask, dtype=np.float64)
        self.params = np.ascontiguousarray(sas_params, dtype=np.float64)

    def upload_state(self, name, arr):
        self.state[name] = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).reshape(self.n, self.A)).copy()

    def download_state(self, name):
        return self.state[name].copy()

    def step(self, forcing):
        f = np.ascontiguousarray(forcing, dtype=np.float64)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        s = self.state
        self._install_bounds()
        self._fn(C.c_int64(self.n), C.byref(self.settings), dp(s["sa_rz"]), dp(s["sa_ss"]), dp(s["msa_rz"]), dp(s["msa_ss"]),
                 dp(f), dp(self.mask), dp(self.params), dp(self.out), dp(self.ttbuf) if self.ttbuf is not None else None,
                 dp(self.stats) if self.stats is not None else None, dp(self.balance_prev))

    def set_balance_reference(self, C_s_prev):
        self.balance_prev[:] = np.asarray(C_s_prev, dtype=np.float64).reshape(-1)

    def set_sanity_bounds(self, S_sat_rz, S_sat_ss, S_pwp_rz, S_pwp_ss, atol=1e-2, rtol=1e-2):
        self._bounds = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (S_sat_rz, S_sat_ss, S_pwp_rz, S_pwp_ss)]
        self._tol = (float(atol), float(rtol))

    def set_anion_parameters(self, alpha_transp, alpha_q, lu_id, S_sat_rz):
        self._anion = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (alpha_transp, alpha_q, lu_id, S_sat_rz)]

    # ---- nitrate (settings.solute = 2): roger_b200.nitrate.NitrateModel drives these
    def set_nitrate_parameters(self, params, dx, dy, zroot_to_zsoil_max, dt_hours):
        self._nit_params = np.ascontiguousarray(params, dtype=np.float64)
        self._nit_scal = (float(dx), float(dy), float(zroot_to_zsoil_max), float(dt_hours))
        self._nit_out = np.zeros((7, self.n))
        self.state.setdefault("Nmin_rz", np.zeros((self.n, self.A)))
        self.state.setdefault("Nmin_ss", np.zeros((self.n, self.A)))

    def set_nitrate_forcing(self, planes):
        self._nit_forcing = np.ascontiguousarray(planes, dtype=np.float64)

    def download_nitrate_outputs(self):
        return self._nit_out.copy()

    # ---- crop variant (redistribution after root growth / root loss opens the day)
    def set_crop_parameters(self, sas_params_re_rg, sas_params_re_rl):
        if self.settings.solute == 0 or self.settings.solver != 0:
            raise NotImplementedError("crop variant: anion / nitrate solutes, deterministic solver")
        self._crop_sp = [np.ascontiguousarray(a, dtype=np.float64).reshape(self.n, 8) for a in (sas_params_re_rg, sas_params_re_rl)]
        self._crop_out = np.zeros((6, self.n))
        self._crop_re = None

    def set_redistribution(self, re_rg, re_rl):
        self._crop_re = np.ascontiguousarray(np.stack([np.asarray(re_rg, dtype=np.float64).reshape(-1), np.asarray(re_rl, dtype=np.float64).reshape(-1)]))

    def download_redistribution(self):
        return self._crop_out.copy()

    def upload_state_index(self, which, arr):
        name = (self.sas.STATE + ["Nmin_rz", "Nmin_ss"])[which]
        self.state[name] = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).reshape(self.n, self.A)).copy()

    def download_state_index(self, which):
        return self.state[(self.sas.STATE + ["Nmin_rz", "Nmin_ss"])[which]].copy()

    def _install_bounds(self):
        """the C oracle keeps the bounds in file-scope variables: (re)install this engine's before every step"""
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        fn = lib().sas_oracle_set_sanity_bounds
        fn.argtypes = [C.POINTER(C.c_double)] * 4 + [C.c_double, C.c_double]
        fn.restype = None
        b = getattr(self, "_bounds", None)
        atol, rtol = getattr(self, "_tol", (1e-2, 1e-2))
        fn(*([dp(a) for a in b] if b else [None] * 4), atol, rtol)
        fa = lib().sas_oracle_set_anion_parameters
        fa.argtypes = [C.POINTER(C.c_double)] * 4
        fa.restype = None
        a = getattr(self, "_anion", None)
        fa(*([dp(x) for x in a] if a else [None] * 4))
        fn2 = lib().sas_oracle_set_nitrate
        fn2.argtypes = [C.POINTER(C.c_double)] * 5 + [C.c_double] * 4
        fn2.restype = None
        if self.settings.solute == 2:
            fn2(dp(self._nit_params), dp(self._nit_forcing), dp(self.state["Nmin_rz"]), dp(self.state["Nmin_ss"]), dp(self._nit_out), *self._nit_scal)
        else:
            fn2(None, None, None, None, None, 1.0, 1.0, 0.7, 24.0)

        fc = lib().sas_oracle_set_crop
        fc.argtypes = [C.POINTER(C.c_double)] * 4
        fc.restype = None
        if getattr(self, "_crop_sp", None) is not None:
            if self._crop_re is None:
                raise RuntimeError("crop variant: set_redistribution has not been called for this day")
            fc(dp(self._crop_sp[0]), dp(self._crop_sp[1]), dp(self._crop_re), dp(self._crop_out))
            self._crop_re_used, self._crop_re = self._crop_re, None  # keep the buffer alive through the step
        else:
            fc(None, None, None, None)

    def download_statistics(self):
        return self.stats.copy()

    def download_outputs(self):
        return self.out.copy()

    def download_tt(self, process):
        return self.ttbuf[self.sas.PROCESSES.index(process)].copy()

    def synchronize(self):
        pass
