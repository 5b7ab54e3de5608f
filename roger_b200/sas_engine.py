"""ctypes binding of the StorAge-selection entry points of libroger_b200.so (include/roger_b200_sas.h): the `cuda`
engine of the offline oxygen-18 transport step.  No fallback: a missing library or CUDA device raises."""
import ctypes as C

import numpy as np

from . import sas
from .engine import RogerCudaError, _check, load_library

_bound = False


def _bind(lib):
    global _bound
    if _bound:
        return lib
    dp = C.POINTER(C.c_double)
    hp = C.c_void_p
    lib.roger_sas_create.argtypes = [C.POINTER(hp), C.c_int64, C.POINTER(sas.SasSettings), C.c_int]
    lib.roger_sas_destroy.argtypes = [hp]
    lib.roger_sas_destroy.restype = None
    lib.roger_sas_upload_state.argtypes = [hp, C.c_int, dp]
    lib.roger_sas_download_state.argtypes = [hp, C.c_int, dp]
    lib.roger_sas_state_ptr.argtypes = [hp, C.c_int]
    lib.roger_sas_state_ptr.restype = C.c_uint64
    lib.roger_sas_age_base.argtypes = [hp]
    lib.roger_sas_init_linear.argtypes = [hp, dp, dp, C.c_double]
    lib.roger_sas_set_parameters.argtypes = [hp, dp, dp]
    lib.roger_sas_step.argtypes = [hp, dp]
    lib.roger_sas_step_device.argtypes = [hp, C.c_uint64]
    lib.roger_sas_download_outputs.argtypes = [hp, dp]
    lib.roger_sas_download_tt.argtypes = [hp, C.c_int, dp]
    lib.roger_sas_download_statistics.argtypes = [hp, dp]
    lib.roger_sas_set_balance_reference.argtypes = [hp, dp]
    lib.roger_sas_set_sanity_bounds.argtypes = [hp, dp, dp, dp, dp, C.c_double, C.c_double]
    lib.roger_sas_set_anion_parameters.argtypes = [hp, dp, dp, dp, dp]
    lib.roger_sas_set_nitrate_parameters.argtypes = [hp, dp, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.roger_sas_set_nitrate_forcing.argtypes = [hp, dp]
    lib.roger_sas_download_nitrate_outputs.argtypes = [hp, dp]
    lib.roger_sas_set_crop_parameters.argtypes = [hp, dp, dp]
    lib.roger_sas_set_redistribution.argtypes = [hp, dp, dp]
    lib.roger_sas_download_redistribution.argtypes = [hp, dp]
    lib.roger_sas_synchronize.argtypes = [hp]
    lib.roger_sas_last_timing.argtypes = [hp, dp, C.POINTER(C.c_int64)]
    _bound = True
    return lib


class CudaSasEngine:
    def __init__(self, n_cells, settings, device=0):
        self.lib = _bind(load_library())
        self.n = int(n_cells)
        self.settings = settings
        self.A = int(settings.ages)
        self.h = C.c_void_p()
        _check(self.lib, self.lib.roger_sas_create(C.byref(self.h), self.n, C.byref(settings), int(device)))
        self.kernel_ms = 0.0
        self.launches = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.roger_sas_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def set_parameters(self, mask, sas_params):
        m = np.ascontiguousarray(mask, dtype=np.float64).reshape(-1)
        p = np.ascontiguousarray(sas_params, dtype=np.float64)
        if m.size != self.n or p.shape != (len(sas.PROCESSES), self.n, 8):
            raise ValueError("set_parameters: mask (n_cells,) and sas_params (5, n_cells, 8) expected")
        _check(self.lib, self.lib.roger_sas_set_parameters(self.h, self._dp(m), self._dp(p)))

    def upload_state(self, name, arr):
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).reshape(self.n, self.A))
        _check(self.lib, self.lib.roger_sas_upload_state(self.h, sas.STATE.index(name), self._dp(a)))

    def download_state(self, name):
        out = np.empty((self.n, self.A))
        _check(self.lib, self.lib.roger_sas_download_state(self.h, sas.STATE.index(name), self._dp(out)))
        return out

    def init_linear(self, S_rz, S_ss, conc0):
        a = np.ascontiguousarray(S_rz, dtype=np.float64).reshape(-1)
        b = np.ascontiguousarray(S_ss, dtype=np.float64).reshape(-1)
        if a.size != self.n or b.size != self.n:
            raise ValueError("init_linear: n_cells values expected")
        _check(self.lib, self.lib.roger_sas_init_linear(self.h, self._dp(a), self._dp(b), float(conc0)))

    def _timing(self):
        ms = C.c_double()
        n = C.c_int64()
        self.lib.roger_sas_last_timing(self.h, C.byref(ms), C.byref(n))
        self.kernel_ms += ms.value
        self.launches += n.value

    def step(self, forcing):
        f = np.ascontiguousarray(forcing, dtype=np.float64)
        if f.shape != (len(sas.FORCING), self.n):
            raise ValueError("step: forcing (9, n_cells) expected")
        _check(self.lib, self.lib.roger_sas_step(self.h, self._dp(f)))
        self._timing()

    def step_device(self, forcing_ptr):
        _check(self.lib, self.lib.roger_sas_step_device(self.h, int(forcing_ptr)))
        self._timing()

    def download_outputs(self, out=None):
        """(22, n_cells) result planes of the last step; pass a preallocated (ideally page-locked) array to avoid the
        staging copy of pageable memory"""
        if out is None:
            out = np.empty((len(sas.OUTPUTS), self.n))
        elif out.shape != (len(sas.OUTPUTS), self.n) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("download_outputs: contiguous float64 array of shape (22, n_cells) expected")
        _check(self.lib, self.lib.roger_sas_download_outputs(self.h, self._dp(out)))
        return out

    def download_statistics(self):
        out = np.empty((len(sas.STATISTICS), self.n))
        _check(self.lib, self.lib.roger_sas_download_statistics(self.h, self._dp(out)))
        return out

    def set_balance_reference(self, C_s_prev):
        a = np.ascontiguousarray(C_s_prev, dtype=np.float64).reshape(-1)
        if a.size != self.n:
            raise ValueError("set_balance_reference: n_cells values expected")
        _check(self.lib, self.lib.roger_sas_set_balance_reference(self.h, self._dp(a)))

    def set_sanity_bounds(self, S_sat_rz, S_sat_ss, S_pwp_rz, S_pwp_ss, atol=1e-2, rtol=1e-2):
        arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (S_sat_rz, S_sat_ss, S_pwp_rz, S_pwp_ss)]
        if any(a.size != self.n for a in arrs):
            raise ValueError("set_sanity_bounds: n_cells values expected")
        _check(self.lib, self.lib.roger_sas_set_sanity_bounds(self.h, *[self._dp(a) for a in arrs], float(atol), float(rtol)))

    def set_anion_parameters(self, alpha_transp, alpha_q, lu_id, S_sat_rz):
        arrs = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (alpha_transp, alpha_q, lu_id, S_sat_rz)]
        if any(a.size != self.n for a in arrs):
            raise ValueError("set_anion_parameters: n_cells values expected")
        _check(self.lib, self.lib.roger_sas_set_anion_parameters(self.h, *[self._dp(a) for a in arrs]))

    # ---- nitrate (settings.solute = 2): roger_b200.nitrate.NitrateModel drives these
    def set_nitrate_parameters(self, params, dx, dy, zroot_to_zsoil_max, dt_hours):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.shape != (13, self.n):
            raise ValueError("set_nitrate_parameters: (13, n_cells) planes expected")
        _check(self.lib, self.lib.roger_sas_set_nitrate_parameters(self.h, self._dp(p), float(dx), float(dy), float(zroot_to_zsoil_max), float(dt_hours)))

    def set_nitrate_forcing(self, planes):
        p = np.ascontiguousarray(planes, dtype=np.float64)
        if p.shape != (4, self.n):
            raise ValueError("set_nitrate_forcing: (4, n_cells) planes expected")
        _check(self.lib, self.lib.roger_sas_set_nitrate_forcing(self.h, self._dp(p)))

    def download_nitrate_outputs(self):
        out = np.empty((7, self.n))
        _check(self.lib, self.lib.roger_sas_download_nitrate_outputs(self.h, self._dp(out)))
        return out

    # ---- crop variant (svat_crop_transport_model_deterministic): redistribution after root growth / root loss opens the day
    def set_crop_parameters(self, sas_params_re_rg, sas_params_re_rl):
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(self.n, 8) for x in (sas_params_re_rg, sas_params_re_rl)]
        _check(self.lib, self.lib.roger_sas_set_crop_parameters(self.h, self._dp(a[0]), self._dp(a[1])))

    def set_redistribution(self, re_rg, re_rl):
        a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1) for x in (re_rg, re_rl)]
        if a[0].size != self.n or a[1].size != self.n:
            raise ValueError("set_redistribution: n_cells values each expected")
        _check(self.lib, self.lib.roger_sas_set_redistribution(self.h, self._dp(a[0]), self._dp(a[1])))

    def download_redistribution(self):
        out = np.empty((6, self.n))
        _check(self.lib, self.lib.roger_sas_download_redistribution(self.h, self._dp(out)))
        return out

    def upload_state_index(self, which, arr):
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).reshape(self.n, self.A))
        _check(self.lib, self.lib.roger_sas_upload_state(self.h, int(which), self._dp(a)))

    def download_state_index(self, which):
        out = np.empty((self.n, self.A))
        _check(self.lib, self.lib.roger_sas_download_state(self.h, int(which), self._dp(out)))
        return out

    def download_tt(self, process):
        out = np.empty((self.n, self.A))
        _check(self.lib, self.lib.roger_sas_download_tt(self.h, sas.PROCESSES.index(process), self._dp(out)))
        return out

    def synchronize(self):
        _check(self.lib, self.lib.roger_sas_synchronize(self.h))


__all__ = ["CudaSasEngine", "RogerCudaError"]
