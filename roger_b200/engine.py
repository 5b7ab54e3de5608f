"""ctypes binding of libroger_b200.so (include/roger_b200.h): the `cuda` engine of the SVAT / oneD path.

There is deliberately no fallback: if the shared library is missing or no CUDA device is usable,
construction raises.  PyTorch is not needed by this binding; when the caller passes torch tensors
they are used only as host/device buffers.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .build import LIB

_lib = None
FLAG_REDUCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_uint32), C.c_int, C.c_void_p)


class RogerCudaError(RuntimeError):
    pass


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise RogerCudaError(
            f"{LIB} not found: the CUDA extension has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "roger_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    hp = C.c_void_p
    lib.roger_last_error.restype = C.c_char_p
    lib.roger_svat_create.argtypes = [C.POINTER(hp), C.c_int64, C.POINTER(abi.SvatSettings), C.c_int]
    lib.roger_svat_destroy.argtypes = [hp]
    lib.roger_svat_destroy.restype = None
    lib.roger_svat_field_count.restype = C.c_int
    lib.roger_svat_field_name.argtypes = [C.c_int]
    lib.roger_svat_field_name.restype = C.c_char_p
    lib.roger_svat_field_kind.argtypes = [C.c_int]
    lib.roger_svat_upload.argtypes = [hp, C.c_int, dp]
    lib.roger_svat_download.argtypes = [hp, C.c_int, dp]
    lib.roger_svat_download_async.argtypes = [hp, C.c_int, dp]
    lib.roger_svat_download_diagnostics_async.argtypes = [hp, dp, C.c_int, C.POINTER(C.c_int64)]
    lib.roger_svat_download_fence.argtypes = [hp]
    lib.roger_svat_fence_wait.argtypes = [hp, C.c_int]
    lib.roger_svat_download_wait.argtypes = [hp]
    lib.roger_svat_set_d2h_mode.argtypes = [hp, C.c_int]
    lib.roger_svat_reserve_snapshots.argtypes = [hp, C.c_int]
    lib.roger_svat_tile_layout.argtypes = [hp, C.POINTER(abi.SvatLayout)]
    lib.roger_svat_field_slot.argtypes = [C.c_int]
    lib.roger_svat_optimize_cell_order.argtypes = [hp]
    lib.roger_svat_set_scalars.argtypes = [hp, C.POINTER(abi.SvatScalars)]
    lib.roger_svat_get_scalars.argtypes = [hp, C.POINTER(abi.SvatScalars)]
    lib.roger_svat_set_forcing_day.argtypes = [hp, dp, dp, dp]
    lib.roger_svat_set_forcing_day_device.argtypes = [hp, C.c_uint64, C.c_uint64, C.c_uint64]
    lib.roger_svat_step.argtypes = [hp, C.c_int]
    lib.roger_svat_run_day.argtypes = [hp, C.c_int, C.POINTER(C.c_int)]
    lib.roger_svat_run_steps.argtypes = [hp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.roger_svat_set_block_stepping.argtypes = [hp, C.c_int, C.c_int]
    lib.roger_svat_last_step_dts.argtypes = [hp, C.POINTER(C.c_int), C.c_int]
    lib.roger_svat_set_rate_fields.argtypes = [hp, C.POINTER(C.c_int), C.c_int]
    lib.roger_svat_set_diagnostics.argtypes = [hp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    lib.roger_svat_download_diagnostics.argtypes = [hp, dp, C.c_int]
    lib.roger_svat_download_rates.argtypes = [hp, dp, C.c_int]
    lib.roger_svat_rates_ptr.argtypes = [hp]
    lib.roger_svat_rates_ptr.restype = C.c_uint64
    lib.roger_svat_synchronize.argtypes = [hp]
    lib.roger_svat_checkpoint.argtypes = [hp]
    lib.roger_svat_rollback.argtypes = [hp]
    lib.roger_svat_last_timing.argtypes = [hp, dp, C.POINTER(C.c_int64)]
    lib.roger_svat_set_flag_reducer.argtypes = [hp, FLAG_REDUCE_FN, C.c_void_p]
    lib.roger_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    lib.roger_svat_enable_nccl_flags.argtypes = [hp, C.POINTER(C.c_uint8), C.c_int, C.c_int]
    # the field table compiled into the library must be the one this package parsed from the header
    n = lib.roger_svat_field_count()
    names = [lib.roger_svat_field_name(i).decode() for i in range(n)]
    if names != abi.FIELD_NAMES:
        raise RogerCudaError("libroger_b200.so was built from a different include/roger_b200_fields.h; rebuild")
    _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise RogerCudaError(f"roger_b200 error {rc}: {lib.roger_last_error().decode()}")


class CudaSvatEngine:
    def __init__(self, n_cells, settings, device=0, flag_reducer=None):
        self.lib = load_library()
        self.n = int(n_cells)
        self.settings = settings
        self.h = C.c_void_p()
        _check(self.lib, self.lib.roger_svat_create(C.byref(self.h), self.n, C.byref(settings), int(device)))
        self._reducer = None
        if flag_reducer is not None:
            # flag_reducer(list_of_uint32) -> list_of_uint32 OR-ed over all ranks (roger_b200.parallel)
            def _cb(ptr, n, _user):
                vals = flag_reducer([int(ptr[i]) for i in range(n)])
                for i in range(n):
                    ptr[i] = vals[i]

            self._reducer = FLAG_REDUCE_FN(_cb)
            _check(self.lib, self.lib.roger_svat_set_flag_reducer(self.h, self._reducer, None))
        self._keep = None
        self.kernel_ms = 0.0
        self.launches = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.roger_svat_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    def upload(self, name, arr):
        a = np.ascontiguousarray(np.asarray(arr, dtype=np.float64).reshape(-1))
        if a.size != self.n:
            raise ValueError(f"upload({name}): expected {self.n} values, got {a.size}")
        _check(self.lib, self.lib.roger_svat_upload(self.h, abi.FIELD_ID[name], self._dp(a)))

    def download(self, name):
        out = np.empty(self.n, dtype=np.float64)
        _check(self.lib, self.lib.roger_svat_download(self.h, abi.FIELD_ID[name], self._dp(out)))
        return out

    def download_async(self, name, host_pinned):
        """host_pinned: page-locked float64 array of n_cells values (e.g. a pinned torch tensor's .numpy()); valid after
        download_wait()."""
        if host_pinned.size != self.n or host_pinned.dtype != np.float64 or not host_pinned.flags["C_CONTIGUOUS"]:
            raise ValueError("download_async: contiguous float64 buffer of n_cells values expected")
        _check(self.lib, self.lib.roger_svat_download_async(self.h, abi.FIELD_ID[name], self._dp(host_pinned)))

    def download_wait(self):
        _check(self.lib, self.lib.roger_svat_download_wait(self.h))

    def set_d2h_mode(self, mode):
        """1: snapshots leave the device through an SM copy kernel into the mapped page-locked buffer (default), 0: copy engine"""
        _check(self.lib, self.lib.roger_svat_set_d2h_mode(self.h, int(mode)))

    def download_diagnostics_async(self, host_pinned, reset=True):
        """Hands the accumulators (len(spec), n_cells) to `host_pinned` off the critical path; returns the number of accumulated
        steps ("average" accumulators arrive as sums).  The buffer is valid after fence_wait(download_fence())."""
        n = len(self._rate_names)
        if host_pinned.size != n * self.n or host_pinned.dtype != np.float64 or not host_pinned.flags["C_CONTIGUOUS"]:
            raise ValueError("download_diagnostics_async: contiguous float64 buffer of (fields, n_cells) values expected")
        cnt = C.c_int64(0)
        _check(self.lib, self.lib.roger_svat_download_diagnostics_async(self.h, self._dp(host_pinned), int(bool(reset)), C.byref(cnt)))
        return int(cnt.value)

    def download_fence(self):
        t = int(self.lib.roger_svat_download_fence(self.h))
        if t < 0:
            _check(self.lib, t)
        return t

    def fence_wait(self, token):
        """may be called from another thread"""
        _check(self.lib, self.lib.roger_svat_fence_wait(self.h, int(token)))

    def reserve_snapshots(self, n):
        """device snapshot planes for n asynchronous plane hand-overs per output time, allocated now (never inside the stepping loop)"""
        _check(self.lib, self.lib.roger_svat_reserve_snapshots(self.h, int(n)))

    def enable_nccl_flags(self, dist, device=None):
        """Flag exchange of the adaptive time stepping by ncclAllGather on the engine's own stream.  `dist` is an initialised
        torch.distributed (any backend): it only carries the 128-byte NCCL id from rank 0 to the others."""
        import torch

        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            _check(self.lib, self.lib.roger_nccl_unique_id(buf))
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=0)
        buf = (C.c_uint8 * 128)(*[int(v) for v in t.cpu().tolist()])
        _check(self.lib, self.lib.roger_svat_enable_nccl_flags(self.h, buf, rank, world))

    def optimize_cell_order(self):
        """sort the cells on the device by infiltration capacity (warp coherence of the ponding-dependent loops); results and
        the cell numbering of every call are unchanged"""
        _check(self.lib, self.lib.roger_svat_optimize_cell_order(self.h))

    def tile_layout(self):
        lay = abi.SvatLayout()
        _check(self.lib, self.lib.roger_svat_tile_layout(self.h, C.byref(lay)))
        return lay

    def set_scalars(self, s):
        _check(self.lib, self.lib.roger_svat_set_scalars(self.h, C.byref(s)))

    def get_scalars(self):
        s = abi.SvatScalars()
        _check(self.lib, self.lib.roger_svat_get_scalars(self.h, C.byref(s)))
        return s

    def set_forcing_day(self, prec, ta, pet):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (prec, ta, pet)]
        want = abi.SLOTS_PER_DAY * (1 if self.settings.forcing_uniform else self.n)
        for a in arrs:
            if a.size != want:
                raise ValueError(f"forcing day block: expected {want} values, got {a.size}")
        self._keep = arrs  # the copy is asynchronous for per-cell forcing
        _check(self.lib, self.lib.roger_svat_set_forcing_day(self.h, *[self._dp(a) for a in arrs]))

    def set_forcing_day_device(self, prec_ptr, ta_ptr, pet_ptr):
        _check(self.lib, self.lib.roger_svat_set_forcing_day_device(self.h, int(prec_ptr), int(ta_ptr), int(pet_ptr)))

    def _timing(self):
        ms = C.c_double()
        n = C.c_int64()
        self.lib.roger_svat_last_timing(self.h, C.byref(ms), C.byref(n))
        self.kernel_ms += ms.value
        self.launches += n.value

    def step(self, write_fluxes=False):
        _check(self.lib, self.lib.roger_svat_step(self.h, int(bool(write_fluxes))))
        self._timing()

    def run_day(self, write_fluxes=False):
        n = C.c_int()
        _check(self.lib, self.lib.roger_svat_run_day(self.h, int(bool(write_fluxes)), C.byref(n)))
        self._timing()
        return n.value

    def run_steps(self, max_steps, write_fluxes=False):
        """Up to max_steps adaptive steps, never past the next midnight; returns the number taken."""
        n = C.c_int()
        _check(self.lib, self.lib.roger_svat_run_steps(self.h, int(bool(write_fluxes)), int(max_steps), C.byref(n)))
        self._timing()
        return n.value

    # "maximum_t" / "minimum_t": variables with a time-step dimension in the reference keep the extreme of the last two time
    # levels instead of a running one (roger/diagnostics/maximum.py:60-75)
    DIAG_MODES = {"rate": 0, "average": 1, "maximum": 2, "minimum": 3, "maximum_t": 4, "minimum_t": 5}

    def set_diagnostics(self, spec):
        """spec: list of (field name, "rate" | "average" | "maximum" | "minimum"); accumulated on the device after every step
        until download_rates(reset=True)."""
        self._rate_names = [n for n, _ in spec]
        n = len(spec)
        ids = (C.c_int * n)(*[abi.FIELD_ID[name] for name, _ in spec])
        modes = (C.c_int * n)(*[self.DIAG_MODES[m] for _, m in spec])
        _check(self.lib, self.lib.roger_svat_set_diagnostics(self.h, ids, modes, n))

    def set_rate_fields(self, names):
        """RATE diagnostics: sum the named planes over the steps until the next download_rates(reset=True)."""
        self.set_diagnostics([(n, "rate") for n in names])

    def download_rates(self, reset=True):
        out = np.empty((len(self._rate_names), self.n))
        _check(self.lib, self.lib.roger_svat_download_rates(self.h, self._dp(out), int(bool(reset))))
        return dict(zip(self._rate_names, out))

    def last_step_dts(self):
        """dt_secs of every step of the last step() / run_steps() call"""
        buf = (C.c_int * 144)()
        n = int(self.lib.roger_svat_last_step_dts(self.h, buf, 144))
        if n < 0:
            _check(self.lib, n)
        return [int(buf[k]) for k in range(min(n, 144))]

    def set_block_stepping(self, enabled, max_block=0):
        """enabled: False / 0 one launch and one flag read-back per step; True / 1 several steps in one launch (default up to
        262144 cells); 2 one launch per step queued back to back, one read-back per batch (default for larger grids)"""
        _check(self.lib, self.lib.roger_svat_set_block_stepping(self.h, int(enabled), int(max_block)))

    def checkpoint(self):
        """device-resident copy of every plane and of the scheduler state (the restart file of the reference, kept in HBM)"""
        _check(self.lib, self.lib.roger_svat_checkpoint(self.h))

    def rollback(self):
        _check(self.lib, self.lib.roger_svat_rollback(self.h))

    def synchronize(self):
        _check(self.lib, self.lib.roger_svat_synchronize(self.h))
