"""Diagnostic output files of the cuda backend: the step either side of the hot path.

The reference accumulates its diagnostics on the host after every step (roger/diagnostics/{rate,average,maximum,minimum}.py)
and hands the file write to a thread (`threaded_io`, roger/io_tools/netcdf.py:145-210).  Here the accumulators live on the
device (`roger_svat_set_diagnostics`), an output is a snapshot + asynchronous device-to-host copy on a second stream
(`roger_svat_download_diagnostics_async`), and a writer thread waits for that copy (`roger_svat_fence_wait`) and appends the
record to the file while the next steps run.

File layout = what `initialize_file` / `initialize_variable` / `write_variable` / `advance_time` produce
(io_tools/netcdf.py:23-142, diagnostics/base.py:79-125): dimensions x, y (interior cells) and an unlimited Time; coordinate
variables x, y, Time (days since `time_origin`); every output variable with transposed dimensions (Time, y, x), `_FillValue` /
`missing_value` -9999 outside the catchment mask, `long_name` and `units` of the reference's variable table; file attributes
date_created, roger_version, comment, setup_identifier.  Container: the reference writes netCDF-4 / HDF5 through h5netcdf, which
this image does not have; this writer produces NetCDF-3 (64-bit offset) through `scipy.io.netcdf_file`, which the netCDF
library, xarray and the reference's own post-processing open the same way.
"""
import datetime
import os
import queue
import threading

import numpy as np

from .field_meta import FIELD_META

FLOAT_FILL_VALUE = -9999  # roger/variables.py:57
MODES = ("rate", "average", "maximum", "minimum", "maximum_t", "minimum_t", "collect")


class DiagnosticFile:
    """One output file `{identifier}.{diagnostic}.nc` (diagnostics/base.py:66-77)."""

    def __init__(self, path, shape, variables, identifier="roger", diagnostic="rate", time_origin="1900-01-01 00:00:00", x=None, y=None,
                 mask=None, force_overwrite=False, version="cuda"):
        from scipy.io import netcdf_file

        if os.path.isfile(path) and not force_overwrite:  # diagnostics/base.py:89-95
            raise IOError(f'output file {path} for diagnostic "{diagnostic}" exists (change the output path, pass force_overwrite=True '
                          "or delete existing files in the output directory)")
        self.path = path
        self.nx, self.ny = int(shape[0]), int(shape[1])
        self.variables = list(variables)
        self.mask = None if mask is None else np.asarray(mask).reshape(self.nx, self.ny).astype(bool)
        f = netcdf_file(path, "w", version=2)
        f.date_created = datetime.datetime.today().isoformat()
        f.roger_version = str(version)
        f.comment = ("First timestep (t=0) contains initial values. Simulations start are written from second timestep (t=1) to "
                     "last timestep (t=N).")
        f.setup_identifier = str(identifier)
        f.createDimension("Time", None)
        f.createDimension("x", self.nx)
        f.createDimension("y", self.ny)
        for dim, n, vals in (("x", self.nx, x), ("y", self.ny, y)):
            v = f.createVariable(dim, "d", (dim,))
            v[:] = np.arange(1, n + 1, dtype=np.float64) if vals is None else np.asarray(vals, dtype=np.float64)
            v.long_name = dim
            v.units = "m"
        t = f.createVariable("Time", "d", ("Time",))
        t.long_name = "Time"
        t.units = "days"
        t.time_origin = str(time_origin)
        for name in self.variables:
            v = f.createVariable(name, "d", ("Time", "y", "x"))  # dims transposed (netcdf.py:106-107)
            long_name, units = FIELD_META.get(name, (name, ""))
            v.long_name = long_name
            v.units = units
            v._FillValue = np.float64(FLOAT_FILL_VALUE)
            v.missing_value = np.float64(FLOAT_FILL_VALUE)
        self._f = f
        self.n_records = 0

    def append(self, time_seconds, planes):
        """planes: (len(variables), nx * ny) values of this output time"""
        f = self._f
        k = self.n_records
        f.variables["Time"][k] = time_seconds / 86400.0  # time.convert_time(vs.time, "seconds", "days")
        planes = np.asarray(planes, dtype=np.float64).reshape(len(self.variables), self.nx, self.ny)
        for j, name in enumerate(self.variables):
            data = planes[j]
            if self.mask is not None:
                data = np.where(self.mask, data, np.float64(FLOAT_FILL_VALUE))  # netcdf.py:121-124
            f.variables[name][k] = data.T  # netcdf.py:128
        self.n_records += 1
        f.flush()

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None


def _pinned(n):
    """page-locked float64 buffer (torch when present: asynchronous copies need it; plain NumPy otherwise)"""
    try:
        import torch

        if torch.cuda.is_available():
            t = torch.empty(n, dtype=torch.float64).pin_memory()
            return t.numpy(), t
    except Exception:  # noqa: BLE001 -- a missing / CPU-only torch only costs the overlap
        pass
    return np.empty(n), None


class AsyncDiagnosticOutput:
    """Device accumulators -> file, off the critical path.

        out = AsyncDiagnosticOutput(model.engine, "SVAT.rate.nc", shape, [("q_ss", "rate"), ("theta_rz", "average")], ...)
        for day in range(ndays):
            model.run_steps(144)
            out.output(model.engine.get_scalars().time)   # returns immediately; the record is written by the writer thread
        out.close()

    `spec` modes: "rate" (sum over the steps, rate.py:72-81), "average" (average.py:71-97), "maximum" / "minimum"
    (maximum.py:76-92), "collect" (the value at the output time, collect.py:61-71 -- a plain asynchronous plane download).
    """

    def __init__(self, engine, path, shape, spec, depth=2, files=None, **file_kwargs):
        """`files` (optional): {path: [variable names]} splits the variables of `spec` over several files, e.g. one per
        diagnostic of the reference ("{identifier}.rate.nc", "{identifier}.average.nc", ...); default: all of them in `path`."""
        self.engine = engine
        self.spec = [(n, m) for n, m in spec]
        for _, m in self.spec:
            if m not in MODES:
                raise ValueError(f"unknown diagnostic mode {m!r}")
        self.acc = [(n, m) for n, m in self.spec if m != "collect"]
        self.col = [n for n, m in self.spec if m == "collect"]
        names = [n for n, _ in self.acc] + self.col
        self.C = int(shape[0]) * int(shape[1])
        if files is None:
            files = {path: names}
        self.files = []  # (DiagnosticFile, rows of the record buffer it takes)
        for fpath, fvars in files.items():
            rows = [names.index(v) for v in fvars]
            kw = dict(file_kwargs)
            if "diagnostic" not in kw:
                kw["diagnostic"] = os.path.basename(fpath).split(".")[-2] if fpath.count(".") >= 2 else "rate"
            self.files.append((DiagnosticFile(fpath, shape, list(fvars), **kw), rows))
        self.file = self.files[0][0]
        if self.acc:
            engine.set_diagnostics(self.acc)
        self._buffers = [_pinned(len(names) * self.C) for _ in range(max(1, int(depth)))]
        if hasattr(engine, "reserve_snapshots"):
            engine.reserve_snapshots(len(names))  # the device side of every hand-over exists before the first step
        self._free = queue.Queue()
        for b in self._buffers:
            self._free.put(b)
        self._work = queue.Queue()
        self._error = None
        self._thread = threading.Thread(target=self._writer, name="roger-output", daemon=True)
        self._thread.start()

    def output(self, time_seconds):
        """Snapshot the accumulators (and reset them), start the copies, queue the record for the writer thread."""
        if self._error is not None:
            raise self._error
        buf = self._free.get()  # blocks only when `depth` records are still in flight
        arr = buf[0]
        na = len(self.acc)
        self.engine.download_wait()  # recycle the snapshot planes of records that have already been copied
        count = self.engine.download_diagnostics_async(arr[: na * self.C], reset=True) if na else 0
        for j, name in enumerate(self.col):
            self.engine.download_async(name, arr[(na + j) * self.C: (na + j + 1) * self.C])
        token = self.engine.download_fence()
        self._work.put((float(time_seconds), buf, count, token))

    def _writer(self):
        while True:
            item = self._work.get()
            if item is None:
                return
            time_seconds, buf, count, token = item
            try:
                self.engine.fence_wait(token)
                planes = buf[0].reshape(len(self.acc) + len(self.col), self.C)
                for j, (_, mode) in enumerate(self.acc):
                    if mode == "average" and count > 0:  # average.py:93-97
                        planes[j] = planes[j] / count
                for f, rows in self.files:
                    f.append(time_seconds, planes[rows])
            except BaseException as e:  # noqa: BLE001 -- surfaced by the next output() / close()
                self._error = e
            finally:
                self._free.put(buf)

    def close(self):
        self._work.put(None)
        self._thread.join()
        self.engine.download_wait()
        for f, _ in self.files:
            f.close()
        if self._error is not None:
            raise self._error


class ThreadedRecordWriter:
    """Host-side records -> files on a writer thread (`threaded_io`, io_tools/netcdf.py:145-210) for engines whose results are
    already on the host after a step (the daily offline-transport step).  files: {path: [variable names]}."""

    def __init__(self, shape, files, **file_kwargs):
        self.files = []
        for fpath, fvars in files.items():
            kw = dict(file_kwargs)
            if "diagnostic" not in kw:
                kw["diagnostic"] = os.path.basename(fpath).split(".")[-2] if fpath.count(".") >= 2 else "collect"
            self.files.append(DiagnosticFile(fpath, shape, list(fvars), **kw))
        self._work = queue.Queue(maxsize=4)
        self._error = None
        self._thread = threading.Thread(target=self._writer, name="roger-output", daemon=True)
        self._thread.start()

    def output(self, time_seconds, values):
        """values: {variable name: array of nx * ny values}; copied, so the caller may reuse its arrays"""
        if self._error is not None:
            raise self._error
        self._work.put((float(time_seconds), {k: np.array(v, dtype=np.float64) for k, v in values.items()}))

    def _writer(self):
        while True:
            item = self._work.get()
            if item is None:
                return
            t, values = item
            try:
                for f in self.files:
                    f.append(t, np.stack([values[v].reshape(-1) for v in f.variables]))
            except BaseException as e:  # noqa: BLE001 -- surfaced by the next output() / close()
                self._error = e

    def close(self):
        self._work.put(None)
        self._thread.join()
        for f in self.files:
            f.close()
        if self._error is not None:
            raise self._error

