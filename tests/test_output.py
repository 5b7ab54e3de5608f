"""Diagnostic output files (roger_b200/output.py): layout rules of the reference's netCDF tools (io_tools/netcdf.py:23-142,
diagnostics/base.py:79-125) and the asynchronous hand-over, on the CPU with the oracle engine plugged in; the CUDA engine
writes the same bytes (tests/test_svat_gpu.py::test_async_diagnostic_output_file)."""
import numpy as np
import pytest

from roger_b200 import output, synth
from roger_b200.svat import SvatModel

SPEC = [("q_ss", "rate"), ("aet", "rate"), ("theta_rz", "average"), ("S_s", "maximum"), ("ta", "minimum"), ("S", "collect")]


def run_with_output(engine_cls, path, ndays=3, shape=(5, 4)):
    forcing = synth.mixed_forcing(ndays + 1, seed=11, start_doy=140)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=12, masked_fraction=0.15)
    m = SvatModel(engine_cls, shape, forcing, prim, th_rz, th_ss)
    m.setup()
    out = output.AsyncDiagnosticOutput(m.engine, path, shape, SPEC, identifier="SVAT", diagnostic="rate", time_origin="2022-12-31 00:00:00",
                                       mask=m.get("maskCatch"), force_overwrite=True)
    expected = []
    acc = None
    for day in range(ndays):
        steps = 0
        while True:
            sc = m.step(write_fluxes=True)
            vals = {n: m.get(n).reshape(-1).copy() for n, _ in SPEC}
            if acc is None:
                acc = {n: (np.zeros_like(v) if mode in ("rate", "average") else v.copy()) for (n, mode), v in zip(SPEC, vals.values())}
                fresh = True
            steps += 1
            for n, mode in SPEC:
                v = vals[n]
                if mode in ("rate", "average"):
                    acc[n] = acc[n] + v
                elif mode == "maximum" and not fresh:
                    acc[n] = np.maximum(acc[n], v)
                elif mode == "minimum" and not fresh:
                    acc[n] = np.minimum(acc[n], v)
                elif mode == "collect":
                    acc[n] = v
            fresh = False
            if sc.time % 86400 == 0:
                break
        acc["theta_rz"] = acc["theta_rz"] / steps
        expected.append((sc.time, {n: acc[n].copy() for n, _ in SPEC}))
        acc = None
        out.output(sc.time)
    out.close()
    return m, expected


def test_file_layout_and_values_follow_the_reference_rules(tmp_path):
    from oracle.engine import OracleSvatEngine
    from scipy.io import netcdf_file

    path = str(tmp_path / "SVAT.rate.nc")
    shape = (5, 4)
    m, expected = run_with_output(OracleSvatEngine, path, shape=shape)
    mask = m.get("maskCatch").reshape(shape).astype(bool)
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.dimensions["x"] == 5 and f.dimensions["y"] == 4 and f.dimensions["Time"] is None
        assert f.setup_identifier == b"SVAT" and b"t=0" in f.comment
        t = f.variables["Time"]
        assert t.units == b"days" and t.time_origin == b"2022-12-31 00:00:00"
        assert np.array_equal(t[:], [e[0] / 86400.0 for e in expected])
        for name, mode in SPEC:
            v = f.variables[name]
            assert v.dimensions == ("Time", "y", "x")  # transposed (netcdf.py:106)
            assert v._FillValue == -9999 and v.missing_value == -9999
            assert v.long_name.decode() == output.FIELD_META[name][0] and v.units.decode() == output.FIELD_META[name][1]
            data = v[:]
            assert data.shape == (len(expected), 4, 5)
            for k, (_, vals) in enumerate(expected):
                want = np.where(mask, vals[name].reshape(shape), -9999.0).T
                assert np.array_equal(data[k], want), (name, mode, k)


def test_existing_file_is_refused(tmp_path):
    path = str(tmp_path / "x.nc")
    output.DiagnosticFile(path, (2, 2), ["prec"]).close()
    with pytest.raises(IOError):
        output.DiagnosticFile(path, (2, 2), ["prec"])
    output.DiagnosticFile(path, (2, 2), ["prec"], force_overwrite=True).close()
