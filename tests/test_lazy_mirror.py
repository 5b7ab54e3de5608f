"""Dirty-flag mirrors of the drop-in driver (roger/state.py:213-251 extended on the instance): host assignments are uploaded
before the next step without being announced, device results are downloaded on first read (mirror="lazy") -- checked here on
the CPU oracle engine with the stand-in setup; tests/test_svat_gpu.py repeats it on the CUDA engine and
tests/test_cuda_backend_patch.py through the patched reference."""
import numpy as np
import pytest

from oracle.engine import OracleSvatEngine
from roger_b200 import synth
from roger_b200.integration import CudaStepDriver
from tests import util
from tests.mini_setup import MiniSvatSetup


def _drivers(**kw):
    shape = (5, 4)
    forcing = synth.mixed_forcing(40, seed=42, start_doy=115)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, masked_fraction=0.1)
    out = []
    for mode in ("eager", "lazy"):
        setup = MiniSvatSetup(shape, forcing, prim, th_rz, th_ss)
        setup.setup()
        out.append(CudaStepDriver(setup, engine_cls=OracleSvatEngine, mirror=mode, **kw))
    return out


def _interior(vs, n):
    a = np.asarray(getattr(vs, n))
    return a[2:-2, 2:-2, 1] if a.ndim == 3 else a[2:-2, 2:-2]


def test_lazy_equals_eager_and_downloads_on_read_only():
    eager, lazy = _drivers(fluxes=["aet"])
    for step in range(260):
        if step == 90:
            for d in (eager, lazy):
                a = np.array(d.state.variables.z0)
                a[2:-2, 2:-2, 1] += 3.0  # ponded water appears from nowhere: a host-side edit of a state array
                d.state.variables.z0 = a
        a, b = eager.step(), lazy.step()
        assert (a.time, a.dt_secs, a.itt_day, a.event_id[1]) == (b.time, b.dt_secs, b.itt_day, b.event_id[1])
        if step in (0, 91, 130, 259):
            for n in ("S", "theta_rz", "z0", "aet", "q_hof" if False else "S_fp_ss"):
                assert util.same_bits(_interior(eager.state.variables, n), _interior(lazy.state.variables, n)), (step, n)
    assert lazy.downloads <= 4 * 5 + 4  # five variables at four steps (+ vs.ta at the month changes)
    assert eager.downloads > 100 * lazy.downloads
    assert lazy.uploads == eager.uploads  # the z0 edit and the monthly land-use refreshes, nothing else


def test_both_levels_of_a_materialised_mirror_show_the_result():
    _, lazy = _drivers()
    for _ in range(5):
        lazy.step()
    s = np.asarray(lazy.state.variables.S)
    assert util.same_bits(s[2:-2, 2:-2, 0], s[2:-2, 2:-2, 1])
    assert not s.flags.writeable


def test_unrequested_flux_is_refused_not_served_stale():
    _, lazy = _drivers()
    lazy.step()
    with pytest.raises(RuntimeError, match="fluxes"):
        lazy.state.variables.q_ss
    _ = lazy.state.variables.S  # state variables are always available
