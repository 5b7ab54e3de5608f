"""The drop-in driver refuses setups whose settings switch on branches of RogerSetup.step the CUDA engine does not
implement (roger/settings.py:86-113), instead of running them silently without those branches."""
import pytest

from oracle.engine import OracleSvatEngine
from roger_b200 import integration, synth
from tests.mini_setup import MiniSvatSetup


def _setup(**settings):
    shape = (2, 2)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=3)
    m = MiniSvatSetup(shape, synth.toy_forcing("rain", ndays=1, seed=1), prim, th_rz, th_ss)
    for k, v in settings.items():
        setattr(m.state.settings, k, v)
    return m


@pytest.mark.parametrize("flag", integration.UNSUPPORTED_SETTINGS)
def test_unsupported_setting_is_refused(flag):
    with pytest.raises(NotImplementedError, match=flag):
        integration.CudaStepDriver(_setup(**{flag: True}), engine_cls=OracleSvatEngine)


def test_the_real_routing_setting_names_are_guarded():
    # the reference has no `enable_routing`; its switches are enable_routing_1D / enable_routing_2D (settings.py:108-109)
    assert "enable_routing_1D" in integration.UNSUPPORTED_SETTINGS and "enable_routing_2D" in integration.UNSUPPORTED_SETTINGS
    assert "enable_routing" not in integration.UNSUPPORTED_SETTINGS


def test_fixed_time_stepping_is_refused():
    with pytest.raises(NotImplementedError, match="enable_adaptive_time_stepping"):
        integration.CudaStepDriver(_setup(enable_adaptive_time_stepping=False), engine_cls=OracleSvatEngine)


def test_offline_transport_setups_are_sent_to_the_transport_driver():
    with pytest.raises(ValueError, match="offline transport"):
        integration.CudaStepDriver(_setup(enable_offline_transport=True), engine_cls=OracleSvatEngine)


def test_supported_setup_is_accepted():
    integration.CudaStepDriver(_setup(), engine_cls=OracleSvatEngine)


def test_uniform_forcing_detection_handles_nan_slots_and_signed_zeros():
    """day blocks of (nx, ny, 144): one series shared by all cells -- with gaps (NaN slots, which nanmean tolerates) and a -0.0 written
    where another cell holds +0.0 -- counts as uniform; a single deviating cell does not"""
    import numpy as np

    from roger_b200.integration import CudaStepDriver

    rng = np.random.default_rng(0)
    full = np.empty((70, 5, 144))
    full[:] = rng.random(144)
    full[:, :, 7] = np.nan
    a = full[2:-2, 1:-1, :]  # a strided interior view, as the driver sees it
    assert CudaStepDriver._is_uniform(a)
    full[40, 2, 9] = -0.0
    full[:, :, 9] = np.where(np.arange(70)[:, None] == 40, -0.0, 0.0)
    assert CudaStepDriver._is_uniform(a)
    full[66, 3, 100] += 1e-9
    assert not CudaStepDriver._is_uniform(a)
    assert CudaStepDriver._is_uniform(full[2:-2, 1:-1, :].astype(np.float32)[:, :, :99])  # other dtypes take the value comparison
