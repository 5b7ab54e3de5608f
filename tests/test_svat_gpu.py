"""GPU parity: the fused sm_100a SVAT / oneD step (through the C ABI) against (a) the committed reference
fixtures and (b) the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

# float64 state within 1e-10 relative (north star); the normalised-max idiom of test/restart_test.py:7-17
TOL = 1e-10


def _engine():
    from roger_b200.engine import CudaSvatEngine

    return CudaSvatEngine


@pytest.mark.parametrize("case", util.SVAT_CASES + util.ONED_CASES)
def test_cuda_matches_reference_fixture(case):
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, _engine(), from_reference_setup=True)
    worst = util.replay_and_compare(fx, model, TOL, write_fluxes=True)
    print(case, sorted(worst.items(), key=lambda kv: -kv[1][0])[:4])


@pytest.mark.parametrize("case", ["svat_rain", "svat_mixed_dist", "oned_mixed"])
def test_cuda_state_only_mode_matches(case):
    """write_fluxes=0 (the production mode) must give the same state as the diagnostics mode."""
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, _engine(), from_reference_setup=True)
    util.replay_and_compare(fx, model, TOL, kinds=("S",), write_fluxes=False)


def test_async_download_is_a_snapshot():
    """roger_svat_download_async returns the plane as of the call, even when more steps are issued before the wait."""
    import torch

    fx = util.load_case("svat_mixed_dist")
    model = util.model_from_fixture(fx, _engine(), from_reference_setup=True)
    eng = model.engine
    for _ in range(5):
        model.step(False)
    want = {n: eng.download(n) for n in ("theta_rz", "S")}
    pins = {n: torch.empty(eng.n, dtype=torch.float64).pin_memory().numpy() for n in want}
    for n in want:
        eng.download_async(n, pins[n])
    for _ in range(30):
        model.step(False)
    eng.download_wait()
    for n in want:
        assert np.array_equal(pins[n], want[n])
    # the state has moved on in the meantime
    assert any(not np.array_equal(eng.download(n), want[n]) for n in want)
