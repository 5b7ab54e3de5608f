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
