"""Pins the CPU oracle (oracle/svat_oracle.c) against the committed outputs of the reference's own
NumPy backend (tests/golden/*.npz, produced by tests/golden/make_svat_golden.py)."""
import pytest

from oracle.engine import OracleSvatEngine
from tests import util

# glibc pow/exp/log are the same libm NumPy calls: the oracle is expected to agree to rounding;
# the bound below is the reference's own restart-parity tolerance (test/restart_test.py:7-17).
TOL = 1e-10


@pytest.mark.parametrize("case", util.SVAT_CASES + util.ONED_CASES)
def test_oracle_matches_reference(case):
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, OracleSvatEngine, from_reference_setup=True)
    util.replay_and_compare(fx, model, TOL)
