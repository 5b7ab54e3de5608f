"""Pins the CPU oracle (oracle/svat_oracle.c) against the committed outputs of the reference's own NumPy backend
(tests/golden/svat_*.npz / oned_*.npz, produced by tests/golden/make_svat_golden.py from the unmodified reference).

The fixtures are generated with NumPy's float64 `power` routed to the C library's pow (NumPy's SIMD dispatch off), the
function the oracle calls: every per-cell field of every recorded step and every scheduler scalar then agrees BIT FOR BIT."""
import pytest

from oracle.engine import OracleSvatEngine, lib
from tests import util


@pytest.fixture(autouse=True)
def _libm_pow():
    lib().roger_oracle_set_pow(0)
    yield
    lib().roger_oracle_set_pow(0)


@pytest.mark.parametrize("case", util.SVAT_CASES + util.ONED_CASES)
def test_oracle_reproduces_reference_bit_for_bit(case):
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, OracleSvatEngine, from_reference_setup=True)
    util.replay_and_compare(fx, model, 0.0, exact=True)


@pytest.mark.parametrize("case", util.SVAT_CASES + util.ONED_CASES)
def test_oracle_with_shared_pow_reproduces_reference_bit_for_bit(case):
    """roger_pow (the operation sequence the CUDA kernel evaluates) returns the C library's bits (tests/test_pow.py), so this mode
    reproduces the fixtures bit for bit as well -- and with it the kernel, which equals this mode of the oracle bit for bit
    (tests/test_svat_gpu.py)"""
    lib().roger_oracle_set_pow(1)
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, OracleSvatEngine, from_reference_setup=True)
    util.replay_and_compare(fx, model, 0.0, exact=True)
