"""Pins the CPU oracle of the StorAge-selection step (oracle/sas_oracle.c) against outputs of the reference's own NumPy
backend (tests/golden/sas_*.npz, produced by tests/golden/make_sas_golden.py).

With the C library's pow on both sides the oracle reproduces the reference BIT FOR BIT -- age distributions, travel-time
distributions, concentrations, delta values, NaN patterns, age-bin bookkeeping.  The remaining tests document the
reference's own sensitivity to the libm ("SAS parity chain")."""
import numpy as np
import pytest

from oracle.engine import OracleSasEngine, lib
from tests import util_sas as u


@pytest.fixture(autouse=True)
def _libm_pow():
    lib().sas_oracle_set_pow(0)
    yield
    lib().sas_oracle_set_pow(0)


@pytest.mark.parametrize("case", u.SAS_CASES + u.ANION_CASES + u.EULER_CASES + u.RK4_CASES)
def test_oracle_reproduces_reference_bit_for_bit(case):
    fx = u.load_case(case)
    m = u.model_from_fixture(fx, OracleSasEngine)
    u.replay_against_fixture(fx, m, exact=True)


@pytest.mark.parametrize("case", u.SAS_CASES + u.ANION_CASES + u.EULER_CASES + u.RK4_CASES)
def test_oracle_with_the_shared_pow_reproduces_reference_bit_for_bit(case):
    """the same with roger_pow / roger_exp (the operation sequence the CUDA kernels evaluate) in place of the C library's
    functions: identical, because roger_pow.h returns the library's bits (tests/test_pow.py) -- the device results, which equal
    this oracle bit for bit (tests/test_sas_gpu.py), therefore equal the reference fixtures bit for bit as well"""
    lib().sas_oracle_set_pow(1)
    fx = u.load_case(case)
    m = u.model_from_fixture(fx, OracleSasEngine)
    u.replay_against_fixture(fx, m, exact=True)


@pytest.mark.parametrize("case", u.GAMMA_CASES)
def test_oracle_reproduces_reference_with_the_gamma_family(case):
    """gamma (id 4) and backwards-read exponential (id 52) SAS functions, deterministic solver: the incomplete gamma function is a
    restatement of the published series / continued fraction, not of scipy's code -- 1e-10 (measured 3e-14), delta values 1e-9"""
    fx = u.load_case(case)
    m = u.model_from_fixture(fx, OracleSasEngine)
    worst = u.replay_against_fixture(fx, m, exact=False)
    bad = {k: v for k, v in worst.items() if v[0] > (1e-9 if k.startswith("C_iso") else 1e-10)}
    assert not bad, bad
    assert float(np.abs(fx["rec_tt_transp"]).max()) > 0  # the gamma-controlled flux withdraws water


def test_groundwater_boundary_variant_is_the_same_transport():
    """settings.enable_groundwater_boundary: svat_lbc_transport_model_deterministic (transport.py:995-1040) adds the capillary rise into
    the subsoil, whose transport kernels are empty stubs (capillary_rise.py:590-611) -- bit for bit the plain model; the reference
    then computes no balance errors (numerics.py:423-428: the planes stay 0)"""
    fx = u.load_case("sas_power_lbc")
    assert float(np.abs(fx["rec_dS_num_error"]).max()) == 0.0 and float(np.abs(fx["rec_dC_num_error"]).max()) == 0.0
    for k in ("rec_dS_num_error", "rec_dC_num_error"):
        fx.pop(k)
    m = u.model_from_fixture(fx, OracleSasEngine)
    u.replay_against_fixture(fx, m, exact=True)


def test_host_side_setup_matches_reference_initial_state():
    """roger_b200.sas.SasModel.setup (initial age distributions) against the reference's set_initial_conditions."""
    fx = u.load_case("sas_power")
    C, A = int(fx["nx"]) * int(fx["ny"]), int(fx["ages"])
    from roger_b200 import sas

    daily = {k[3:]: fx[k] for k in fx if k.startswith("in_")}
    m = sas.SasModel(OracleSasEngine, C, A, daily, {p: fx["sas_" + p] for p in sas.PROCESSES})
    m.setup()
    for k in sas.STATE:
        assert u.rel_err(m.state(k), fx["init_" + k]) < 1e-15, k


@pytest.mark.parametrize("case", ["sas_power", "sas_power_snow"])
def test_reference_itself_is_libm_sensitive(case):
    """Two runs of the UNMODIFIED reference that differ only in NumPy's `power` loop (SVML vs C library) drift apart by
    far more than 1e-10 on the power-law configurations: 1e-10 parity against BOTH of them is not a property any
    implementation of pow can have.  This backend evaluates the C library's pow (roger_pow.h returns its bits), so it is
    bit-identical to the reference run with the C library's pow and inherits that run's distance to the SVML run."""
    fx = u.load_case(case)
    drift_ref = max(u.rel_err(fx["simd_rec_sa_ss"][d], fx["rec_sa_ss"][d]) for d in range(int(fx["ndays"])))
    assert drift_ref > 1e-8
    lib().sas_oracle_set_pow(1)
    m = u.model_from_fixture(fx, OracleSasEngine)
    worst = u.replay_against_fixture(fx, m, exact=False)
    assert all(v[0] == 0.0 for v in worst.values()), worst
    # until the first rounding-decided clamp the two `power` loops of the reference agree to rounding
    for day in range(5):
        for k in ("sa_rz", "sa_ss"):
            assert u.rel_err(fx["simd_rec_" + k][day], fx["rec_" + k][day]) < 1e-12


def test_well_conditioned_case_is_libm_insensitive():
    fx = u.load_case("sas_families")
    lib().sas_oracle_set_pow(1)
    m = u.model_from_fixture(fx, OracleSasEngine)
    worst = u.replay_against_fixture(fx, m, exact=False)
    for k in ("sa_rz", "sa_ss", "msa_rz", "msa_ss", "C_rz", "C_ss", "C_s", "tt_transp", "tt_q_ss"):
        assert worst.get(k, (0, 0))[0] < 1e-10, (k, worst[k])
    drift_ref = max(u.rel_err(fx["simd_rec_sa_ss"][d], fx["rec_sa_ss"][d]) for d in range(int(fx["ndays"])))
    assert drift_ref < 1e-12


def test_mass_balance_closure_of_the_oracle():
    """water balance of the age-ranked storage: d(sum sa) = inflow - outflow, to rounding (numerics.py:475-560)"""
    fx = u.load_case("sas_power")
    m = u.model_from_fixture(fx, OracleSasEngine)
    d = {k[3:]: fx[k] for k in fx if k.startswith("in_")}
    for day in range(int(fx["ndays"])):
        s0 = m.state("sa_rz").sum(axis=1) + m.state("sa_ss").sum(axis=1)
        m.step()
        s1 = m.state("sa_rz").sum(axis=1) + m.state("sa_ss").sum(axis=1)
        t = day + 1
        net = d["INF_MAT_RZ"][:, t] + d["INF_PF_RZ"][:, t] + d["INF_PF_SS"][:, t] - d["EVAP_SOIL"][:, t] - d["TRANSP"][:, t] - d["Q_SS"][:, t]
        ref_s1 = fx["rec_sa_rz"][day].sum(axis=1) + fx["rec_sa_ss"][day].sum(axis=1)
        assert np.max(np.abs((s1 - s0) - net)) <= np.max(np.abs((ref_s1 - s0) - net)) + 1e-12
