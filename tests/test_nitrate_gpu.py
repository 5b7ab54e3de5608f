"""GPU parity of the nitrate transport (k_sas_day_nitrate: anion transport + soil nitrogen cycle, through the C ABI): bit for bit
against the CPU oracle evaluating the same pow, and against the reference fixture."""
import numpy as np
import pytest

from tests import util_sas as u
from tests.test_nitrate_oracle_golden import model_from_fixture

pytestmark = pytest.mark.gpu


def _cuda():
    from roger_b200.sas_engine import CudaSasEngine

    return CudaSasEngine


@pytest.fixture
def oracle_roger_pow():
    from oracle.engine import OracleSasEngine, lib

    lib().sas_oracle_set_pow(1)
    yield OracleSasEngine
    lib().sas_oracle_set_pow(0)


def _compare(a, b, day, exact=True, tol=None, skip=()):
    from roger_b200 import nitrate

    for name in nitrate.STATE:
        x, y = a.state(name), b.state(name) if not isinstance(b, dict) else b[name]
        if exact:
            assert u.same(x, y), f"day {day}: {name} (max abs diff {np.nanmax(np.abs(x - y))})"
        else:
            assert u.rel_err(x, y) <= tol, (day, name, u.rel_err(x, y))
    oa = dict(a.nitrate_outputs(), **a.outputs())
    ob = dict(b.nitrate_outputs(), **b.outputs()) if not isinstance(b, dict) else b
    for k in ("nit_s", "denit_s", "min_s", "ngas_s", "nfix_s", "nh4_up", "Nmin_s", "C_rz", "C_ss", "C_s", "C_transp", "C_q_rz", "C_q_ss", "C_cpr_rz"):
        if k not in ob or k in skip:
            continue
        if exact:
            assert u.same(oa[k], ob[k]), f"day {day}: {k}: {oa[k]} vs {ob[k]}"
        else:
            assert u.rel_err(oa[k], ob[k]) <= tol, (day, k, u.rel_err(oa[k], ob[k]))


def test_cuda_equals_oracle_bit_for_bit_on_fixture_inputs(oracle_roger_pow):
    fx = u.load_case("sas_nitrate")
    a, b = model_from_fixture(fx, _cuda()), model_from_fixture(fx, oracle_roger_pow)
    for day in range(int(fx["ndays"])):
        a.step()
        b.step()
        _compare(a, b, day)


def test_cuda_against_reference_fixture():
    """the same run against the reference's own numbers (libm pow there, roger_pow here): 1e-10 for 17 days for every compared quantity and for 24 days for the
    age-ranked arrays and storage concentrations; then the first
    rounding-decided clamp of the power-law solver separates the two pow implementations (1e-7, the reference's own libm
    envelope, tests/test_sas_gpu.py) and the crop cell's uptake switch at 80 % saturation (evapotranspiration.py:931-938), a
    hard threshold on sum(sa_rz), flips on single days -- C_transp is left out from there on"""
    fx = u.load_case("sas_nitrate")
    a = model_from_fixture(fx, _cuda())
    for day in range(int(fx["ndays"])):
        a.step()
        ref = {k[4:]: fx[k][day] for k in fx if k.startswith("rec_")}
        if day < 17:
            _compare(a, ref, day, exact=False, tol=1e-10)
        elif day < 24:  # the capillary-rise concentration (exponent 0.25) is the first quantity to feel the other pow: 9e-9 on day 17
            _compare(a, ref, day, exact=False, tol=1e-10, skip=("C_cpr_rz",))
        else:
            _compare(a, ref, day, exact=False, tol=1e-5, skip=("C_transp",))


@pytest.mark.parametrize("C_,A,ndays", [(300, 96, 10), (65, 513, 4), (1, 5, 6)])
def test_cuda_equals_oracle_bit_for_bit_on_seeded_inputs(C_, A, ndays, oracle_roger_pow):
    from roger_b200 import nitrate, sas

    rng = np.random.default_rng(100 + C_)
    t = ndays + 1
    wet = rng.random((C_, t)) < 0.4
    prec = np.where(wet, rng.gamma(0.8, 6.0, (C_, t)), 0.0)
    daily = {
        "PREC": prec, "INF_MAT_RZ": prec * rng.uniform(0.5, 0.8, (C_, t)), "INF_PF_RZ": prec * rng.uniform(0.0, 0.15, (C_, t)),
        "INF_PF_SS": prec * rng.uniform(0.0, 0.1, (C_, t)), "EVAP_SOIL": rng.uniform(0, 1.5, (C_, t)), "TRANSP": rng.uniform(0, 3.0, (C_, t)) * (rng.random((C_, t)) < 0.8),
        "Q_RZ": rng.gamma(0.5, 2.0, (C_, t)), "Q_SS": rng.gamma(0.5, 1.5, (C_, t)), "CPR_RZ": rng.uniform(0, 0.5, (C_, t)) * (rng.random((C_, t)) < 0.4),
        "S_RZ": np.repeat(rng.uniform(40, 160, (C_, 1)), t, axis=1), "S_SS": np.repeat(rng.uniform(150, 700, (C_, 1)), t, axis=1),
        "S_SNOW": np.zeros((C_, t)), "C_IN": rng.uniform(0, 60, (C_, t)) * (rng.random((C_, t)) < 0.2),
        "TA": rng.uniform(-5, 28, t), "DOY": np.arange(100, 100 + t), "Z_ROOT": rng.uniform(200, 900, (C_, t)),
        "LU_ID": rng.choice([8.0, 525.0, 583.0, 541.0], (C_, t)), "NUP": rng.uniform(0, 80, (C_, t)),
    }
    spec = dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.25), transp=(6, 0.5), q_rz=(6, 2.0), q_ss=(6, 3.0))
    params = {k: rng.uniform(lo, hi, C_) for k, (lo, hi) in dict(
        kfix_rz=(10, 60), kmin_rz=(20, 120), kmin_ss=(5, 40), km_nit_rz=(5, 40), dmax_nit_rz=(20, 200), kngl_rz=(1, 20), km_denit_rz=(5, 40),
        dmax_denit_rz=(10, 150), km_denit_ss=(5, 40), dmax_denit_ss=(5, 60), z_soil=(1200, 2500), damp_soil_temp=(12, 16)).items()}
    params["phi_soil_temp"] = np.full(C_, 91.0)
    params["S_sat_rz"] = daily["S_RZ"][:, 0] * rng.uniform(1.02, 1.7, C_)
    params["S_sat_ss"] = daily["S_SS"][:, 0] * rng.uniform(1.02, 1.7, C_)
    anion = dict(alpha_transp=rng.uniform(0.2, 0.9, C_), alpha_q=rng.uniform(0.4, 1.0, C_), lu_id=daily["LU_ID"][:, 0], S_sat_rz=params["S_sat_rz"])
    mask = (rng.random(C_) > 0.05).astype(np.float64)
    models = []
    for eng in (_cuda(), oracle_roger_pow):
        m = nitrate.NitrateModel(eng, C_, A, daily, spec, params, anion, nitt=t, mask=mask)
        m.setup()
        sa_rz, sa_ss = sas.initial_age_distribution(daily["S_RZ"][:, 0], A), sas.initial_age_distribution(daily["S_SS"][:, 0], A)
        decay = np.exp(-np.linspace(0.001, 7, A))[None, :]
        for k, arr in enumerate([sa_rz, sa_ss, 30.0 * sa_rz, 15.0 * sa_ss, 600.0 * decay * np.ones((C_, 1)), 200.0 * decay * np.ones((C_, 1))]):
            m.engine.upload_state_index(k, arr)
        models.append(m)
    for day in range(ndays):
        for m in models:
            m.step()
        _compare(models[0], models[1], day)


# ---- crop variant (svat_crop_transport_model_deterministic): redistribution after root growth / root loss -----------------------
@pytest.mark.parametrize("case", ["sas_nitrate_crop", "sas_bromide_crop", "sas_nitrate_crop_stats", "sas_bromide_crop_stats", "sas_virtualtracer_crop"])
def test_crop_variant_cuda_equals_oracle_bit_for_bit_on_fixture_inputs(case, oracle_roger_pow):
    from tests.test_nitrate_oracle_golden import crop_model_from_fixture, crop_snapshot

    fx = u.load_case(case)
    a, b = crop_model_from_fixture(fx, _cuda()), crop_model_from_fixture(fx, oracle_roger_pow)
    moved = 0.0
    for day in range(int(fx["ndays"])):
        a.step()
        b.step()
        sa, sb = crop_snapshot(a, fx), crop_snapshot(b, fx)
        for k in sa:
            assert u.same(sa[k], sb[k]), f"day {day}: {k} (max abs diff {np.nanmax(np.abs(sa[k] - sb[k]))})"
        moved += float(np.sum(sa["M_re_rg"]) + np.sum(sa["M_re_rl"]))
    assert moved > 0


@pytest.mark.parametrize("case", ["sas_nitrate_crop", "sas_bromide_crop"])
def test_crop_variant_cuda_against_reference_fixture(case):
    """against the reference's own numbers (libm pow there, roger_pow here): 1e-10 for every compared quantity on every day (no
    rounding-decided clamp of the power-law solver falls into these fixtures; the oracle with roger_pow stays below 1e-14)"""
    from tests.test_nitrate_oracle_golden import crop_model_from_fixture, crop_snapshot

    fx = u.load_case(case)
    a = crop_model_from_fixture(fx, _cuda())
    for day in range(int(fx["ndays"])):
        a.step()
        for k, v in crop_snapshot(a, fx).items():
            assert u.rel_err(v, fx["rec_" + k][day]) <= 1e-10, (day, k, u.rel_err(v, fx["rec_" + k][day]))


def test_crop_variant_needs_the_days_fluxes_and_rejects_other_configurations():
    from roger_b200 import sas
    from roger_b200.engine import RogerCudaError

    eng = _cuda()(4, sas.default_settings(8, solute=sas.ANION))
    rows = sas.sas_parameter_block(4, dict(re_rg=(6, 0.5), re_rl=(6, 2.0)), processes=("re_rg", "re_rl"))
    spec = dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))
    eng.set_parameters(np.ones(4), sas.sas_parameter_block(4, spec))
    eng.set_crop_parameters(rows[0], rows[1])
    with pytest.raises(RogerCudaError, match="set_redistribution"):
        eng.step(np.zeros((9, 4)))
    eng.set_redistribution(np.zeros(4), np.zeros(4))
    eng.step(np.zeros((9, 4)))
    for kw in (dict(), dict(solute=sas.ANION, solver=sas.EULER)):
        with pytest.raises(RogerCudaError, match="crop variant"):
            _cuda()(4, sas.default_settings(8, **kw)).set_crop_parameters(rows[0], rows[1])


@pytest.mark.parametrize("case", ["sas_nitrate", "sas_nitrate_crop", "sas_bromide_crop", "sas_nitrate_crop_stats", "sas_bromide_crop_stats",
                                  "sas_virtualtracer_crop"])
def test_cuda_reproduces_reference_fixture_bit_for_bit(case):
    """nitrate (anion transport + soil nitrogen cycle) and the crop variants (the *_stats ones with age statistics and balance errors) against the reference's own numbers, bit for bit on every
    day (the soil temperature, a host-side expression, is taken from the fixture as in the oracle pin)"""
    from roger_b200 import nitrate
    from tests.test_nitrate_oracle_golden import crop_model_from_fixture, crop_snapshot

    fx = u.load_case(case)
    if case == "sas_nitrate":
        m = model_from_fixture(fx, _cuda())
        for day in range(int(fx["ndays"])):
            m.step()
            for name in nitrate.STATE:
                assert u.same(m.state(name), fx["rec_" + name][day]), (day, name)
            out = dict(m.nitrate_outputs(), **m.outputs())
            for k in ("nit_s", "denit_s", "min_s", "ngas_s", "nfix_s", "nh4_up", "Nmin_s", "C_rz", "C_ss", "C_s", "C_transp", "C_q_ss"):
                assert u.same(out[k], fx["rec_" + k][day]), (day, k)
        return
    m = crop_model_from_fixture(fx, _cuda())
    for day in range(int(fx["ndays"])):
        m.step()
        for k, v in crop_snapshot(m, fx).items():
            assert u.same(v, fx["rec_" + k][day]), (day, k)


def test_nitrate_age_statistics_cuda_equals_reference_fixture_and_oracle(oracle_roger_pow):
    """settings.enable_age_statistics with nitrate (k_sas_day_nitrate<STATS = true>): the reference fixture bit for bit, the oracle too"""
    from tests.test_nitrate_oracle_golden import nitrate_stats_snapshot

    fx = u.load_case("sas_nitrate_stats")
    a, b = model_from_fixture(fx, _cuda()), model_from_fixture(fx, oracle_roger_pow)
    for day in range(int(fx["ndays"])):
        a.step()
        b.step()
        sa, sb = nitrate_stats_snapshot(a), nitrate_stats_snapshot(b)
        for k in sa:
            assert u.same(sa[k], sb[k]), (day, k, "oracle")
            assert u.same(sa[k], fx["rec_" + k][day]), (day, k, "reference fixture")
