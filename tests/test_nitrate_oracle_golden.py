"""Pins the CPU oracle of the nitrate transport -- anion transport of the nitrate mass plus the soil nitrogen cycle on the
age-ranked mineral nitrogen (roger/core/nitrate.py) -- against the reference's NumPy backend (tests/golden/sas_nitrate.npz,
produced by tests/golden/make_sas_golden.py from the unmodified reference, settings.enable_nitrate): bit for bit."""
import numpy as np

from oracle.engine import OracleSasEngine, lib
from roger_b200 import nitrate, sas
from tests import util_sas as u


def model_from_fixture(fx, engine_cls, engine_kwargs=None):
    """fixtures sas_nitrate and sas_nitrate_stats (the latter with settings.enable_age_statistics)"""
    C, A = int(fx["nx"]) * int(fx["ny"]), int(fx["ages"])
    stats = "rec_dC_num_error" in fx
    daily = {k[3:]: fx[k] for k in fx if k.startswith("in_")}
    spec = {p: fx["sas_" + p] for p in sas.PROCESSES}
    params = {k: fx["par_" + k] for k in nitrate.PARAMS + ["phi_soil_temp", "damp_soil_temp"]}
    anion = dict(alpha_transp=fx["par_alpha_transp"], alpha_q=fx["par_alpha_q"], lu_id=fx["par_lu_id"], S_sat_rz=fx["par_S_sat_rz"])
    m = nitrate.NitrateModel(engine_cls, C, A, daily, spec, params, anion, nitt=int(fx["ndays"]) + 1, pi=3.14159265358979323846264338327950588, substeps=int(fx["substeps"]),
                             store_tt=True, engine_kwargs=engine_kwargs, age_statistics=stats)
    m.setup()
    for k, name in enumerate(nitrate.STATE):  # start from the reference's own post-setup state
        m.engine.upload_state_index(k, fx["init_" + name])
    if stats:
        m.engine.set_balance_reference(fx["init_C_s"])
    # The soil temperature is a host-side NumPy expression with a sine and an exponential (nitrate.py:6-66): its last bit follows
    # NumPy's SIMD dispatch (the fixture was generated with it switched off, see make_sas_golden.py).  The engine is pinned with
    # the recorded temperatures; the host expression is checked against them to rounding.
    m.temp_soil_given = np.concatenate([fx["rec_temp_soil"][:1], fx["rec_temp_soil"]])
    return m


def test_oracle_reproduces_reference_nitrate_bit_for_bit():
    lib().sas_oracle_set_pow(0)
    fx = u.load_case("sas_nitrate")
    m = model_from_fixture(fx, OracleSasEngine)
    for day in range(int(fx["ndays"])):
        m.step()
        assert u.rel_err(m.temp_soil_host, fx["rec_temp_soil"][day]) < 1e-14, f"day {day}: soil temperature (host expression)"
        for name in nitrate.STATE:
            assert u.same(m.state(name), fx["rec_" + name][day]), f"day {day}: {name} (max abs diff {np.nanmax(np.abs(m.state(name) - fx['rec_' + name][day]))})"
        out = m.nitrate_outputs()
        for k in ("nit_s", "denit_s", "min_s", "ngas_s", "nfix_s", "nh4_up", "Nmin_s"):
            assert u.same(out[k], fx["rec_" + k][day]), f"day {day}: {k}: {out[k]} vs {fx['rec_' + k][day]}"
        o = m.outputs()
        for k in ("C_rz", "C_ss", "C_s", "C_transp", "C_q_ss"):
            assert u.same(o[k], fx["rec_" + k][day]), f"day {day}: {k}"
    # the fixture exercises every process: fixation on the crop cell, nitrification, denitrification, gaseous loss, uptake
    for k in ("nit_s", "denit_s", "ngas_s", "nfix_s", "nh4_up", "min_s"):
        assert np.nanmax(fx["rec_" + k]) > 0, k


def nitrate_stats_snapshot(m):
    snap = {name: m.state(name) for name in nitrate.STATE}
    snap.update(m.nitrate_outputs())
    snap.update({k: v for k, v in m.statistics().items() if k != "sanity_ok"})
    o = m.outputs()
    for k in ("C_rz", "C_ss", "C_s", "C_transp", "C_q_ss"):
        snap[k] = o[k]
    return snap


def test_oracle_reproduces_reference_nitrate_age_statistics_bit_for_bit():
    """settings.enable_age_statistics with nitrate: water-age percentiles and mean ages of the fluxes and storages, the water balance
    error and the nitrate balance error, which carries the day's denitrified and nitrified mass (numerics.py:673-705)"""
    lib().sas_oracle_set_pow(0)
    fx = u.load_case("sas_nitrate_stats")
    m = model_from_fixture(fx, OracleSasEngine)
    for day in range(int(fx["ndays"])):
        m.step()
        for k, v in nitrate_stats_snapshot(m).items():
            assert u.same(v, fx["rec_" + k][day]), f"day {day}: {k}"
    assert np.nanmax(fx["rec_dC_num_error"]) > 0 and np.nanmax(fx["rec_denit_s"]) > 0 and np.nanmax(fx["rec_nit_s"]) > 0


# ---- crop variant of the offline model: redistribution after root growth / root loss opens the day ----------------------------
def crop_model_from_fixture(fx, engine_cls, engine_kwargs=None):
    """fixtures sas_nitrate_crop / sas_bromide_crop (settings.enable_crop_phenology, svat_crop_transport_model_deterministic) and
    their *_stats siblings (settings.enable_age_statistics on top)"""
    C, A = int(fx["nx"]) * int(fx["ny"]), int(fx["ages"])
    stats = "rec_dC_num_error" in fx
    daily = {k[3:]: fx[k] for k in fx if k.startswith("in_")}
    spec = {p: fx["sas_" + p] for p in sas.PROCESSES + ["re_rg", "re_rl"]}
    anion = dict(alpha_transp=fx["par_alpha_transp"], alpha_q=fx["par_alpha_q"], lu_id=fx["par_lu_id"], S_sat_rz=fx["par_S_sat_rz"])
    if int(fx["solute"]) == 2:
        params = {k: fx["par_" + k] for k in nitrate.PARAMS + ["phi_soil_temp", "damp_soil_temp"]}
        m = nitrate.NitrateModel(engine_cls, C, A, daily, spec, params, anion, nitt=int(fx["ndays"]) + 1, pi=3.14159265358979323846264338327950588,
                                 substeps=int(fx["substeps"]), store_tt=True, engine_kwargs=engine_kwargs, crop=True, age_statistics=stats)
        m.setup()
        for k, name in enumerate(nitrate.STATE):
            m.engine.upload_state_index(k, fx["init_" + name])
        m.temp_soil_given = np.concatenate([fx["rec_temp_soil"][:1], fx["rec_temp_soil"]])
    else:
        m = sas.SasModel(engine_cls, C, A, daily, spec, substeps=int(fx["substeps"]), store_tt=True, solute=int(fx["solute"]), anion_parameters=anion,
                         engine_kwargs=engine_kwargs, crop=True, age_statistics=stats)
        m.setup()
        for k in sas.STATE:
            m.engine.upload_state(k, fx["init_" + k])
    if stats:
        m.engine.set_balance_reference(fx["init_C_s"])
    return m


def crop_snapshot(m, fx):
    snap = {k: m.state(k) for k in sas.STATE}
    if int(fx["solute"]) == 2:
        snap.update({k: m.state(k) for k in ("Nmin_rz", "Nmin_ss")})
        snap.update(m.nitrate_outputs())
    snap.update(m.redistribution())
    o = m.outputs()
    for k in ("C_rz", "C_ss", "C_s", "C_transp", "C_q_rz", "C_q_ss", "C_cpr_rz"):
        snap[k] = o[k]
    if int(fx["solute"]) == 3:  # virtual tracer: the soil evaporation carries solute (evapotranspiration.py:723-791)
        snap["C_evap_soil"], snap["M_evap_soil"] = o["C_evap_soil"], o["C_iso_evap_soil"]
    for p in sas.PROCESSES:
        snap["tt_" + p] = m.tt(p)
    if "rec_dC_num_error" in fx:
        snap.update({k: v for k, v in m.statistics().items() if k != "sanity_ok"})
    return snap


import pytest  # noqa: E402


@pytest.mark.parametrize("case", ["sas_nitrate_crop", "sas_bromide_crop", "sas_nitrate_crop_stats", "sas_bromide_crop_stats", "sas_virtualtracer_crop"])
def test_oracle_reproduces_reference_crop_variant_bit_for_bit(case):
    lib().sas_oracle_set_pow(0)
    fx = u.load_case(case)
    m = crop_model_from_fixture(fx, OracleSasEngine)
    for day in range(int(fx["ndays"])):
        m.step()
        for k, v in crop_snapshot(m, fx).items():
            assert u.same(v, fx["rec_" + k][day]), f"day {day}: {k} (max abs diff {np.nanmax(np.abs(v - fx['rec_' + k][day]))})"
    # both directions move water and solute, and the clamp to the source layer's water is exercised
    assert np.nanmax(fx["rec_M_re_rg"]) > 0 and np.nanmax(fx["rec_M_re_rl"]) > 0
    assert (fx["rec_re_rl"] < fx["in_RE_RL"][:, 1:].T - 1.0).any() and (fx["rec_re_rg"] < fx["in_RE_RG"][:, 1:].T - 1.0).any()


def test_crop_variant_is_rejected_where_the_reference_has_none():
    daily = {"C_ISO_IN": np.zeros((2, 3)), "S_SNOW": np.zeros((2, 3))}
    spec = dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5), re_rg=(6, 0.5), re_rl=(6, 2.0))
    m = sas.SasModel(OracleSasEngine, 2, 4, dict(daily, S_RZ=np.ones((2, 3)), S_SS=np.ones((2, 3))), spec, crop=True)
    with pytest.raises(NotImplementedError):
        m.setup()
