"""Shared helpers of the StorAge-selection parity tests: fixture -> SasModel, replay, comparison."""
import os

import numpy as np

from roger_b200 import sas

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAS_CASES = ["sas_power", "sas_power_snow", "sas_families", "sas_power_stats", "sas_power_300", "sas_power_1500"]
ANION_CASES = ["sas_bromide", "sas_virtualtracer"]  # the latter: settings.enable_virtualtracer (the evaporation carries the tracer)
RK4_CASES = ["sas_rk4", "sas_rk4_stats", "sas_rk4_52"]
EULER_CASES = ["sas_euler", "sas_euler_families", "sas_euler_stats", "sas_euler_bromide", "sas_euler_52", "sas_euler_bromide_stats", "sas_euler_virtualtracer_stats"]
# gamma family (scipy.special.gammainc in the reference, roger_gamma.h here): pinned to 1e-10, not bit for bit
GAMMA_CASES = ["sas_gamma52"]


def load_case(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
        return {k: d[k] for k in d.files}


def model_from_fixture(fx, engine_cls, store_tt=True, engine_kwargs=None, mask=None, age_statistics=None):
    C = int(fx["nx"]) * int(fx["ny"])
    A = int(fx["ages"])
    daily = {k[3:]: fx[k] for k in fx if k.startswith("in_")}
    spec = {p: fx["sas_" + p] for p in sas.PROCESSES}
    if age_statistics is None:
        age_statistics = "rec_dC_num_error" in fx
    anion = "solute" in fx and int(fx["solute"]) in (1, 3)
    kw = {}
    if anion:
        kw = dict(solute=int(fx["solute"]), anion_parameters=dict(alpha_transp=fx["par_alpha_transp"], alpha_q=fx["par_alpha_q"], lu_id=fx["par_lu_id"],
                                                          S_sat_rz=fx["par_S_sat_rz"]))
    if "solver" in fx:
        kw["solver"] = int(fx["solver"])
    m = sas.SasModel(engine_cls, C, A, daily, spec, substeps=int(fx["substeps"]), store_tt=store_tt, mask=mask,
                     engine_kwargs=engine_kwargs, age_statistics=age_statistics, **kw)
    m.setup()
    for k in sas.STATE:  # start from the reference's own post-setup state
        m.engine.upload_state(k, fx["init_" + k])
    if age_statistics:
        m.engine.set_balance_reference(fx["init_C_s"])
        if "par_S_sat_rz" in fx:
            m.engine.set_sanity_bounds(fx["par_S_sat_rz"], fx["par_S_sat_ss"], fx["par_S_pwp_rz"], fx["par_S_pwp_ss"])
    return m


def same(a, b):
    """bit-level equality of values (NaN == NaN, -0.0 == 0.0)"""
    a = np.asarray(a)
    b = np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def rel_err(a, b):
    """max |a - b| / max |b| with NaN patterns required to agree"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if (np.isnan(a) != np.isnan(b)).any():
        return np.inf
    ok = ~np.isnan(b)
    if not ok.any():
        return 0.0
    scale = np.max(np.abs(b[ok]))
    d = np.max(np.abs(a[ok] - b[ok]))
    return float(d / scale) if scale > 0 else float(d)


def snapshot(model):
    """every compared quantity of the current step"""
    out = {k: model.state(k) for k in sas.STATE}
    if model.settings.store_tt:
        for p in sas.PROCESSES:
            out["tt_" + p] = model.tt(p)
    out.update(model.outputs())
    if model.settings.solute in (1, 3):  # anion transport / virtual tracer: the C_iso_* result planes carry the solute masses M_*
        for k in list(out):
            if k.startswith("C_iso_"):
                out["M_" + k[len("C_iso_"):]] = out.pop(k)
    if model.settings.age_statistics:
        out.update(model.statistics())
    out["C_in"] = model.C_in.copy()
    return out


def replay_against_fixture(fx, model, exact=True, tol=None, skip_nan_pattern=()):
    """Steps through the fixture's days; returns {name: (worst error, day)}; asserts when exact / tol is given."""
    worst = {}
    for day in range(int(fx["ndays"])):
        model.step()
        snap = snapshot(model)
        if "rec_sanity_all" in fx and "sanity_ok" in snap and not np.isnan(snap["sanity_ok"]).all():
            assert bool(np.all(snap["sanity_ok"] == 1.0)) == bool(fx["rec_sanity_all"][day]), f"day {day}: sanity_check differs"
        for k, x in snap.items():
            if "rec_" + k not in fx:
                continue
            r = fx["rec_" + k][day]
            if exact:
                assert same(x, r), f"day {day}: {k} differs from the reference fixture (max abs diff {np.nanmax(np.abs(x - r))})"
            else:
                e = rel_err(x, r)
                if e > worst.get(k, (0, 0))[0]:
                    worst[k] = (e, day)
    if tol is not None:
        bad = {k: v for k, v in worst.items() if v[0] > tol}
        assert not bad, f"above {tol}: {bad}"
    return worst


def replay_pair(model_a, model_b, ndays, exact=True):
    """Steps two models side by side and compares every quantity each day."""
    worst = {}
    for day in range(ndays):
        model_a.step()
        model_b.step()
        sa, sb = snapshot(model_a), snapshot(model_b)
        for k in sa:
            if exact:
                assert same(sa[k], sb[k]), f"day {day}: {k} differs (max abs diff {np.nanmax(np.abs(sa[k] - sb[k]))})"
            else:
                e = rel_err(sa[k], sb[k])
                if e > worst.get(k, (0, 0))[0]:
                    worst[k] = (e, day)
    return worst
