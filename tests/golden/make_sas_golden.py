"""Generates tests/golden/sas_*.npz: the reference's deterministic oxygen-18 StorAge-selection transport
(roger/core/transport.py:948-991) run by the UNMODIFIED NumPy backend on seeded synthetic inputs.
Daily hydrologic forcing comes from an online SVAT run of the (already pinned) CPU oracle.

    python tests/golden/make_sas_golden.py [case ...]

Which libm.  The SAS solver amplifies last-bit differences of `power` ("SAS parity chain"), so the
fixture has to say which `power` produced it.  The main records (`rec_*`) are produced with NumPy's AVX512 / AVX2
SIMD dispatch switched off (NPY_DISABLE_CPU_FEATURES, a documented NumPy runtime switch): float64 `power` then calls
the C library's pow, which is what the C oracle calls, and oracle and reference agree bit for bit.  A second run with
NumPy's default dispatch (Intel SVML `power` on this AVX512 host) is stored as `simd_rec_*` for a few variables: it
documents how far two runs of the *unmodified reference itself* drift apart when only the libm differs.
"""
import os
import subprocess
import sys
import tempfile

NO_SIMD = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2 FMA3"
if __name__ == "__main__" and "--simd" not in sys.argv and os.environ.get("NPY_DISABLE_CPU_FEATURES") != NO_SIMD:
    os.environ["NPY_DISABLE_CPU_FEATURES"] = NO_SIMD  # must be set before numpy is imported
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402
from oracle.engine import OracleSvatEngine  # noqa: E402
from roger_b200 import sas_inputs, synth  # noqa: E402
from roger_b200.svat import SvatModel  # noqa: E402

CASES = {
    # power-law SAS of benchmarks/SVATOXYGEN18_benchmark.py:129-138
    "sas_power": dict(shape=(2, 3), ndays=30, ages=60, doy=100, sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    # winter start: snow storage, isotope mixing in the snow pack
    "sas_power_snow": dict(shape=(2, 2), ndays=25, ages=50, doy=20, sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5)), var_input=True),
    # age statistics (percentile ages by numpy.interp) and the water / solute balance errors
    "sas_power_stats": dict(shape=(2, 2), ndays=16, ages=40, doy=105, age_statistics=True,
                            sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    # BASELINE-sized age axes: NumPy's pairwise summation recurses for n > 128 (sums over the age axis), np.interp walks long CDFs
    "sas_power_300": dict(shape=(2, 2), ndays=8, ages=300, doy=105, age_statistics=True,
                          sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    "sas_power_1500": dict(shape=(1, 2), ndays=4, ages=1500, doy=100,
                           sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    # groundwater boundary condition (settings.enable_groundwater_boundary): svat_lbc_transport_model_deterministic
    "sas_power_lbc": dict(shape=(2, 2), ndays=12, ages=40, doy=100, groundwater_boundary=True, age_statistics=True,
                          sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    # conservative anion (bromide): solute mass per age bin, partition coefficients, a crop cell whose uptake stops when wet
    "sas_bromide": dict(shape=(2, 3), ndays=26, ages=50, doy=100, solute="bromide", age_statistics=True,
                        sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0))),
    # nitrate: the anion transport plus the soil nitrogen cycle on age-ranked mineral nitrogen (nitrate.py), a nitrogen-fixing crop
    # cell, a fertiliser pulse, wet cells above the denitrification threshold
    "sas_nitrate": dict(shape=(2, 3), ndays=30, ages=50, doy=100, solute="nitrate",
                        sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.25), transp=(6, 0.5), q_rz=(6, 2.0), q_ss=(6, 3.0))),
    "sas_nitrate_stats": dict(shape=(2, 2), ndays=14, ages=40, doy=100, solute="nitrate", age_statistics=True,
                              sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.25), transp=(6, 0.5), q_rz=(6, 2.0), q_ss=(6, 3.0))),
    # the crop variant of the offline model (svat_crop_transport_model_deterministic): redistribution between subsoil and root zone
    # after root growth (re_rg) / root loss (re_rl) opens the day; a harvest pulse larger than the root zone holds is clamped
    "sas_nitrate_crop": dict(shape=(2, 3), ndays=24, ages=50, doy=100, solute="nitrate", crop=True,
                             sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.25), transp=(6, 0.5), q_rz=(6, 2.0), q_ss=(6, 3.0),
                                      re_rg=(6, 0.5), re_rl=(6, 10.0))),
    "sas_bromide_crop": dict(shape=(2, 2), ndays=16, ages=40, doy=100, solute="bromide", crop=True,
                             sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0),
                                      re_rg=(3, 0.7, 1.6), re_rl=(1,))),
    # the crop variant with age statistics and balance errors (numerics.py:445-467: the water balance of the crop variant;
    # :642-715: the bromide / nitrate solute balances do not exclude the crop model)
    "sas_bromide_crop_stats": dict(shape=(2, 2), ndays=14, ages=40, doy=100, solute="bromide", crop=True, age_statistics=True,
                                   sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0),
                                            re_rg=(6, 0.5), re_rl=(6, 10.0))),
    "sas_nitrate_crop_stats": dict(shape=(2, 2), ndays=16, ages=40, doy=100, solute="nitrate", crop=True, age_statistics=True,
                                   sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.25), transp=(6, 0.5), q_rz=(6, 2.0), q_ss=(6, 3.0),
                                            re_rg=(3, 0.7, 1.6), re_rl=(6, 10.0))),
    # virtual tracer (settings.enable_virtualtracer): the anion model whose soil evaporation carries the tracer with partition
    # coefficient 1 (evapotranspiration.py:723-791); its own solute balance (numerics.py:517-549); with and without the crop variant
    "sas_virtualtracer": dict(shape=(2, 3), ndays=22, ages=50, doy=100, solute="virtualtracer", age_statistics=True,
                              sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0))),
    "sas_virtualtracer_crop": dict(shape=(2, 2), ndays=16, ages=40, doy=100, solute="virtualtracer", crop=True,
                                   sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0),
                                            re_rg=(6, 0.5), re_rl=(6, 10.0))),
    # mixed families: uniform, Kumaraswamy, exponential, storage-dependent power, dirac
    "sas_families": dict(shape=(2, 3), ndays=20, ages=45, doy=130,
                         sas=dict(evap_soil=(1,), cpr_rz=(51, 2.0), transp=(3, 0.7, 1.6), q_rz=(61, 0.0, 0.0, 0.5, 2.0, 60.0, 160.0), q_ss=(35, 0.0, 0.0, 0.4, 1.5, 150.0, 700.0))),
    # explicit Euler solver (settings.sas_solver = "Euler", 6 steps per day): benchmark power law; snow + variable input with mixed families
    "sas_euler": dict(shape=(2, 3), ndays=24, ages=50, doy=100, sas_solver="Euler",
                      sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    "sas_euler_stats": dict(shape=(2, 2), ndays=14, ages=40, doy=105, sas_solver="Euler", age_statistics=True,
                            sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    "sas_euler_bromide": dict(shape=(2, 3), ndays=20, ages=50, doy=100, solute="bromide", sas_solver="Euler",
                              sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0))),
    "sas_euler_bromide_stats": dict(shape=(2, 2), ndays=16, ages=40, doy=100, solute="bromide", sas_solver="Euler", age_statistics=True,
                                    sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0))),
    # the virtual tracer under the Euler scheme: the anion branch of svat_transport_model_euler (transport.py:2308-2331 -- no solute
    # leaves with the evaporation there), the virtual tracer's own balance error (numerics.py:517-549)
    "sas_euler_virtualtracer_stats": dict(shape=(2, 2), ndays=16, ages=40, doy=100, solute="virtualtracer", sas_solver="Euler", age_statistics=True,
                                          sas=dict(evap_soil=(6, 0.1), cpr_rz=(6, 0.1), transp=(6, 0.3), q_rz=(6, 2.0), q_ss=(6, 3.0))),
    # Runge-Kutta solver (settings.sas_solver = "RK4")
    "sas_rk4": dict(shape=(2, 2), ndays=12, ages=40, doy=100, sas_solver="RK4",
                    sas=dict(evap_soil=(6, 0.2), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    "sas_rk4_stats": dict(shape=(2, 2), ndays=10, ages=36, doy=105, sas_solver="RK4", age_statistics=True, var_input=True,
                          sas=dict(evap_soil=(6, 0.2), cpr_rz=(3, 0.7, 1.6), transp=(6, 0.5), q_rz=(61, 0.0, 0.0, 0.5, 2.0, 60.0, 160.0), q_ss=(6, 1.5))),
    # the gamma family (id 4, scipy.special.gammainc) and the exponential family read backwards along the age axis (id 52), with
    # each of the three solvers
    "sas_gamma52": dict(shape=(2, 2), ndays=16, ages=45, doy=110, var_input=True,
                        sas=dict(evap_soil=(52, 2.0), cpr_rz=(4, 2.0, 3.0), transp=(4, 1.5, 4.0), q_rz=(52, 1.0), q_ss=(4, 0.8, 2.5))),
    # (under the Euler / RK4 schemes the gamma family makes the reference divide by a rounding residue -- its cumulative function
    # returns to 0 at S, the normalising sum telescopes to ~0, "Solution of SAS diverged" -- so only family 52 is pinned there)
    "sas_euler_52": dict(shape=(2, 2), ndays=14, ages=40, doy=110, sas_solver="Euler", var_input=True,
                         sas=dict(evap_soil=(52, 2.0), cpr_rz=(6, 0.2), transp=(6, 0.5), q_rz=(52, 1.0), q_ss=(6, 1.5))),
    "sas_rk4_52": dict(shape=(2, 2), ndays=8, ages=36, doy=110, sas_solver="RK4",
                       sas=dict(evap_soil=(6, 0.2), cpr_rz=(52, 2.0), transp=(52, 1.5), q_rz=(6, 1.5), q_ss=(6, 1.5))),
    "sas_euler_families": dict(shape=(2, 2), ndays=22, ages=45, doy=20, sas_solver="Euler", var_input=True,
                               sas=dict(evap_soil=(1,), cpr_rz=(51, 2.0), transp=(3, 0.7, 1.6), q_rz=(61, 0.0, 0.0, 0.5, 2.0, 60.0, 160.0),
                                        q_ss=(35, 0.0, 0.0, 0.4, 1.5, 150.0, 700.0))),
}


def build(name):
    c = CASES[name]
    nx, ny = c["shape"]
    ndays, ages = c["ndays"], c["ages"]
    forcing = synth.mixed_forcing(ndays + 2, seed=7, start_doy=c["doy"])
    prim, th_rz, th_ss = synth.heterogeneous_parameters((nx, ny), seed=5, lu_choices=(5, 6, 8, 9, 10))
    svat = SvatModel(OracleSvatEngine, (nx, ny), forcing, prim, th_rz, th_ss)
    svat.setup()
    daily = sas_inputs.daily_fluxes_from_svat(svat, ndays)
    case = dict(nx=nx, ny=ny, ndays=ndays, ages=ages, age_statistics=bool(c.get("age_statistics", False)),
                sas_solver=c.get("sas_solver", "deterministic"), groundwater_boundary=bool(c.get("groundwater_boundary", False)))
    for k in list(sas_inputs.SUMS) + list(sas_inputs.STATES):
        case[k] = daily[k].reshape(nx, ny, ndays + 1)
    for k in ("S_pwp_rz", "S_pwp_ss"):
        case[k] = daily[k].reshape(nx, ny)
    case["S_sat_rz"] = svat.get("S_pwp_rz").reshape(nx, ny) * 0 + 1e4
    case["S_sat_ss"] = svat.get("S_pwp_ss").reshape(nx, ny) * 0 + 1e4
    rng = np.random.default_rng(11)
    ciso = np.full((nx, ny, ndays + 1), -7.0)
    if c.get("var_input"):
        ciso = rng.uniform(-12, -5, (nx, ny, ndays + 1))
    ciso[:, :, 0] = np.nan
    case["C_ISO_IN"] = ciso
    if c.get("solute") in ("bromide", "nitrate", "virtualtracer"):
        # a bromide pulse dissolved in the rain of days 2..4 (mg/l), nothing afterwards
        cin = np.zeros((nx, ny, ndays + 1))
        wet = np.flatnonzero(case["PREC"][0, 0, 1:] > 0) + 1
        for d in list(wet[:3]) + list(wet[4:5]):  # the first three wet days and one later day carry bromide
            cin[:, :, d] = rng.uniform(20, 80, (nx, ny)) * (case["PREC"][:, :, d] > 0)
        case["C_IN"] = cin
        case["alpha_transp"] = rng.uniform(0.2, 0.9, (nx, ny))
        case["alpha_q"] = rng.uniform(0.4, 1.0, (nx, ny))
        lu = np.full((nx, ny), 8.0)
        lu[0, 1] = 525.0  # a crop cell ...
        case["lu_id"] = lu
        case["S_sat_rz"] = np.full((nx, ny), 1e4)
        case["S_sat_rz"][0, 1] = case["S_RZ"][0, 1, 0] * 1.1  # ... close to saturation: uptake is switched off on wet days
        case["solute"] = c["solute"]
    if c.get("solute") == "nitrate":
        t = ndays + 1
        case["TA"] = forcing["TA"][::144][:t].copy()
        case["DOY"] = forcing["DOY"][::144][:t].astype(np.int64)
        case["Z_ROOT"] = np.repeat(rng.uniform(300, 900, (nx, ny, 1)), t, axis=2)
        lu_t = np.repeat(case["lu_id"][:, :, None], t, axis=2)
        lu_t[1, 0, :] = 583  # a nitrogen-fixing crop (nitrate.py:339-343)
        case["LU_ID"] = lu_t
        case["NUP"] = rng.uniform(0.0, 80.0, (nx, ny, t))
        case["z_soil"] = rng.uniform(1200, 2500, (nx, ny))
        # storages close to the pore volume in some cells: denitrification needs S >= 0.7 S_sat, nitrification S < 0.9 S_sat
        case["S_sat_rz"] = case["S_RZ"][:, :, 0] * rng.uniform(1.05, 1.6, (nx, ny))
        case["S_sat_ss"] = case["S_SS"][:, :, 0] * rng.uniform(1.05, 1.6, (nx, ny))
        for k, (lo, hi) in dict(kfix_rz=(10, 60), kmin_rz=(20, 120), kmin_ss=(5, 40), km_nit_rz=(5, 40), dmax_nit_rz=(20, 200), kngl_rz=(1, 20),
                                km_denit_rz=(5, 40), dmax_denit_rz=(10, 150), km_denit_ss=(5, 40), dmax_denit_ss=(5, 60), damp_soil_temp=(12, 16)).items():
            case[k] = rng.uniform(lo, hi, (nx, ny))
        case["phi_soil_temp"] = np.full((nx, ny), 91.0)
        case.update(C_rz0=30.0, C_ss0=15.0, NMIN_RZ0=60.0, NMIN_SS0=20.0)
    if c.get("crop"):
        t = ndays + 1
        rg = np.zeros((nx, ny, t))
        rl = np.zeros((nx, ny, t))
        rg[:, :, 2:9] = rng.uniform(0.2, 3.0, (nx, ny, 7)) * (rng.random((nx, ny, 7)) < 0.8)  # roots grow for a week ...
        rl[:, :, 12] = rng.uniform(2.0, 25.0, (nx, ny))                                         # ... and are lost at harvest
        rl[0, 0, 12] = 1e4                                                                      # more than the root zone holds
        rg[1, 0, 14] = 1e4                                                                      # more than the subsoil holds
        rl[:, :, 15:18] = rng.uniform(0.0, 1.0, rl[:, :, 15:18].shape)
        case["RE_RG"], case["RE_RL"] = rg, rl
        case["crop"] = True
    sas = {}
    for proc, tup in c["sas"].items():
        arr = np.zeros((nx, ny, 8))
        for j, v in enumerate(tup):
            arr[:, :, j] = v
        sas[proc] = arr
    case["sas"] = sas
    return case


SIMD_KEEP = ["sa_rz", "sa_ss", "C_rz", "C_iso_transp", "C_iso_q_ss"]


def run(name, simd_out=None):
    case = build(name)
    nitrate = case.get("solute") == "nitrate"
    tracer = case.get("solute") == "virtualtracer"
    bromide = case.get("solute") in ("bromide", "nitrate", "virtualtracer")
    Setup = rh.make_nitrate_setup(case) if nitrate else (rh.make_virtualtracer_setup(case) if tracer else (rh.make_bromide_setup(case) if bromide else rh.make_o18_setup(case)))
    crop = bool(case.get("crop"))
    if crop:
        Setup = rh.with_crop_redistribution(Setup, case)
    model = Setup()
    model.setup()
    model.warmup()
    state = model.state
    vs = state.variables
    nx, ny, ages = case["nx"], case["ny"], case["ages"]
    C = nx * ny
    age_vars = ["sa_rz", "sa_ss", "msa_rz", "msa_ss", "sa_s", "msa_s"]
    tt_vars = ["tt_evap_soil", "tt_transp", "tt_q_rz", "tt_q_ss", "tt_cpr_rz"]
    scal_vars = ["C_evap_soil", "C_transp", "C_q_rz", "C_q_ss", "C_cpr_rz", "C_iso_evap_soil", "C_iso_transp", "C_iso_q_rz", "C_iso_q_ss",
                 "C_iso_cpr_rz", "C_inf_mat_rz", "C_inf_pf_rz", "C_inf_pf_ss", "C_in"]
    lvl_vars = ["C_rz", "C_ss", "C_s", "C_iso_rz", "C_iso_ss", "C_iso_s", "C_snow"]
    if bromide:
        scal_vars = ["C_transp", "C_q_rz", "C_q_ss", "C_cpr_rz", "C_inf_mat_rz", "C_inf_pf_rz", "C_inf_pf_ss", "C_in",
                     "M_transp", "M_q_rz", "M_q_ss", "M_cpr_rz", "M_inf_mat_rz", "M_inf_pf_rz", "M_inf_pf_ss"]
        lvl_vars = ["C_rz", "C_ss", "C_s", "M_rz", "M_ss", "M_s"]
    if tracer:
        scal_vars = scal_vars + ["C_evap_soil", "M_evap_soil"]
    if nitrate:
        age_vars = age_vars + ["Nmin_rz", "Nmin_ss"]
        scal_vars = scal_vars + ["nit_s", "denit_s", "min_s", "ngas_s", "nfix_s", "nh4_up"]
        lvl_vars = lvl_vars + ["Nmin_s", "temp_soil"]
    if crop:
        tt_vars = tt_vars + ["tt_re_rg", "tt_re_rl"]
        scal_vars = scal_vars + ["re_rg", "re_rl", "C_re_rg", "C_re_rl", "M_re_rg", "M_re_rl"]
    from roger_b200 import sas as _sas

    if case["age_statistics"]:
        scal_vars = scal_vars + [n for n in _sas.STATISTICS if n != "sanity_ok"]

    def grab():
        d = {}
        for v in age_vars:
            d[v] = np.asarray(getattr(vs, v))[2:-2, 2:-2, 1, :].reshape(C, -1).copy()
        for v in tt_vars:
            d[v] = np.asarray(getattr(vs, v))[2:-2, 2:-2, :].reshape(C, -1).copy()
        for v in scal_vars:
            d[v] = np.asarray(getattr(vs, v))[2:-2, 2:-2].reshape(C).copy()
        for v in lvl_vars:
            d[v] = np.asarray(getattr(vs, v))[2:-2, 2:-2, 1].reshape(C).copy()
        return d

    # the scalar result of numerics.sanity_check as write_output sees it (first call of every step, transport.py:3406)
    import roger.core.numerics as numerics
    import roger.core.transport as transport_mod

    sanity = []
    orig_check = numerics.sanity_check
    seen = {"itt": None}

    euler = case["sas_solver"] in ("Euler", "RK4")

    def recording_check(st):
        r = orig_check(st)
        if seen["itt"] != int(vs.itt):
            seen["itt"] = int(vs.itt)
            sanity.append(bool(r))
        elif euler:  # the Euler loop checks after every step of the day: the last one is what the day's record shows
            sanity[-1] = bool(r)
        return r

    numerics.sanity_check = recording_check
    init = grab()
    recs = []
    for day in range(case["ndays"]):
        model.step(state)
        recs.append(grab())
    numerics.sanity_check = orig_check
    if simd_out is not None:
        np.savez(simd_out, **{"simd_rec_" + k: np.stack([r[k] for r in recs]) for k in SIMD_KEEP if k in recs[0]})
        return
    out = dict(nx=nx, ny=ny, ndays=case["ndays"], ages=ages, substeps=6)
    with tempfile.TemporaryDirectory() as td:
        tmp = os.path.join(td, "simd.npz")
        env = {k: v for k, v in os.environ.items() if k != "NPY_DISABLE_CPU_FEATURES"}
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--simd", tmp, name], env=env, stdout=subprocess.DEVNULL)
        with np.load(tmp) as z:
            out.update({k: z[k] for k in z.files})
    for k in list(sas_inputs.SUMS) + list(sas_inputs.STATES) + ["C_ISO_IN"] + (["C_IN"] if bromide else []):
        out["in_" + k] = case[k].reshape(C, -1)
    if case["sas_solver"] in ("Euler", "RK4"):
        out["solver"] = np.asarray({"Euler": 1, "RK4": 2}[case["sas_solver"]])
    if bromide:
        out["solute"] = np.asarray(2 if nitrate else (3 if tracer else 1))
        for k in ("alpha_transp", "alpha_q", "lu_id"):
            out["par_" + k] = np.asarray(case[k], dtype=np.float64).reshape(C)
    if nitrate:
        for k in ("kfix_rz", "kmin_rz", "kmin_ss", "km_nit_rz", "dmax_nit_rz", "kngl_rz", "km_denit_rz", "dmax_denit_rz", "km_denit_ss",
                  "dmax_denit_ss", "phi_soil_temp", "damp_soil_temp", "z_soil"):
            out["par_" + k] = np.asarray(case[k], dtype=np.float64).reshape(C)
        out["in_TA"] = case["TA"]
        out["in_DOY"] = case["DOY"]
        for k in ("Z_ROOT", "LU_ID", "NUP"):
            out["in_" + k] = np.asarray(case[k], dtype=np.float64).reshape(C, -1)
        for k in ("C_rz0", "C_ss0", "NMIN_RZ0", "NMIN_SS0"):
            out["par_" + k] = np.asarray(case[k])
    if crop:
        out["in_RE_RG"] = case["RE_RG"].reshape(C, -1)
        out["in_RE_RL"] = case["RE_RL"].reshape(C, -1)
    for proc, arr in case["sas"].items():
        out["sas_" + proc] = arr.reshape(C, 8)
    for k in ("S_sat_rz", "S_sat_ss", "S_pwp_rz", "S_pwp_ss"):
        out["par_" + k] = np.asarray(case[k], dtype=np.float64).reshape(C)
    if len(sanity) == case["ndays"]:
        out["rec_sanity_all"] = np.asarray(sanity)
    for k, v in init.items():
        out["init_" + k] = v
    for k in recs[0]:
        out["rec_" + k] = np.stack([r[k] for r in recs])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "days", case["ndays"], "size %.1f KB" % (os.path.getsize(path) / 1024),
          "sum sa_rz", float(np.nansum(recs[-1]["sa_rz"])), "C_q_ss", recs[-1]["C_q_ss"][:3])


if __name__ == "__main__":
    if "--simd" in sys.argv:
        k = sys.argv.index("--simd")
        run(sys.argv[k + 2], simd_out=sys.argv[k + 1])
    else:
        for n in sys.argv[1:] or list(CASES):
            run(n)
