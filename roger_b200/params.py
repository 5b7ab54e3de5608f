"""Host-side (setup-time / monthly) parameter derivation for the SVAT / oneD models.

These run once at setup and once per month, not per time step, so they stay NumPy on the host
(SURVEY.md 2.1 "Parameter derivation"); the per-step hot path is CUDA only.  Each function
restates one reference kernel and is checked against the reference's own output in
tests/test_params.py (fixtures `setup_*` in tests/golden/*.npz).

All arrays are flat per-cell vectors (interior cells only, C-order of the (nx, ny) grid).
"""
import os

import numpy as np

_LUT = None


def luts():
    """Numeric look-up tables of roger/lookuptables.py (ARR_ILU, ARR_GC, ARR_GCM, ARR_RDLU, ARR_MLMS)."""
    global _LUT
    if _LUT is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lut.npz")
        with np.load(path) as d:
            _LUT = {k: np.asarray(d[k], dtype=np.float64) for k in d.files}
    return _LUT


def _row_no(col, i):
    """roger/core/utilities.py:6-27 `_get_row_no`: first row whose key equals i, else row 0."""
    hit = np.nonzero(col == i)[0]
    return int(hit[0]) if hit.size else 0


def masks(lu_id):
    """roger/core/surface.py:40-71 calc_topo_kernel."""
    lu = np.asarray(lu_id)
    river = lu == 20
    lake = lu == 14
    catch = (lu != 14) & (lu != 20) & (lu != 999)
    return catch, river, lake


def surface_parameters(lu_id, month, ta, swe_top_tot, c_int=None):
    """roger/core/surface.py:74-343 calc_parameters_surface_kernel (run at setup and on month change,
    models/svat/svat.py:116-120).  month is vs.month[tau] (1..12) used as LUT column."""
    L = luts()
    lu = np.asarray(lu_id).astype(np.int64)
    C = lu.shape[0]
    catch, river, lake = masks(lu)
    m = catch.astype(np.float64)
    c_int = np.ones(C) if c_int is None else np.asarray(c_int, dtype=np.float64)
    month = int(month)
    ilu, gc, gcm = L["ILU"], L["GC"], L["GCM"]

    S_top = np.zeros(C)
    for i in range(10, 16):
        if i in (10, 11, 12, 15, 17):
            S_top = np.where(lu == i, ilu[_row_no(ilu[:, 0], i), month], S_top) * m
        else:
            S_top = S_top * m
    S_int_top_tot = S_top * c_int * m

    ground_list = (0, 5, 6, 7, 8, 9, 13, 98, 31, 32, 33, 40, 41, 50, 60, 98)
    S_gr = np.zeros(C)
    for i in range(0, 81):
        if i in ground_list:
            S_gr = np.where(lu == i, ilu[_row_no(ilu[:, 0], i), month], S_gr) * m
    for i in range(10, 16):
        if i in (10, 11, 12, 15, 16):
            S_gr = np.where(lu == i, 1.0, S_gr) * m
    S_int_ground_tot = S_gr * c_int * m

    cc_list = (0, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 98, 31, 32, 33, 40, 41, 50, 60, 90, 98)
    ground_cover = np.zeros(C)
    btc = np.zeros(C)
    bec = np.zeros(C)
    with np.errstate(all="ignore"):
        for i in range(0, 81):
            if i in cc_list:
                r = _row_no(gc[:, 0], i)
                sel = lu == i
                ground_cover = np.where(sel, gc[r, month], ground_cover) * m
                btc = np.where(sel, gc[r, month] / gcm[r, 1], btc) * m
                bec = np.where(sel, 1 - ((gc[r, month] / gcm[r, 1]) * gcm[r, 1]), bec) * m
        btc = np.where(river | lake, 0, btc)
        basal_transp_coeff = btc * m
        # surface.py:230 `maskRiver | maskLake | lu_id == 0` parses as `(maskRiver | maskLake | lu_id) == 0`
        bec = np.where((river.astype(np.int64) | lake.astype(np.int64) | lu) == 0, 1, bec)
        basal_evap_coeff = bec * m

        ta = np.asarray(ta, dtype=np.float64)
        s = np.asarray(swe_top_tot, dtype=np.float64).copy()
        for lid, v in ((10, 9), (11, 15), (12, 25)):
            s = np.where((ta > -1) & (lu == lid), v, s) * m
        for lid, v in ((10, 9), (11, 15), (12, 25)):
            s = np.where((ta >= -3) & (ta <= -1) & (lu == lid), 2.5 + 0.5 * ta * v, s) * m
        for lid, v in ((10, 18), (11, 30), (12, 50)):
            s = np.where((ta < -3) & (lu == lid), v, s) * m

        lai = np.log(1 / (1 - ground_cover)) / np.log(1 / 0.7) * m
        tf = np.where(lai > 1, 0.1, 1.0 - lai)
        throughfall_coeff_top = np.where(np.isin(lu, (10, 11, 12)), tf, 0) * m
        throughfall_coeff_ground = np.where((lu >= 500) & (lu < 598), tf, 0) * m
    return dict(
        S_int_top_tot=S_int_top_tot,
        S_int_ground_tot=S_int_ground_tot,
        ground_cover=ground_cover * m,
        basal_transp_coeff=basal_transp_coeff,
        basal_evap_coeff=basal_evap_coeff,
        swe_top_tot=s,
        lai=lai,
        throughfall_coeff_top=throughfall_coeff_top,
        throughfall_coeff_ground=throughfall_coeff_ground,
    )


SETTINGS_DEFAULTS = dict(
    # roger/settings.py:18-136
    pi=3.14159265358979323846264338327950588, r_mp=2.5, l_sc=10000.0, sf=3.0, ta_fm=0.0, rmax=30.0,
    end_event=21600, hpi=5, a_bc=2, b_bc=2, clay_min=0.01, clay_max=0.71, theta_rew_min=0.02,
    theta_rew_max=0.24, zroot_to_zsoil_max=0.7, rew_min=2.0, rew_max=12.0, z_evap_max=150.0,
    transp_water_stress=0.75, atol=1e-2, rtol=1e-2, dx=1.0,
)


def soil_parameters(p, settings=None, lateral=False, slope=None):
    """roger/core/soil.py:143-560 calc_parameters_soil_kernel / _root_zone_kernel / _subsoil_kernel
    (+ :560-640 lateral-flow macropore velocities).  `p` holds the primary per-cell parameters
    lu_id, z_soil, dmpv, lmpv, theta_ac, theta_ufc, theta_pwp, ks, kf (+ c_root)."""
    st = dict(SETTINGS_DEFAULTS)
    if settings:
        st.update(settings)
    L = luts()
    lu = np.asarray(p["lu_id"]).astype(np.int64)
    C = lu.shape[0]
    catch, river, lake = masks(lu)
    m = catch.astype(np.float64)
    z_soil = np.asarray(p["z_soil"], dtype=np.float64)
    theta_ac = np.asarray(p["theta_ac"], dtype=np.float64)
    theta_ufc = np.asarray(p["theta_ufc"], dtype=np.float64)
    theta_pwp = np.asarray(p["theta_pwp"], dtype=np.float64)
    dmpv = np.asarray(p["dmpv"], dtype=np.float64)
    c_root = np.asarray(p.get("c_root", np.ones(C)), dtype=np.float64)
    out = {}
    with np.errstate(all="ignore"):
        out["S_pwp_s"] = (z_soil * theta_pwp) * m
        theta_sat = (theta_ac + theta_ufc + theta_pwp) * m
        theta_fc = (theta_ufc + theta_pwp) * m
        lambda_bc = ((np.log(theta_fc / theta_sat) - np.log(theta_pwp / theta_sat)) / (np.log(15850) - np.log(63))) * m
        ha = ((theta_pwp / theta_sat) ** (1.0 / lambda_bc) * (-15850)) * m
        m_bc = ((st["a_bc"] + st["b_bc"] * lambda_bc) / lambda_bc) * m
        n_salv = (st["a_bc"] + st["b_bc"] * lambda_bc) * m
        wfs = (((2 + 3 * lambda_bc) / (1 + 3 * lambda_bc) * ha / 2) * (-10)) * m
        theta_27 = ((ha / (-(10**2.7))) ** lambda_bc * theta_sat) * m
        theta_4 = ((ha / (-(10**4))) ** lambda_bc * theta_sat) * m
        theta_6 = ((ha / (-(10**6))) ** lambda_bc * theta_sat) * m
        clay = (st["clay_max"] * (theta_6 - st["clay_min"]) / 0.3) * m
        clay = np.where(clay < st["clay_min"], st["clay_min"], clay) * m
        z_sc_max = (clay * 700) * m
        mp_drain_area = 1 - np.exp((-1) * (dmpv / 82) ** 0.887) * m
        # root zone (soil.py:258-470)
        rew = np.zeros(C)
        rew = np.where(theta_pwp < st["theta_rew_min"], st["rew_min"], rew) * m
        rew = np.where((theta_pwp >= st["theta_rew_min"]) & (theta_pwp <= st["theta_rew_max"]), theta_pwp / st["theta_rew_max"], rew) * m
        rew = np.where(theta_pwp > st["theta_rew_max"], st["rew_max"], rew) * m
        z_evap = ((rew / st["rew_max"]) * st["z_evap_max"]) * m
        tew = ((theta_fc - 0.5 * theta_pwp) * z_evap) * m
        z_root = np.zeros(C)
        rd = L["RDLU"]
        cc_list = (0, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 98, 31, 32, 33, 40, 41, 50, 60, 98)
        for i in range(0, 61):
            if i in cc_list:
                z_root = np.where(lu == i, rd[_row_no(rd[:, 0], i), 1], z_root) * m
        z_root = np.where(river | lake, 0, z_root)
        z_root = np.where(np.isin(lu, (10, 11, 12, 15, 16, 17)), 1500, z_root) * m
        z_root = np.where(lu == 100, 300, z_root) * m
        z_root = np.where(z_root >= z_soil, st["zroot_to_zsoil_max"] * z_soil, z_root) * m
        z_root = z_root * c_root
        z_root = np.where((lu >= 500) & (lu < 600), 200, z_root) * m
        z_root = np.where(z_root < z_soil, z_root, z_soil * 0.9)
        out.update(
            theta_sat=theta_sat, theta_fc=theta_fc, lambda_bc=lambda_bc, ha=ha, m_bc=m_bc, n_salv=n_salv, wfs=wfs,
            theta_27=theta_27, theta_4=theta_4, theta_6=theta_6, clay=clay, z_sc_max=z_sc_max,
            mp_drain_area=mp_drain_area, rew=rew, z_evap=z_evap, tew=tew, z_root=z_root,
            S_ac_rz=(theta_ac * z_root) * m, S_ufc_rz=(theta_ufc * z_root) * m, S_pwp_rz=(theta_pwp * z_root) * m,
            S_ac_ss=(theta_ac * (z_soil - z_root)) * m, S_ufc_ss=(theta_ufc * (z_soil - z_root)) * m,
            S_pwp_ss=(theta_pwp * (z_soil - z_root)) * m,
            maskCatch=m, maskRiverLake=(river | lake).astype(np.float64),
        )
        if lateral:
            ml = L["MLMS"]
            slope_per = (np.asarray(slope, dtype=np.float64) * 100).astype(np.int64)
            v = np.zeros((8, C))
            for i in range(1, int(slope_per.max()) + 1):
                r = _row_no(ml[:, 0], i)
                sel = slope_per == i
                for k in range(8):  # v_mp_layer_(k+1) <- lut column 8-k, m/h -> mm/h
                    v[k] = np.where(sel, ml[r, 8 - k] * 1000, v[k]) * m
            for k in range(8):
                out["v_mp_layer_%d" % (k + 1)] = v[k] * m
    return out


def initial_conditions(p, d, theta_rz0, theta_ss0):
    """roger/core/soil.py:742-1010 + surface.py:398-427 initial storages from initial theta."""
    m = d["maskCatch"]
    theta_pwp = np.asarray(p["theta_pwp"], dtype=np.float64)
    theta_ufc = np.asarray(p["theta_ufc"], dtype=np.float64)
    z_soil = np.asarray(p["z_soil"], dtype=np.float64)
    ks = np.asarray(p["ks"], dtype=np.float64)
    theta_fc, theta_sat, z_root = d["theta_fc"], d["theta_sat"], d["z_root"]
    out = {}
    with np.errstate(all="ignore"):
        for tag, th0, depth, S_pwp in (("rz", theta_rz0, z_root, d["S_pwp_rz"]), ("ss", theta_ss0, z_soil - z_root, d["S_pwp_ss"])):
            th0 = np.asarray(th0, dtype=np.float64)
            fp = np.zeros_like(th0)
            fp = np.where(th0 > theta_pwp, th0 - theta_pwp, fp) * m
            fp = np.where(th0 <= theta_pwp, 0, fp) * m
            fp = np.where(fp >= theta_ufc, theta_ufc, fp) * m
            lp = np.zeros_like(th0)
            lp = np.where(th0 > theta_fc, th0 - theta_fc, lp) * m
            lp = np.where(th0 <= theta_fc, 0, lp) * m
            S_fp = (fp * depth) * m
            S_lp = (lp * depth) * m
            S = (S_pwp + S_fp + S_lp) * m
            theta = ((S_fp + S_lp) / depth + theta_pwp) * m
            out["S_fp_" + tag], out["S_lp_" + tag], out["S_" + tag], out["theta_" + tag] = S_fp, S_lp, S, theta
            out["k_" + tag] = (ks / (1 + (theta / theta_sat) ** (-d["m_bc"]))) * m
            out["h_" + tag] = (d["ha"] / ((theta / theta_sat) ** (1 / d["lambda_bc"]))) * m
        out["S_s"] = (out["S_rz"] + out["S_ss"]) * m
        out["theta"] = (out["S_s"] / z_soil) * m
        out["S_sur"] = np.zeros_like(m)
        out["S"] = out["S_sur"] + out["S_s"] * m
    return out
