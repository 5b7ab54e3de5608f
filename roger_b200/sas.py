"""Host-side driver of the offline oxygen-18 StorAge-selection transport: the parts of the reference's
SVATOXYGEN18 setup that surround the per-cell transport step (benchmarks/SVATOXYGEN18_benchmark.py:150-444,
roger/models/svat_oxygen18): initial age distributions, daily forcing hand-over with the snow-pack isotope
mixing of `set_forcing`.  The step itself is delegated to an engine (CUDA in the product, the CPU oracle in tests).
"""
import ctypes as C

import numpy as np

FORCING = ["inf_mat_rz", "inf_pf_rz", "inf_pf_ss", "evap_soil", "transp", "q_rz", "q_ss", "cpr_rz", "C_in"]
PROCESSES = ["evap_soil", "transp", "q_rz", "q_ss", "cpr_rz"]
OUTPUTS = [
    "C_inf_mat_rz", "C_inf_pf_rz", "C_inf_pf_ss", "C_evap_soil", "C_transp", "C_q_rz", "C_q_ss", "C_cpr_rz",
    "C_iso_inf_mat_rz", "C_iso_inf_pf_rz", "C_iso_inf_pf_ss", "C_iso_evap_soil", "C_iso_transp", "C_iso_q_rz", "C_iso_q_ss",
    "C_iso_cpr_rz", "C_rz", "C_ss", "C_s", "C_iso_rz", "C_iso_ss", "C_iso_s",
]
STATE = ["sa_rz", "sa_ss", "msa_rz", "msa_ss"]
STATISTICS = [
    "tt10_transp", "tt25_transp", "tt50_transp", "tt75_transp", "tt90_transp", "ttavg_transp",
    "tt10_q_ss", "tt25_q_ss", "tt50_q_ss", "tt75_q_ss", "tt90_q_ss", "ttavg_q_ss",
    "rt25_rz", "rt50_rz", "rt75_rz", "rtavg_rz", "rt25_ss", "rt50_ss", "rt75_ss", "rtavg_ss",
    "rt10_s", "rt25_s", "rt50_s", "rt75_s", "rt90_s", "rtavg_s", "dS_num_error", "dC_num_error", "sanity_ok",
]
VSMOW_conc18O = 2005.2e-6  # roger/settings.py
D18O_MIN, D18O_MAX = -20.0, 0.0


class SasSettings(C.Structure):
    _fields_ = [("ages", C.c_int32), ("substeps", C.c_int32), ("VSMOW_conc18O", C.c_double), ("d18O_min", C.c_double),
                ("d18O_max", C.c_double), ("store_tt", C.c_int32), ("age_statistics", C.c_int32), ("solute", C.c_int32),
                ("solver", C.c_int32)]


ISOTOPE, ANION = 0, 1
VIRTUALTRACER = 3  # roger_sas_settings.solute: the anion model whose soil evaporation carries the tracer (settings.enable_virtualtracer)
DETERMINISTIC, EULER, RK4 = 0, 1, 2  # settings.sas_solver of the reference (roger/settings.py); RK4: isotope branch


def default_settings(ages, substeps=6, store_tt=False, age_statistics=False, solute=ISOTOPE, solver=DETERMINISTIC):
    return SasSettings(int(ages), int(substeps), VSMOW_conc18O, D18O_MIN, D18O_MAX, int(store_tt), int(age_statistics), int(solute), int(solver))


def delta_to_conc(delta):
    """roger/core/transport.py:315-325"""
    delta = np.asarray(delta, dtype=np.float64)
    return VSMOW_conc18O * (delta / 1000.0 + 1.0) / (1.0 + (delta / 1000.0 + 1.0) * VSMOW_conc18O)


def conc_to_delta(conc):
    """roger/core/transport.py:328-340"""
    conc = np.asarray(conc, dtype=np.float64)
    with np.errstate(all="ignore"):
        d = 1000.0 * (conc / (VSMOW_conc18O * (1.0 - conc)) - 1.0)
        return np.where((d < D18O_MIN) | (d > D18O_MAX), np.nan, d)


def sas_parameter_block(n_cells, spec, processes=None):
    """spec: process -> tuple (id, p1, p2, ...) or (n_cells, 8) array.  Returns (len(processes), n_cells, 8), by default for
    the five fluxes of PROCESSES."""
    processes = PROCESSES if processes is None else list(processes)
    out = np.zeros((len(processes), n_cells, 8))
    for k, proc in enumerate(processes):
        v = spec[proc]
        a = np.asarray(v, dtype=np.float64)
        if a.ndim == 1:
            out[k, :, : a.size] = a[None, :]
        else:
            out[k] = a.reshape(n_cells, 8)
    return out


def initial_age_distribution(S, ages):
    """benchmarks/SVATOXYGEN18_benchmark.py:172-190: sa[1:] = diff(linspace(0, S, ages)), sa[0] = 0."""
    S = np.asarray(S, dtype=np.float64).reshape(-1)
    lin = np.linspace(np.zeros_like(S), S, ages, axis=-1)
    sa = np.zeros((S.size, ages))
    sa[:, 1:] = np.diff(lin, axis=-1)
    return sa


REDISTRIBUTION = ["re_rg", "re_rl", "C_re_rg", "C_re_rl", "M_re_rg", "M_re_rl"]  # enum roger_crop_output


class SasModel:
    def __init__(self, engine_cls, n_cells, ages, daily, sas_spec, d18O_soil=-10.0, mask=None, substeps=6, store_tt=False,
                 engine_kwargs=None, age_statistics=False, solute=ISOTOPE, anion_parameters=None, solver=DETERMINISTIC, crop=False):
        """daily: dict of (n_cells, ndays + 1) arrays PREC, INF_MAT_RZ, INF_PF_RZ, INF_PF_SS, TRANSP, EVAP_SOIL, CPR_RZ,
        Q_RZ, Q_SS, S_RZ, S_SS, S_SNOW, C_ISO_IN (index 0 = initial state, as in the reference's SVAT.nc).
        crop: the crop variant of the offline model (settings.enable_crop_phenology, svat_crop_transport_model_deterministic): `daily`
        then carries RE_RG, RE_RL (the redistribution fluxes of the SVAT-CROP run) and `sas_spec` the rows re_rg, re_rl."""
        self.crop = bool(crop)
        self.C = int(n_cells)
        self.ages = int(ages)
        self.daily = {k: np.asarray(v, dtype=np.float64) for k, v in daily.items()}
        self.solute = int(solute)
        # anion transport (roger/models/svat_bromide): dict alpha_transp, alpha_q, lu_id, S_sat_rz of n_cells values each
        self.anion_parameters = anion_parameters
        self.settings = default_settings(ages, substeps, store_tt, age_statistics, solute, solver)
        if self.solute == VIRTUALTRACER:  # the host side is the anion model's; the engine's settings carry the 3
            self.solute = ANION
        self.engine = engine_cls(self.C, self.settings, **(engine_kwargs or {}))
        self.mask = np.ones(self.C) if mask is None else np.asarray(mask, dtype=np.float64)
        self.sas_params = sas_parameter_block(self.C, sas_spec)
        self.sas_spec = sas_spec
        self.d18O_soil = d18O_soil
        self.itt = 0
        self.C_IN = delta_to_conc(self.daily["C_ISO_IN"]) if self.solute == ISOTOPE else self.daily["C_IN"]
        self.C_snow = np.full((self.C, 2), 0.0)  # vs.C_snow[taum1, tau], initial 0 (roger/variables.py)
        self.S_snow_m1 = self.daily["S_SNOW"][:, 0].copy()
        self.prec_m1 = np.zeros(self.C)
        self.C_in = np.zeros(self.C)

    def setup(self):
        d = self.daily
        sa_rz = initial_age_distribution(d["S_RZ"][:, 0], self.ages)
        sa_ss = initial_age_distribution(d["S_SS"][:, 0], self.ages)
        c0 = float(delta_to_conc(self.d18O_soil))
        msa = np.full((self.C, self.ages), c0)
        msa[:, 0] = 0
        if self.solute == ANION:  # no solute in the soil initially (svat_bromide.py:235-266)
            msa[:] = 0
            p = self.anion_parameters or {}
            one = np.ones(self.C)
            self.engine.set_anion_parameters(p.get("alpha_transp", one), p.get("alpha_q", one), p.get("lu_id", 0 * one),
                                             p.get("S_sat_rz", np.inf * one))
        self.engine.set_parameters(self.mask, self.sas_params)
        if self.crop:
            rows = sas_parameter_block(self.C, self.sas_spec, processes=("re_rg", "re_rl"))
            self.engine.set_crop_parameters(rows[0], rows[1])
        self.engine.upload_state("sa_rz", sa_rz)
        self.engine.upload_state("sa_ss", sa_ss)
        self.engine.upload_state("msa_rz", msa)
        self.engine.upload_state("msa_ss", msa.copy())
        if self.settings.age_statistics:
            # vs.C_s[taum1] of the first step: the storage concentration of the initial state (what the reference's warmup
            # leaves there, roger/roger.py:497-521)
            c_s = np.where(np.sum(sa_rz + sa_ss, axis=-1) > 0, c0, 0.0)
            self.engine.set_balance_reference(c_s)

    def _set_forcing(self):
        """benchmarks/SVATOXYGEN18_benchmark.py:385-444 (per-cell vectors); roger/models/svat_bromide/svat_bromide.py:325-348."""
        d, t = self.daily, self.itt
        if self.solute == ANION:
            self.C_in = self.C_IN[:, t].copy()
            f = np.empty((len(FORCING), self.C))
            f[0], f[1], f[2] = d["INF_MAT_RZ"][:, t], d["INF_PF_RZ"][:, t], d["INF_PF_SS"][:, t]
            f[3], f[4], f[5], f[6], f[7] = d["EVAP_SOIL"][:, t], d["TRANSP"][:, t], d["Q_RZ"][:, t], d["Q_SS"][:, t], d["CPR_RZ"][:, t]
            f[8] = self.C_in
            return f
        prec, S_snow = d["PREC"][:, t], d["S_SNOW"][:, t]
        C_in = self.C_IN[:, t].copy()
        cs_m1, cs = self.C_snow[:, 0], self.C_snow[:, 1]
        with np.errstate(all="ignore"):
            mixed = (prec / (prec + S_snow)) * C_in + (S_snow / (prec + S_snow)) * cs_m1
            new = np.where(S_snow > 0, np.where(np.isnan(cs), C_in, mixed), np.nan)
            new = np.where(S_snow <= 0, np.nan, new)
        self.C_snow[:, 1] = new
        C_in = np.where(np.isfinite(cs_m1), cs_m1, np.where(prec > 0, self.C_IN[:, t], 0))
        self.C_in = C_in
        f = np.empty((len(FORCING), self.C))
        f[0], f[1], f[2] = d["INF_MAT_RZ"][:, t], d["INF_PF_RZ"][:, t], d["INF_PF_SS"][:, t]
        f[3], f[4], f[5], f[6], f[7] = d["EVAP_SOIL"][:, t], d["TRANSP"][:, t], d["Q_RZ"][:, t], d["Q_SS"][:, t], d["CPR_RZ"][:, t]
        f[8] = C_in
        return f

    def step(self):
        """RogerSetup.step, offline-transport branch (roger/roger.py:467-485)."""
        self.itt += 1
        f = self._set_forcing()
        if self.crop:  # svat_crop_nitrate.py:603-604
            self.engine.set_redistribution(self.daily["RE_RG"][:, self.itt], self.daily["RE_RL"][:, self.itt])
        self.engine.step(f)
        self.C_snow[:, 0] = self.C_snow[:, 1]  # after_timestep
        return self.itt

    def outputs(self):
        o = self.engine.download_outputs()
        return {n: o[k] for k, n in enumerate(OUTPUTS)}

    def redistribution(self):
        o = self.engine.download_redistribution()
        return {n: o[k] for k, n in enumerate(REDISTRIBUTION)}

    def statistics(self):
        o = self.engine.download_statistics()
        return {n: o[k] for k, n in enumerate(STATISTICS)}

    def state(self, name):
        return self.engine.download_state(name)

    def tt(self, process):
        return self.engine.download_tt(process)
