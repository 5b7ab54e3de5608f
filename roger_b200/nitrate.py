"""Host side of the nitrate transport (settings.enable_nitrate): the parts of roger/core/nitrate.py that are per-cell scalars
of forcing statistics -- the soil temperature -- and the driver that feeds the engine's nitrogen cycle.

`NitrateModel` mirrors the offline nitrate setup (examples/plot_scale/svat_crop_nitrate_tutorial/svat_crop_nitrate.py without
the crop redistribution: transport.svat_transport_model_deterministic with settings.enable_nitrate, roger/core/transport.py:948-991):
anion transport of the nitrate mass (msa_rz, msa_ss), the soil nitrogen cycle on the age-ranked mineral nitrogen (Nmin_rz, Nmin_ss),
ageing of all four.  The age-array work runs in the engine (CUDA in the product, the CPU oracle in tests)."""
import numpy as np

from . import sas

PARAMS = ["kfix_rz", "kmin_rz", "kmin_ss", "km_nit_rz", "dmax_nit_rz", "kngl_rz", "km_denit_rz", "dmax_denit_rz", "km_denit_ss",
          "dmax_denit_ss", "z_soil", "S_sat_rz", "S_sat_ss"]  # include/roger_b200_sas.h: roger_nitrate_param
FORCING = ["temp_soil", "z_root", "lu_id", "nup"]                # roger_nitrate_forcing
OUTPUTS = ["nit_s", "denit_s", "min_s", "ngas_s", "nfix_s", "nh4_up", "Nmin_s"]  # roger_nitrate_output
STATE = sas.STATE + ["Nmin_rz", "Nmin_ss"]
NITRATE = 2  # roger_sas_settings.solute


def soil_temperature(TA, itt, nitt, doy, phi_soil_temp, damp_soil_temp, z_soil, S_s, S_sat_rz, S_sat_ss, mask, pi):
    """nitrate.py:6-66 (calc_soil_temperature_kernel1 / 2) and :717-731: annual mean and amplitude of the air temperature from the
    coming 364 days (or, near the end of the series, from the last 364), damped and phase-shifted with depth."""
    TA = np.asarray(TA, dtype=np.float64)
    if itt + 364 < nitt:
        ta_year = np.mean(TA[itt:itt + 364])
        a_year = 2 * np.mean(np.abs(TA[itt:itt + 364] - ta_year), axis=-1)
    else:
        ta_year = np.mean(TA[nitt - 364:])
        a_year = 2 * np.mean(np.abs(TA[np.newaxis, nitt - 364:] - ta_year), axis=-1)[0]
    rel = damp_soil_temp * (S_s / (S_sat_rz + S_sat_ss))
    return (ta_year
            + a_year
            * np.sin((2 * pi) * (doy / 365) - (2 * pi) * (phi_soil_temp / 365) / 2 - ((0.5 * (z_soil / 1000)) / rel))
            * np.exp((-0.5 * (z_soil / 1000)) / rel)
            * mask)


class NitrateModel(sas.SasModel):
    def __init__(self, engine_cls, n_cells, ages, daily, sas_spec, params, anion_parameters, nitt=None, pi=3.14159265358979323846264338327950588, dx=1.0, dy=1.0,
                 zroot_to_zsoil_max=0.7, **kw):
        """daily: the anion inputs plus TA, DOY (ndays + 1 values), Z_ROOT, LU_ID, NUP (n_cells, ndays + 1); params: dict of the
        PARAMS planes plus phi_soil_temp, damp_soil_temp"""
        super().__init__(engine_cls, n_cells, ages, daily, sas_spec, solute=NITRATE, anion_parameters=anion_parameters, **kw)
        self.params = {k: np.asarray(v, dtype=np.float64).reshape(-1) for k, v in params.items()}
        self.nitt = int(nitt) if nitt is not None else int(np.asarray(daily["TA"]).shape[-1])
        self.pi, self.dx, self.dy, self.zmax = float(pi), float(dx), float(dy), float(zroot_to_zsoil_max)
        self.temp_soil = np.zeros(self.C)
        self.temp_soil_given = None  # optional (ndays + 1, n_cells) soil temperatures to use instead of soil_temperature()

    def setup(self):
        solute = self.solute
        self.solute = sas.ANION  # the anion part of SasModel.setup (parameters, zero solute)
        try:
            super().setup()
        finally:
            self.solute = solute
        block = np.stack([self.params[k] for k in PARAMS])
        self.engine.set_nitrate_parameters(block, self.dx, self.dy, self.zmax, 24.0)

    def _set_forcing(self):
        solute = self.solute
        self.solute = sas.ANION
        try:
            f = super()._set_forcing()
        finally:
            self.solute = solute
        d, t, p = self.daily, self.itt, self.params
        S_s = d["S_RZ"][:, t] + d["S_SS"][:, t]
        self.temp_soil = soil_temperature(d["TA"], t, self.nitt, float(d["DOY"][t]), p["phi_soil_temp"], p["damp_soil_temp"], p["z_soil"], S_s,
                                          p["S_sat_rz"], p["S_sat_ss"], self.mask, self.pi)
        self.temp_soil_host = self.temp_soil
        if self.temp_soil_given is not None:
            self.temp_soil = np.asarray(self.temp_soil_given[t], dtype=np.float64)
        nf = np.stack([self.temp_soil, d["Z_ROOT"][:, t], d["LU_ID"][:, t], d["NUP"][:, t]])
        self.engine.set_nitrate_forcing(nf)
        return f

    def nitrate_outputs(self):
        o = self.engine.download_nitrate_outputs()
        return {n: o[k] for k, n in enumerate(OUTPUTS)}

    def state(self, name):
        if name in ("Nmin_rz", "Nmin_ss"):
            return self.engine.download_state_index(4 + ("Nmin_rz", "Nmin_ss").index(name))
        return super().state(name)
