"""Daily hydrologic forcing of the offline StorAge-selection transport, aggregated from an online SVAT run
(the data flow of the reference's test/make_data_for_svat_transport.py:279-300: daily sums of the fluxes,
end-of-day storages, index 0 = initial state)."""
import numpy as np

SUMS = {  # transport forcing name -> SVAT flux fields summed over the day
    "PREC": ("prec",), "INF_MAT_RZ": ("inf_mat_rz",), "INF_PF_RZ": ("inf_mp_rz", "inf_sc_rz"), "INF_PF_SS": ("inf_ss",),
    "TRANSP": ("transp",), "EVAP_SOIL": ("evap_soil",), "CPR_RZ": ("cpr_rz",), "Q_RZ": ("q_rz",), "Q_SS": ("q_ss",),
}
STATES = {"S_RZ": "S_rz", "S_SS": "S_ss", "S_SNOW": "S_snow"}


def daily_fluxes_from_svat(model, ndays):
    """Runs `model` (roger_b200.svat.SvatModel, already set up) for ndays and returns dict name -> (C, ndays+1)."""
    C = model.C
    out = {k: np.zeros((C, ndays + 1)) for k in list(SUMS) + list(STATES)}
    for k, f in STATES.items():
        out[k][:, 0] = model.get(f)
    day = 1
    acc = {k: np.zeros(C) for k in SUMS}
    while day <= ndays:
        sc = model.step(write_fluxes=True)
        for k, fields in SUMS.items():
            for f in fields:
                acc[k] += model.get(f)
        if sc.time % 86400 == 0:
            for k in SUMS:
                out[k][:, day] = acc[k]
                acc[k][:] = 0
            for k, f in STATES.items():
                out[k][:, day] = model.get(f)
            day += 1
    for k in ("S_pwp_rz", "S_pwp_ss"):
        out[k] = model.get(k)
    return out
