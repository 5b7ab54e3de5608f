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
    # RATE diagnostics of the engine: the flux planes are summed over the day's steps where the state lives (on the device
    # for the CUDA engine); one read-back per day
    names = sorted({f for fields in SUMS.values() for f in fields})
    model.engine.set_rate_fields(names)
    while day <= ndays:
        sc = model.step(write_fluxes=False)  # the accumulators themselves make the step store the flux planes they read -- and only those
        if sc.time % 86400 == 0:
            rates = model.engine.download_rates(reset=True)
            for k, fields in SUMS.items():
                acc = np.zeros(C)
                for f in fields:
                    acc += rates[f]
                out[k][:, day] = acc
            for k, f in STATES.items():
                out[k][:, day] = model.get(f)
            day += 1
    model.engine.set_rate_fields([])
    for k in ("S_pwp_rz", "S_pwp_ss"):
        out[k] = model.get(k)
    return out
