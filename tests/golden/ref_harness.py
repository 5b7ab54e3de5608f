"""Runs the UNMODIFIED reference (read-only at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY -- used by the ``make_*_golden.py`` scripts to produce the
committed fixtures under ``tests/golden/``.  Nothing on the GPU box imports this file
(`/root/reference` does not exist there).

The reference's NumPy backend is imported as-is; the only shims are empty stand-ins
for the three IO packages that are absent offline (h5netcdf, h5py, cftime) -- they are
never called because forcing is fed from memory and ``diskless_mode`` is on
(roger/diagnostics/base.py:85, roger/restart.py:134).
"""
import sys
import types
import numpy as onp

REFERENCE_ROOT = "/root/reference"


def import_reference(float_type="float64"):
    for name in ("h5netcdf", "h5py", "cftime"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.File = None
            m.date2num = None
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import roger  # noqa

    if not getattr(import_reference, "_done", False):
        roger.runtime_settings.update(
            backend="numpy",
            float_type=float_type,
            int_type="int64" if float_type == "float64" else "int32",
            diskless_mode=True,
            loglevel="error",
        )
        import_reference._done = True
    return roger


def make_svat_setup(case):
    """Build a RogerSetup subclass mirroring benchmarks/SVAT_benchmark.py but with in-memory,
    optionally per-cell forcing and heterogeneous parameters.

    case: dict with nx, ny, ndays, PREC/TA/PET (1-D 10-min series), YEAR/MONTH/DOY, params dict of
    (nx, ny) arrays, optional prec_weight/ta_offset/pet_weight (nx, ny), lateral (bool -> oneD).
    """
    roger = import_reference()
    from roger import RogerSetup, roger_routine, roger_kernel, KernelOutput
    from roger.variables import allocate
    from roger.core.operators import numpy as npx, update, at
    from roger.core.surface import calc_parameters_surface_kernel
    import roger.lookuptables as lut

    nx, ny, ndays = case["nx"], case["ny"], case["ndays"]
    P = case["params"]
    lateral = bool(case.get("lateral", False))
    # groundwater boundary condition: case["Z_GW"] (nx, ny, ndays + 1) is the depth of the water table (m), column 0 the initial
    # state (examples/hillslope_scale/oneD_with_groundwater_boundary_condition_distributed_tutorial/oneD.py:266, :317-361)
    gwb = case.get("Z_GW") is not None

    copy_tau = [
        "ta", "z_root", "ground_cover", "S_sur", "S_int_top", "S_int_ground", "S_dep", "S_snow", "swe",
        "S_rz", "S_ss", "S_s", "S", "z_sat", "z_wf", "z_wf_t0", "z_wf_t1", "y_mp", "y_sc", "theta_rz",
        "theta_ss", "theta", "k_rz", "k_ss", "k", "h_rz", "h_ss", "h", "z0", "prec",
    ]
    if gwb:
        copy_tau = copy_tau + ["z_gw", "S_vad"]

    class GoldenSetup(RogerSetup):
        @roger_routine
        def set_settings(self, state):
            s = state.settings
            s.identifier = "GOLDEN"
            s.nx, s.ny = nx, ny
            s.runlen = 24 * 60 * 60 * ndays
            s.nitt_forc = len(case["PREC"])
            s.dx = 1
            s.dy = 1
            s.x_origin = 0.0
            s.y_origin = 0.0
            s.time_origin = "2017-12-31 00:00:00"
            s.enable_groundwater_boundary = gwb
            s.enable_adaptive_time_stepping = True
            if lateral:
                s.enable_lateral_flow = True

        @roger_routine
        def read_data(self, state):
            pass

        @roger_routine(dist_safe=False, local_variables=["x", "y"])
        def set_grid(self, state):
            vs = state.variables
            dx = allocate(state.dimensions, ("x"))
            dx = update(dx, at[:], 1)
            dy = allocate(state.dimensions, ("y"))
            dy = update(dy, at[:], 1)
            vs.x = update(vs.x, at[3:-2], npx.cumsum(dx[3:-2]))
            vs.y = update(vs.y, at[3:-2], npx.cumsum(dy[3:-2]))

        @roger_routine
        def set_look_up_tables(self, state):
            vs = state.variables
            vs.lut_ilu = update(vs.lut_ilu, at[:, :], lut.ARR_ILU)
            vs.lut_gc = update(vs.lut_gc, at[:, :], lut.ARR_GC)
            vs.lut_gcm = update(vs.lut_gcm, at[:, :], lut.ARR_GCM)
            vs.lut_is = update(vs.lut_is, at[:, :], lut.ARR_IS)
            vs.lut_rdlu = update(vs.lut_rdlu, at[:, :], lut.ARR_RDLU)
            if lateral:
                vs.lut_mlms = update(vs.lut_mlms, at[:, :], lut.ARR_MLMS)

        @roger_routine
        def set_topography(self, state):
            pass

        @roger_routine
        def set_parameters_setup(self, state):
            vs = state.variables
            for name, arr in P.items():
                cur = getattr(vs, name)
                setattr(vs, name, update(cur, at[2:-2, 2:-2], onp.asarray(arr).astype(cur.dtype)))
            if lateral:
                vs.slope_per = update(vs.slope_per, at[2:-2, 2:-2], vs.slope[2:-2, 2:-2] * 100)

        @roger_routine
        def set_parameters(self, state):
            vs = state.variables
            if (vs.month[vs.tau] != vs.month[vs.taum1]) & (vs.itt > 1):
                vs.update(calc_parameters_surface_kernel(state))

        @roger_routine
        def set_initial_conditions_setup(self, state):
            pass

        @roger_routine
        def set_initial_conditions(self, state):
            vs = state.variables
            th_rz = onp.asarray(case["theta_rz0"])[:, :, None]
            th_ss = onp.asarray(case["theta_ss0"])[:, :, None]
            vs.theta_rz = update(vs.theta_rz, at[2:-2, 2:-2, : vs.taup1], th_rz)
            vs.theta_ss = update(vs.theta_ss, at[2:-2, 2:-2, : vs.taup1], th_ss)
            if gwb:
                vs.z_gw = update(vs.z_gw, at[2:-2, 2:-2, : vs.taup1], onp.asarray(case["Z_GW"])[:, :, 0][:, :, None])

        @roger_routine
        def set_boundary_conditions_setup(self, state):
            pass

        @roger_routine
        def set_boundary_conditions(self, state):
            pass

        @roger_routine(dist_safe=False, local_variables=["PREC", "TA", "PET", "YEAR", "MONTH", "DOY"])
        def set_forcing_setup(self, state):
            vs = state.variables
            vs.PREC = update(vs.PREC, at[:], case["PREC"])
            vs.TA = update(vs.TA, at[:], case["TA"])
            vs.PET = update(vs.PET, at[:], case["PET"])
            vs.YEAR = update(vs.YEAR, at[:], case["YEAR"])
            vs.MONTH = update(vs.MONTH, at[:], case["MONTH"])
            vs.DOY = update(vs.DOY, at[:], case["DOY"])

        @roger_routine
        def set_forcing(self, state):
            vs = state.variables
            condt = vs.time % (24 * 60 * 60) == 0
            if condt:
                vs.itt_day = 0
                vs.year = update(vs.year, at[1], vs.YEAR[vs.itt_forc])
                vs.month = update(vs.month, at[1], vs.MONTH[vs.itt_forc])
                vs.doy = update(vs.doy, at[1], vs.DOY[vs.itt_forc])
                sl = slice(int(vs.itt_forc), int(vs.itt_forc) + 6 * 24)
                pw = case.get("prec_weight")
                to = case.get("ta_offset")
                ew = case.get("pet_weight")
                prec = vs.PREC[npx.newaxis, npx.newaxis, sl]
                ta = vs.TA[npx.newaxis, npx.newaxis, sl]
                pet = vs.PET[npx.newaxis, npx.newaxis, sl]
                if pw is not None:
                    # Eberbaechle pattern (examples/catchment_scale/eberbaechle/svat_distributed/svat.py:258-297)
                    prec = prec * onp.asarray(pw)[:, :, None]
                    ta = ta + onp.asarray(to)[:, :, None]
                    pet = pet * onp.asarray(ew)[:, :, None]
                    vs.prec_day = update(vs.prec_day, at[2:-2, 2:-2, :], prec)
                    vs.ta_day = update(vs.ta_day, at[2:-2, 2:-2, :], ta)
                    vs.pet_day = update(vs.pet_day, at[2:-2, 2:-2, :], pet)
                else:
                    vs.prec_day = update(vs.prec_day, at[:, :, :], prec)
                    vs.ta_day = update(vs.ta_day, at[:, :, :], ta)
                    vs.pet_day = update(vs.pet_day, at[:, :, :], pet)
                if gwb:
                    day = min(int(vs.itt_forc) // (6 * 24) + 1, onp.asarray(case["Z_GW"]).shape[2] - 1)
                    vs.z_gw = update(vs.z_gw, at[2:-2, 2:-2, vs.tau], onp.asarray(case["Z_GW"])[:, :, day])
                vs.itt_forc = vs.itt_forc + 6 * 24

        @roger_routine
        def set_diagnostics(self, state):
            pass

        @roger_routine
        def after_timestep(self, state):
            vs = state.variables
            vs.update(after_timestep_kernel(state))

    @roger_kernel
    def after_timestep_kernel(state):
        # same variable list as benchmarks/SVAT_benchmark.py:183-418
        vs = state.variables
        out = {}
        for name in copy_tau:
            arr = getattr(vs, name)
            arr = update(arr, at[2:-2, 2:-2, vs.taum1], arr[2:-2, 2:-2, vs.tau])
            setattr(vs, name, arr)
            out[name] = arr
        if not lateral:
            for name in ("S_fp_rz", "S_lp_rz", "S_fp_ss", "S_lp_ss"):
                arr = getattr(vs, name)
                arr = update(arr, at[2:-2, 2:-2], npx.where((arr > -1e-6) & (arr < 0), 0, arr)[2:-2, 2:-2])
                setattr(vs, name, arr)
                out[name] = arr
        for name in ("event_id", "year", "month", "doy"):
            arr = getattr(vs, name)
            arr = update(arr, at[vs.taum1], arr[vs.tau])
            setattr(vs, name, arr)
            out[name] = arr
        return KernelOutput(**out)

    return GoldenSetup


def snapshot(state, names=None):
    """Interior view of every active (x, y[, t]) float/int variable plus the 0-d scalars."""
    vs = state.variables
    out = {}
    for name, meta in state.var_meta.items():
        if not meta.active:
            continue
        if names is not None and name not in names:
            continue
        dims = meta.dims
        arr = onp.asarray(getattr(vs, name))
        if dims is None or len(dims) == 0:
            out[name] = arr.copy()
        elif tuple(dims[:2]) == ("x", "y"):
            if len(dims) == 2:
                out[name] = arr[2:-2, 2:-2].copy()
            elif len(dims) == 3 and dims[2] == "timesteps":
                out[name] = arr[2:-2, 2:-2, :2].copy()
        elif tuple(dims) == ("timesteps",):
            out[name] = arr.copy()
    return out


def make_o18_setup(case):
    """Offline oxygen-18 StorAge-selection transport, restating benchmarks/SVATOXYGEN18_benchmark.py
    (the multi-cell variant of roger/models/svat_oxygen18) with in-memory daily forcing.

    case: nx, ny, ndays, ages, daily arrays of shape (nx, ny, ndays + 1): PREC, INF_MAT_RZ, INF_PF_RZ,
    INF_PF_SS, TRANSP, EVAP_SOIL, CPR_RZ, Q_RZ, Q_SS, S_RZ, S_SS, S_SNOW, C_ISO_IN;
    (nx, ny) arrays S_pwp_rz, S_pwp_ss, S_sat_rz, S_sat_ss; sas: dict process -> (nx, ny, 8) parameter array.
    """
    import_reference()
    from roger import RogerSetup, roger_routine, roger_kernel, KernelOutput
    from roger.variables import allocate
    from roger.core.operators import numpy as npx, update, at
    from roger.core.transport import delta_to_conc, conc_to_delta

    nx, ny, ndays, ages = case["nx"], case["ny"], case["ndays"], case["ages"]

    class O18Setup(RogerSetup):
        @roger_routine
        def set_settings(self, state):
            s = state.settings
            s.identifier = "GOLDEN_O18"
            s.sas_solver = case.get("sas_solver", "deterministic")
            s.sas_solver_substeps = case.get("substeps", 6)
            if s.sas_solver in ("RK4", "Euler"):  # roger/models/svat_oxygen18/svat_oxygen18.py:_set_settings
                s.h = 1 / s.sas_solver_substeps
            s.nx, s.ny = nx, ny
            s.runlen = 24 * 60 * 60 * ndays
            s.runlen_warmup = 0
            s.nitt = ndays + 1
            s.ages = ages
            s.nages = ages + 1
            s.dx = 1
            s.dy = 1
            s.x_origin = 0.0
            s.y_origin = 0.0
            s.time_origin = "01-01-2022"
            s.enable_offline_transport = True
            s.enable_oxygen18 = True
            s.tm_structure = "power"
            s.enable_age_statistics = bool(case.get("age_statistics", False))
            # groundwater boundary: calculate_storage_selection then runs svat_lbc_transport_model_deterministic
            # (transport.py:995-1040, :3304-3349), whose extra kernel -- capillary rise into the subsoil -- is an empty stub
            # (capillary_rise.py:590-611)
            s.enable_groundwater_boundary = bool(case.get("groundwater_boundary", False))

        @roger_routine
        def read_data(self, state):
            pass

        @roger_routine
        def set_grid(self, state):
            vs = state.variables
            settings = state.settings
            vs.dt_secs = 60 * 60 * 24
            vs.dt = 60 * 60 * 24 / (60 * 60)
            vs.ages = update(vs.ages, at[:], npx.arange(1, settings.nages))
            vs.nages = update(vs.nages, at[:], npx.arange(settings.nages))
            dx = allocate(state.dimensions, ("x"))
            dx = update(dx, at[:], settings.dx)
            dy = allocate(state.dimensions, ("y"))
            dy = update(dy, at[:], settings.dy)
            vs.x = update(vs.x, at[3:-2], settings.x_origin + npx.cumsum(dx[3:-2]))
            vs.y = update(vs.y, at[3:-2], settings.y_origin + npx.cumsum(dy[3:-2]))

        @roger_routine
        def set_look_up_tables(self, state):
            pass

        @roger_routine
        def set_topography(self, state):
            pass

        @roger_routine
        def set_parameters_setup(self, state):
            vs = state.variables
            for k in ("S_pwp_rz", "S_pwp_ss", "S_sat_rz", "S_sat_ss"):
                setattr(vs, k, update(getattr(vs, k), at[2:-2, 2:-2], case[k]))
            for proc, arr in case["sas"].items():
                name = "sas_params_" + proc
                setattr(vs, name, update(getattr(vs, name), at[2:-2, 2:-2, :], arr))

        @roger_routine
        def set_parameters(self, state):
            pass

        @roger_routine
        def set_initial_conditions_setup(self, state):
            vs = state.variables
            vs.S_snow = update(vs.S_snow, at[2:-2, 2:-2, : vs.taup1], case["S_SNOW"][:, :, 0, None])
            vs.S_rz = update(vs.S_rz, at[2:-2, 2:-2, : vs.taup1], case["S_RZ"][:, :, 0, None])
            vs.S_ss = update(vs.S_ss, at[2:-2, 2:-2, : vs.taup1], case["S_SS"][:, :, 0, None])
            vs.S_s = update(vs.S_s, at[2:-2, 2:-2, : vs.taup1], vs.S_rz[2:-2, 2:-2, : vs.taup1] + vs.S_ss[2:-2, 2:-2, : vs.taup1])
            vs.S_rz_init = update(vs.S_rz_init, at[2:-2, 2:-2], vs.S_rz[2:-2, 2:-2, 0])
            vs.S_ss_init = update(vs.S_ss_init, at[2:-2, 2:-2], vs.S_ss[2:-2, 2:-2, 0])

        @roger_routine
        def set_initial_conditions(self, state):
            # benchmarks/SVATOXYGEN18_benchmark.py:172-310
            vs = state.variables
            settings = state.settings
            arr0 = allocate(state.dimensions, ("x", "y"))
            vs.sa_rz = update(vs.sa_rz, at[2:-2, 2:-2, : vs.taup1, 1:],
                              npx.diff(npx.linspace(arr0[2:-2, 2:-2], vs.S_rz[2:-2, 2:-2, vs.tau], settings.ages, axis=-1), axis=-1)[:, :, npx.newaxis, :])
            vs.sa_ss = update(vs.sa_ss, at[2:-2, 2:-2, : vs.taup1, 1:],
                              npx.diff(npx.linspace(arr0[2:-2, 2:-2], vs.S_ss[2:-2, 2:-2, vs.tau], settings.ages, axis=-1), axis=-1)[:, :, npx.newaxis, :])
            vs.SA_rz = update(vs.SA_rz, at[2:-2, 2:-2, :, 1:], npx.cumsum(vs.sa_rz[2:-2, 2:-2, :, :], axis=-1))
            vs.SA_ss = update(vs.SA_ss, at[2:-2, 2:-2, :, 1:], npx.cumsum(vs.sa_rz[2:-2, 2:-2, :, :], axis=-1))
            vs.sa_s = update(vs.sa_s, at[2:-2, 2:-2, :, :], vs.sa_rz[2:-2, 2:-2, :, :] + vs.sa_ss[2:-2, 2:-2, :, :])
            vs.SA_s = update(vs.SA_s, at[2:-2, 2:-2, :, 1:], npx.cumsum(vs.sa_s[2:-2, 2:-2, :, :], axis=-1))
            vs.C_iso_snow = update(vs.C_iso_snow, at[2:-2, 2:-2, : vs.taup1], npx.nan)
            vs.C_iso_rz = update(vs.C_iso_rz, at[2:-2, 2:-2, : vs.taup1], case.get("d18O_soil", -10))
            vs.C_iso_ss = update(vs.C_iso_ss, at[2:-2, 2:-2, : vs.taup1], case.get("d18O_soil", -10))
            vs.C_rz = update(vs.C_rz, at[2:-2, 2:-2, : vs.taup1], delta_to_conc(state, vs.C_iso_rz[2:-2, 2:-2, vs.tau, npx.newaxis]))
            vs.msa_rz = update(vs.msa_rz, at[2:-2, 2:-2, : vs.taup1, :], vs.C_rz[2:-2, 2:-2, : vs.taup1, npx.newaxis])
            vs.msa_rz = update(vs.msa_rz, at[2:-2, 2:-2, : vs.taup1, 0], 0)
            vs.C_ss = update(vs.C_ss, at[2:-2, 2:-2, : vs.taup1], delta_to_conc(state, vs.C_iso_ss[2:-2, 2:-2, vs.tau, npx.newaxis]))
            vs.msa_ss = update(vs.msa_ss, at[2:-2, 2:-2, : vs.taup1, :], vs.C_ss[2:-2, 2:-2, : vs.taup1, npx.newaxis])
            vs.msa_ss = update(vs.msa_ss, at[2:-2, 2:-2, : vs.taup1, 0], 0)

        @roger_routine
        def set_boundary_conditions_setup(self, state):
            pass

        @roger_routine
        def set_boundary_conditions(self, state):
            pass

        @roger_routine
        def set_forcing_setup(self, state):
            vs = state.variables
            for big, key in (("PREC_DIST_DAILY", "PREC"), ("INF_MAT_RZ", "INF_MAT_RZ"), ("INF_PF_RZ", "INF_PF_RZ"),
                             ("INF_PF_SS", "INF_PF_SS"), ("TRANSP", "TRANSP"), ("EVAP_SOIL", "EVAP_SOIL"), ("CPR_RZ", "CPR_RZ"),
                             ("Q_RZ", "Q_RZ"), ("Q_SS", "Q_SS"), ("S_RZ", "S_RZ"), ("S_SS", "S_SS"), ("S_SNOW", "S_SNOW"),
                             ("C_ISO_IN", "C_ISO_IN")):
                setattr(vs, big, update(getattr(vs, big), at[2:-2, 2:-2, :], case[key]))
            vs.S_S = update(vs.S_S, at[2:-2, 2:-2, :], vs.S_RZ[2:-2, 2:-2, :] + vs.S_SS[2:-2, 2:-2, :])
            vs.C_IN = update(vs.C_IN, at[2:-2, 2:-2, :], delta_to_conc(state, vs.C_ISO_IN)[2:-2, 2:-2, :])

        @roger_routine
        def set_forcing(self, state):
            # benchmarks/SVATOXYGEN18_benchmark.py:385-444
            vs = state.variables
            vs.prec = update(vs.prec, at[2:-2, 2:-2, vs.tau], vs.PREC_DIST_DAILY[2:-2, 2:-2, vs.itt])
            vs.inf_mat_rz = update(vs.inf_mat_rz, at[2:-2, 2:-2], vs.INF_MAT_RZ[2:-2, 2:-2, vs.itt])
            vs.inf_pf_rz = update(vs.inf_pf_rz, at[2:-2, 2:-2], vs.INF_PF_RZ[2:-2, 2:-2, vs.itt])
            vs.inf_pf_ss = update(vs.inf_pf_ss, at[2:-2, 2:-2], vs.INF_PF_SS[2:-2, 2:-2, vs.itt])
            vs.transp = update(vs.transp, at[2:-2, 2:-2], vs.TRANSP[2:-2, 2:-2, vs.itt])
            vs.evap_soil = update(vs.evap_soil, at[2:-2, 2:-2], vs.EVAP_SOIL[2:-2, 2:-2, vs.itt])
            vs.cpr_rz = update(vs.cpr_rz, at[2:-2, 2:-2], vs.CPR_RZ[2:-2, 2:-2, vs.itt])
            vs.q_rz = update(vs.q_rz, at[2:-2, 2:-2], vs.Q_RZ[2:-2, 2:-2, vs.itt])
            vs.q_ss = update(vs.q_ss, at[2:-2, 2:-2], vs.Q_SS[2:-2, 2:-2, vs.itt])
            vs.S_rz = update(vs.S_rz, at[2:-2, 2:-2, vs.tau], vs.S_RZ[2:-2, 2:-2, vs.itt])
            vs.S_ss = update(vs.S_ss, at[2:-2, 2:-2, vs.tau], vs.S_SS[2:-2, 2:-2, vs.itt])
            vs.S_s = update(vs.S_s, at[2:-2, 2:-2, vs.tau], vs.S_rz[2:-2, 2:-2, vs.tau] + vs.S_ss[2:-2, 2:-2, vs.tau])
            vs.S_snow = update(vs.S_snow, at[2:-2, 2:-2, vs.tau], vs.S_SNOW[2:-2, 2:-2, vs.itt])
            vs.C_in = update(vs.C_in, at[2:-2, 2:-2], vs.C_IN[2:-2, 2:-2, vs.itt])
            vs.C_snow = update(
                vs.C_snow, at[2:-2, 2:-2, vs.tau],
                npx.where(
                    vs.S_snow[2:-2, 2:-2, vs.tau] > 0,
                    npx.where(
                        npx.isnan(vs.C_snow[2:-2, 2:-2, vs.tau]),
                        vs.C_in[2:-2, 2:-2],
                        (vs.prec[2:-2, 2:-2, vs.tau] / (vs.prec[2:-2, 2:-2, vs.tau] + vs.S_snow[2:-2, 2:-2, vs.tau])) * vs.C_in[2:-2, 2:-2]
                        + (vs.S_snow[2:-2, 2:-2, vs.tau] / (vs.prec[2:-2, 2:-2, vs.tau] + vs.S_snow[2:-2, 2:-2, vs.tau])) * vs.C_snow[2:-2, 2:-2, vs.taum1],
                    ),
                    npx.nan,
                ),
            )
            vs.C_snow = update(vs.C_snow, at[2:-2, 2:-2, vs.tau], npx.where(vs.S_snow[2:-2, 2:-2, vs.tau] <= 0, npx.nan, vs.C_snow[2:-2, 2:-2, vs.tau]))
            vs.C_iso_snow = update(vs.C_iso_snow, at[2:-2, 2:-2, vs.tau], conc_to_delta(state, vs.C_snow[2:-2, 2:-2, vs.tau]))
            vs.C_in = update(
                vs.C_in, at[2:-2, 2:-2],
                npx.where(npx.isfinite(vs.C_snow[2:-2, 2:-2, vs.taum1]), vs.C_snow[2:-2, 2:-2, vs.taum1],
                          npx.where(vs.prec[2:-2, 2:-2, vs.tau] > 0, vs.C_IN[2:-2, 2:-2, vs.itt], 0)),
            )
            vs.C_iso_in = update(vs.C_iso_in, at[2:-2, 2:-2], conc_to_delta(state, vs.C_in[2:-2, 2:-2]))

        @roger_routine
        def set_diagnostics(self, state):
            pass

        @roger_routine
        def after_timestep(self, state):
            vs = state.variables
            vs.update(after_timestep_kernel(state))

    @roger_kernel
    def after_timestep_kernel(state):
        vs = state.variables
        out = {}
        for name in ("S_snow", "C_snow", "prec"):  # benchmarks/SVATOXYGEN18_benchmark.py:453-480
            arr = getattr(vs, name)
            arr = update(arr, at[2:-2, 2:-2, vs.taum1], arr[2:-2, 2:-2, vs.tau])
            setattr(vs, name, arr)
            out[name] = arr
        return KernelOutput(**out)

    return O18Setup


def make_bromide_setup(case):
    """Offline bromide (conservative anion) StorAge-selection transport: roger/models/svat_bromide/svat_bromide.py with
    in-memory daily forcing on an (nx, ny) grid.

    case: nx, ny, ndays, ages, daily arrays (nx, ny, ndays + 1): PREC, INF_MAT_RZ, INF_PF_RZ, INF_PF_SS, TRANSP, EVAP_SOIL, CPR_RZ,
    Q_RZ, Q_SS, S_RZ, S_SS, C_IN (bromide concentration of the input); (nx, ny) arrays S_pwp_rz, S_pwp_ss, S_sat_rz, S_sat_ss,
    alpha_transp, alpha_q, lu_id; sas: dict process -> (nx, ny, 8).
    """
    import_reference()
    from roger import RogerSetup, roger_routine
    from roger.variables import allocate
    from roger.core.operators import numpy as npx, update, at

    nx, ny, ndays, ages = case["nx"], case["ny"], case["ndays"], case["ages"]

    class BromideSetup(RogerSetup):
        @roger_routine
        def set_settings(self, state):
            s = state.settings
            s.identifier = "GOLDEN_BR"
            s.sas_solver = case.get("sas_solver", "deterministic")
            s.sas_solver_substeps = case.get("substeps", 6)
            if s.sas_solver in ("RK4", "Euler"):  # roger/models/svat_oxygen18/svat_oxygen18.py:_set_settings
                s.h = 1 / s.sas_solver_substeps
            s.nx, s.ny = nx, ny
            s.runlen = 24 * 60 * 60 * ndays
            s.runlen_warmup = 0
            s.nitt = ndays + 1
            s.ages = ages
            s.nages = ages + 1
            s.dx = 1
            s.dy = 1
            s.x_origin = 0.0
            s.y_origin = 0.0
            s.time_origin = "01-01-2022"
            s.enable_offline_transport = True
            s.enable_bromide = True
            s.tm_structure = "power"
            s.enable_age_statistics = bool(case.get("age_statistics", False))

        @roger_routine
        def read_data(self, state):
            pass

        @roger_routine
        def set_grid(self, state):
            vs = state.variables
            settings = state.settings
            vs.dt_secs = 60 * 60 * 24
            vs.dt = 60 * 60 * 24 / (60 * 60)
            vs.ages = update(vs.ages, at[:], npx.arange(1, settings.nages))
            vs.nages = update(vs.nages, at[:], npx.arange(settings.nages))
            dx = allocate(state.dimensions, ("x"))
            dx = update(dx, at[:], settings.dx)
            dy = allocate(state.dimensions, ("y"))
            dy = update(dy, at[:], settings.dy)
            vs.x = update(vs.x, at[3:-2], settings.x_origin + npx.cumsum(dx[3:-2]))
            vs.y = update(vs.y, at[3:-2], settings.y_origin + npx.cumsum(dy[3:-2]))

        @roger_routine
        def set_look_up_tables(self, state):
            pass

        @roger_routine
        def set_topography(self, state):
            pass

        @roger_routine
        def set_parameters_setup(self, state):
            vs = state.variables
            for k in ("S_pwp_rz", "S_pwp_ss", "S_sat_rz", "S_sat_ss", "alpha_transp", "alpha_q"):
                setattr(vs, k, update(getattr(vs, k), at[2:-2, 2:-2], case[k]))
            vs.lu_id = update(vs.lu_id, at[2:-2, 2:-2], onp.asarray(case["lu_id"]).astype(vs.lu_id.dtype))
            for proc, arr in case["sas"].items():
                name = "sas_params_" + proc
                setattr(vs, name, update(getattr(vs, name), at[2:-2, 2:-2, :], arr))

        @roger_routine
        def set_parameters(self, state):
            pass

        @roger_routine
        def set_initial_conditions_setup(self, state):
            vs = state.variables
            vs.S_rz = update(vs.S_rz, at[2:-2, 2:-2, : vs.taup1], case["S_RZ"][:, :, 0, None])
            vs.S_ss = update(vs.S_ss, at[2:-2, 2:-2, : vs.taup1], case["S_SS"][:, :, 0, None])
            vs.S_s = update(vs.S_s, at[2:-2, 2:-2, : vs.taup1], vs.S_rz[2:-2, 2:-2, : vs.taup1] + vs.S_ss[2:-2, 2:-2, : vs.taup1])
            vs.S_rz_init = update(vs.S_rz_init, at[2:-2, 2:-2], vs.S_rz[2:-2, 2:-2, 0])
            vs.S_ss_init = update(vs.S_ss_init, at[2:-2, 2:-2], vs.S_ss[2:-2, 2:-2, 0])

        @roger_routine
        def set_initial_conditions(self, state):
            # roger/models/svat_bromide/svat_bromide.py:235-266 (no bromide in the soil initially)
            vs = state.variables
            settings = state.settings
            arr0 = allocate(state.dimensions, ("x", "y"))
            vs.sa_rz = update(vs.sa_rz, at[2:-2, 2:-2, : vs.taup1, 1:],
                              npx.diff(npx.linspace(arr0[2:-2, 2:-2], vs.S_rz[2:-2, 2:-2, vs.tau], settings.ages, axis=-1), axis=-1)[:, :, npx.newaxis, :])
            vs.sa_ss = update(vs.sa_ss, at[2:-2, 2:-2, : vs.taup1, 1:],
                              npx.diff(npx.linspace(arr0[2:-2, 2:-2], vs.S_ss[2:-2, 2:-2, vs.tau], settings.ages, axis=-1), axis=-1)[:, :, npx.newaxis, :])
            vs.SA_rz = update(vs.SA_rz, at[2:-2, 2:-2, :, 1:], npx.cumsum(vs.sa_rz[2:-2, 2:-2, :, :], axis=-1))
            vs.SA_ss = update(vs.SA_ss, at[2:-2, 2:-2, :, 1:], npx.cumsum(vs.sa_rz[2:-2, 2:-2, :, :], axis=-1))
            vs.sa_s = update(vs.sa_s, at[2:-2, 2:-2, :, :], vs.sa_rz[2:-2, 2:-2, :, :] + vs.sa_ss[2:-2, 2:-2, :, :])
            vs.SA_s = update(vs.SA_s, at[2:-2, 2:-2, :, 1:], npx.cumsum(vs.sa_s[2:-2, 2:-2, :, :], axis=-1))

        @roger_routine
        def set_boundary_conditions_setup(self, state):
            pass

        @roger_routine
        def set_boundary_conditions(self, state):
            pass

        @roger_routine
        def set_forcing_setup(self, state):
            vs = state.variables
            vs.C_IN = update(vs.C_IN, at[2:-2, 2:-2, :], case["C_IN"])

        @roger_routine
        def set_forcing(self, state):
            # roger/models/svat_bromide/svat_bromide.py:325-348
            vs = state.variables
            t = vs.itt
            vs.prec = update(vs.prec, at[2:-2, 2:-2, vs.tau], case["PREC"][:, :, t])
            for name, key in (("inf_mat_rz", "INF_MAT_RZ"), ("inf_pf_rz", "INF_PF_RZ"), ("inf_pf_ss", "INF_PF_SS"), ("transp", "TRANSP"),
                              ("evap_soil", "EVAP_SOIL"), ("cpr_rz", "CPR_RZ"), ("q_rz", "Q_RZ"), ("q_ss", "Q_SS")):
                setattr(vs, name, update(getattr(vs, name), at[2:-2, 2:-2], case[key][:, :, t]))
            vs.S_rz = update(vs.S_rz, at[2:-2, 2:-2, vs.tau], case["S_RZ"][:, :, t])
            vs.S_ss = update(vs.S_ss, at[2:-2, 2:-2, vs.tau], case["S_SS"][:, :, t])
            vs.S_s = update(vs.S_s, at[2:-2, 2:-2, vs.tau], vs.S_rz[2:-2, 2:-2, vs.tau] + vs.S_ss[2:-2, 2:-2, vs.tau])
            vs.C_in = update(vs.C_in, at[2:-2, 2:-2], vs.C_IN[2:-2, 2:-2, vs.itt])
            vs.M_in = update(vs.M_in, at[2:-2, 2:-2], vs.C_in[2:-2, 2:-2] * vs.prec[2:-2, 2:-2, vs.tau])

        @roger_routine
        def set_diagnostics(self, state):
            pass

        @roger_routine
        def after_timestep(self, state):
            pass

    return BromideSetup


def make_virtualtracer_setup(case):
    """Offline virtual-tracer transport (settings.enable_virtualtracer, examples/plot_scale/freiburg_altheim_kupferzell/svat_transport/
    svat_transport.py:74, :385-392): the conservative-anion model in which the soil evaporation carries the tracer as well
    (evapotranspiration.py:723-791, partition coefficient 1) and whose solute balance counts it (numerics.py:517-549).  `C_IN` is not
    an active variable of this configuration (variables.py:5497-5512): the setup writes the day's `C_in` itself, as the example does."""
    import_reference()
    from roger import roger_routine
    from roger.core.operators import update, at

    Base = make_bromide_setup(case)

    class VirtualTracerSetup(Base):
        @roger_routine
        def set_settings(self, state):
            super().set_settings(state)
            s = state.settings
            s.identifier = "GOLDEN_VT"
            s.enable_bromide = False
            s.enable_virtualtracer = True

        @roger_routine
        def set_forcing_setup(self, state):
            pass

        @roger_routine
        def set_forcing(self, state):
            vs = state.variables
            t = vs.itt
            vs.prec = update(vs.prec, at[2:-2, 2:-2, vs.tau], case["PREC"][:, :, t])
            for name, key in (("inf_mat_rz", "INF_MAT_RZ"), ("inf_pf_rz", "INF_PF_RZ"), ("inf_pf_ss", "INF_PF_SS"), ("transp", "TRANSP"),
                              ("evap_soil", "EVAP_SOIL"), ("cpr_rz", "CPR_RZ"), ("q_rz", "Q_RZ"), ("q_ss", "Q_SS")):
                setattr(vs, name, update(getattr(vs, name), at[2:-2, 2:-2], case[key][:, :, t]))
            vs.S_rz = update(vs.S_rz, at[2:-2, 2:-2, vs.tau], case["S_RZ"][:, :, t])
            vs.S_ss = update(vs.S_ss, at[2:-2, 2:-2, vs.tau], case["S_SS"][:, :, t])
            vs.S_s = update(vs.S_s, at[2:-2, 2:-2, vs.tau], vs.S_rz[2:-2, 2:-2, vs.tau] + vs.S_ss[2:-2, 2:-2, vs.tau])
            vs.C_in = update(vs.C_in, at[2:-2, 2:-2], case["C_IN"][:, :, t])
            vs.M_in = update(vs.M_in, at[2:-2, 2:-2], vs.C_in[2:-2, 2:-2] * vs.prec[2:-2, 2:-2, vs.tau])

    return VirtualTracerSetup


def make_nitrate_setup(case):
    """Offline nitrate StorAge-selection transport with the soil nitrogen cycle: the bromide setup above with
    `settings.enable_nitrate` (transport.svat_transport_model_deterministic, roger/core/transport.py:948-991 -> nitrate.py) and the
    nitrate-specific parameters, forcing and initial conditions of examples/plot_scale/svat_crop_nitrate_tutorial/svat_crop_nitrate.py
    (:255-291 parameters, :337-415 initial mineral nitrogen and nitrate mass, :591-623 forcing incl. a fertiliser pulse).

    case: everything make_bromide_setup takes plus TA (ndays + 1 daily air temperatures), DOY (ndays + 1), Z_ROOT, LU_ID, NUP
    (nx, ny, ndays + 1), z_soil and the nitrate parameter planes (nx, ny): kfix_rz, kmin_rz, kmin_ss, km_nit_rz, dmax_nit_rz, kngl_rz,
    km_denit_rz, dmax_denit_rz, km_denit_ss, dmax_denit_ss, phi_soil_temp, damp_soil_temp; C_rz0, C_ss0, NMIN_RZ0, NMIN_SS0 scalars."""
    import_reference()
    from roger import roger_routine
    from roger.variables import allocate
    from roger.core.operators import numpy as npx, update, at
    import scipy.stats as sstx

    Base = make_bromide_setup(case)
    NIT = ("kfix_rz", "kmin_rz", "kmin_ss", "km_nit_rz", "dmax_nit_rz", "kngl_rz", "km_denit_rz", "dmax_denit_rz", "km_denit_ss",
           "dmax_denit_ss", "phi_soil_temp", "damp_soil_temp", "z_soil")

    class NitrateSetup(Base):
        @roger_routine
        def set_settings(self, state):
            super().set_settings(state)
            s = state.settings
            s.identifier = "GOLDEN_NO3"
            s.enable_bromide = False
            s.enable_nitrate = True
            s.nitt_forc = s.nitt

        @roger_routine
        def set_parameters_setup(self, state):
            super().set_parameters_setup(state)
            vs = state.variables
            for k in NIT:
                setattr(vs, k, update(getattr(vs, k), at[2:-2, 2:-2], case[k]))

        @roger_routine
        def set_initial_conditions(self, state):
            super().set_initial_conditions(state)
            vs = state.variables
            settings = state.settings
            vs.C_rz = update(vs.C_rz, at[2:-2, 2:-2, : vs.taup1], case["C_rz0"])
            vs.C_ss = update(vs.C_ss, at[2:-2, 2:-2, : vs.taup1], case["C_ss0"])
            p_dec = allocate(state.dimensions, ("x", "y", 2, "ages"))
            p_dec = update(p_dec, at[:, :, : vs.taup1, :],
                           sstx.expon.pdf(npx.linspace(sstx.expon.ppf(0.001), sstx.expon.ppf(0.999), settings.ages))[npx.newaxis, npx.newaxis, npx.newaxis, :])
            vs.Nmin_rz = update(vs.Nmin_rz, at[2:-2, 2:-2, : vs.taup1, :], case["NMIN_RZ0"] * p_dec[2:-2, 2:-2, :, :] * settings.dx * settings.dy * 100)
            vs.Nmin_ss = update(vs.Nmin_ss, at[2:-2, 2:-2, : vs.taup1, :], case["NMIN_SS0"] * p_dec[2:-2, 2:-2, :, :] * settings.dx * settings.dy * 100)
            vs.msa_rz = update(vs.msa_rz, at[2:-2, 2:-2, : vs.taup1, :], vs.C_rz[2:-2, 2:-2, : vs.taup1, npx.newaxis] * vs.sa_rz[2:-2, 2:-2, : vs.taup1, :])
            vs.msa_ss = update(vs.msa_ss, at[2:-2, 2:-2, : vs.taup1, :], vs.C_ss[2:-2, 2:-2, : vs.taup1, npx.newaxis] * vs.sa_ss[2:-2, 2:-2, : vs.taup1, :])
            vs.msa_s = update(vs.msa_s, at[2:-2, 2:-2, :, :], vs.msa_rz[2:-2, 2:-2, :, :] + vs.msa_ss[2:-2, 2:-2, :, :])
            vs.M_rz = update(vs.M_rz, at[2:-2, 2:-2, :], npx.sum(vs.msa_rz[2:-2, 2:-2, :, :], axis=-1))
            vs.M_ss = update(vs.M_ss, at[2:-2, 2:-2, :], npx.sum(vs.msa_ss[2:-2, 2:-2, :, :], axis=-1))
            vs.M_s = update(vs.M_s, at[2:-2, 2:-2, :], npx.sum(vs.msa_s[2:-2, 2:-2, :, :], axis=-1))

        @roger_routine
        def set_forcing_setup(self, state):
            super().set_forcing_setup(state)
            vs = state.variables
            vs.TA = update(vs.TA, at[:], case["TA"])
            vs.LU_ID = update(vs.LU_ID, at[2:-2, 2:-2, :], onp.asarray(case["LU_ID"]).astype(vs.LU_ID.dtype))

        @roger_routine
        def set_forcing(self, state):
            super().set_forcing(state)
            vs = state.variables
            t = vs.itt
            vs.ta = update(vs.ta, at[2:-2, 2:-2, vs.tau], vs.TA[t])
            vs.doy = update(vs.doy, at[1], case["DOY"][t])
            vs.z_root = update(vs.z_root, at[2:-2, 2:-2, vs.tau], case["Z_ROOT"][:, :, t])
            vs.nup = update(vs.nup, at[2:-2, 2:-2], case["NUP"][:, :, t])

    return NitrateSetup


def with_crop_redistribution(Base, case):
    """The offline transport setups above with `settings.enable_crop_phenology`: calculate_storage_selection then runs
    svat_crop_transport_model_deterministic (roger/core/transport.py:1042-1087), i.e. the day starts with the redistribution of
    water and solute between subsoil and root zone after root growth / root loss (crop.calculate_redistribution_transport,
    roger/core/crop.py:2461-2867).  The redistribution fluxes are an INPUT of the offline model (the tutorial setup reads them from the
    SVAT-CROP output, examples/plot_scale/svat_crop_nitrate_tutorial/svat_crop_nitrate.py:525-537, :603-604).

    case: plus RE_RG, RE_RL (nx, ny, ndays + 1) and case["sas"]["re_rg"], case["sas"]["re_rl"] (nx, ny, 8)."""
    import_reference()
    from roger import roger_routine
    from roger.core.operators import update, at

    class CropSetup(Base):
        @roger_routine
        def set_settings(self, state):
            super().set_settings(state)
            state.settings.enable_crop_phenology = True
            state.settings.identifier = state.settings.identifier + "_CROP"

        @roger_routine
        def set_parameters_setup(self, state):
            super().set_parameters_setup(state)
            vs = state.variables
            vs.sas_params_re_rg = update(vs.sas_params_re_rg, at[2:-2, 2:-2, :], case["sas"]["re_rg"])
            vs.sas_params_re_rl = update(vs.sas_params_re_rl, at[2:-2, 2:-2, :], case["sas"]["re_rl"])

        @roger_routine
        def set_forcing(self, state):
            super().set_forcing(state)
            vs = state.variables
            t = vs.itt
            vs.re_rg = update(vs.re_rg, at[2:-2, 2:-2], case["RE_RG"][:, :, t])
            vs.re_rl = update(vs.re_rl, at[2:-2, 2:-2], case["RE_RL"][:, :, t])

    return CropSetup
