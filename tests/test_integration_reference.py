"""The drop-in boundary: an UNMODIFIED RogerSetup subclass of the reference is
stepped by roger_b200.integration.CudaStepDriver and must reproduce what the reference's own `RogerSetup.step` computes
-- same hooks, same `vs` arrays with their tau / taum1 levels, same scalars.

Runs only where the reference checkout is importable (the build container); the engine plugged in here is the CPU
oracle (tests may do that), the product default is the CUDA engine, which has no fallback."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/roger"), reason="reference checkout not present on this box")


def _case(name):
    import make_svat_golden as mg

    case, _ = mg.build_case(name)
    return case


@pytest.mark.parametrize("name,nsteps", [("svat_mixed_dist", 260), ("oned_rain", 50), ("svat_snowrain", 50), ("svat_gwb", 120), ("oned_gwb", 120)])
def test_driver_reproduces_reference_step(name, nsteps):
    import ref_harness as rh
    from oracle.engine import OracleSvatEngine
    from roger_b200 import abi
    from roger_b200.integration import CudaStepDriver
    from tests.util import normalised_max_error

    case = _case(name)
    Setup = rh.make_svat_setup(case)
    ref = Setup()
    ref.setup()
    dut = Setup()
    dut.setup()
    driver = CudaStepDriver(dut, engine_cls=OracleSvatEngine)
    names = [n for n in abi.FIELD_NAMES if n in driver._active and n != "maskRiverLake"]
    worst = {}
    for step in range(nsteps):
        ref.step(ref.state)
        driver.step()
        a, b = ref.state.variables, dut.state.variables
        for k in ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt"):
            assert float(np.asarray(getattr(a, k))) == float(np.asarray(getattr(b, k))), (step, k)
        assert list(a.event_id) == list(b.event_id), step
        if step % 7 == 0 or step == nsteps - 1:
            for n in names:
                x, y = np.asarray(getattr(a, n)), np.asarray(getattr(b, n))
                assert x.shape == y.shape and x.dtype == y.dtype
                if n.endswith("_num_error"):
                    continue
                e = normalised_max_error(x[2:-2, 2:-2], y[2:-2, 2:-2])  # both time levels where the variable has them
                if e > worst.get(n, (0, 0))[0]:
                    worst[n] = (e, step)
    bad = {n: v for n, v in worst.items() if v[0] > 1e-10}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0])[:10]


def test_driver_refuses_offline_transport_and_has_no_fallback():
    import ref_harness as rh
    from roger_b200.engine import RogerCudaError
    from roger_b200.integration import CudaStepDriver

    Setup = rh.make_svat_setup(_case("svat_rain"))
    m = Setup()
    m.setup()
    import torch

    if not torch.cuda.is_available():
        d = CudaStepDriver(m)  # product default engine
        with pytest.raises(RogerCudaError, match="no CUDA device"):
            d.step()


@pytest.mark.parametrize("name,ndays,tol", [("sas_families", 12, 1e-10), ("sas_power", 5, 1e-10), ("sas_bromide", 22, 1e-3),
                                            ("sas_euler_families", 10, 1e-10), ("sas_euler_bromide", 10, 1e-3)])
def test_transport_driver_reproduces_reference_step(name, ndays, tol):
    """The offline oxygen-18 branch: an unmodified RogerSetup subclass of the reference stepped by `RogerSetup.step` and by
    `CudaTransportDriver.step` (oracle engine plugged in).  The reference runs on NumPy's default SVML `power` here, so
    the power-law case is compared over the days before its first rounding-decided clamp; the bromide
    case (exponent 0.1, clamp on day 14, solute pulse on day 17) is compared inside the reference's own libm envelope -- its
    bit-for-bit pin is tests/test_sas_oracle_golden.py, where both sides call the same pow."""
    import make_sas_golden as mg
    import ref_harness as rh
    from oracle.engine import OracleSasEngine, lib
    from roger_b200 import sas
    from roger_b200.integration import CudaTransportDriver
    from tests.util_sas import rel_err

    lib().roger_oracle_set_pow(0)
    case = mg.build(name)
    bromide = case.get("solute") == "bromide"
    Setup = rh.make_bromide_setup(case) if bromide else rh.make_o18_setup(case)
    ref = Setup()
    ref.setup()
    ref.warmup()
    dut = Setup()
    dut.setup()
    dut.warmup()
    driver = CudaTransportDriver(dut, engine_cls=OracleSasEngine)
    a, b = ref.state.variables, dut.state.variables
    meta = ref.state.var_meta
    names = list(sas.STATE) + ["tt_" + p for p in sas.PROCESSES] + [n for n in sas.OUTPUTS if n in meta and meta[n].active]
    if bromide:  # (the travel-time distributions of a power law with exponent 0.1 react to the libm by up to 3e-3: left out)
        names = [n for n in names if n != "C_evap_soil" and not n.startswith("tt_")] + ["M_transp", "M_q_rz", "M_q_ss", "M_cpr_rz", "M_rz", "M_ss", "M_s"]
    for day in range(ndays):
        ref.step(ref.state)
        driver.step()
        assert int(a.itt) == int(b.itt) and int(a.time) == int(b.time)
        for n in names:
            x, y = np.asarray(getattr(a, n))[2:-2, 2:-2], np.asarray(getattr(b, n))[2:-2, 2:-2]
            assert x.shape == y.shape
            lim = max(1e-9, tol) if n.startswith("C_iso") else tol  # delta values amplify concentration rounding ~1e3
            assert rel_err(y, x) < lim, (day, n)


def test_transport_driver_reproduces_reference_nitrate_step():
    """settings.enable_nitrate: the reference's nitrate setup (anion transport + soil nitrogen cycle) stepped by `RogerSetup.step`
    and by `CudaTransportDriver.step` (oracle engine): the six age-ranked arrays, the nitrogen-cycle fluxes, the soil temperature
    the driver computes on the host, concentrations and masses."""
    import make_sas_golden as mg
    import ref_harness as rh
    from oracle.engine import OracleSasEngine, lib
    from roger_b200.integration import CudaTransportDriver
    from tests.util_sas import rel_err

    lib().roger_oracle_set_pow(0)
    case = mg.build("sas_nitrate")
    Setup = rh.make_nitrate_setup(case)
    ref, dut = Setup(), Setup()
    for m in (ref, dut):
        m.setup()
        m.warmup()
    driver = CudaTransportDriver(dut, engine_cls=OracleSasEngine)
    a, b = ref.state.variables, dut.state.variables
    names = ["sa_rz", "sa_ss", "msa_rz", "msa_ss", "Nmin_rz", "Nmin_ss", "Nmin_s", "temp_soil", "nit_s", "denit_s", "min_s", "ngas_s", "nfix_s",
             "nh4_up", "C_rz", "C_ss", "C_s", "M_rz", "M_ss", "M_s", "C_q_ss", "M_q_ss", "M_transp"]
    for day in range(16):
        ref.step(ref.state)
        driver.step()
        assert int(a.itt) == int(b.itt) and int(a.time) == int(b.time)
        for n in names:
            x, y = np.asarray(getattr(a, n))[2:-2, 2:-2], np.asarray(getattr(b, n))[2:-2, 2:-2]
            assert x.shape == y.shape
            # the reference runs on NumPy's SVML `power` in this process, the oracle on the C library's: 1e-10 until the first
            # rounding-decided clamp of the power-law solver (day 11 here), the reference's own libm envelope afterwards; the
            # bit-for-bit pin of this model is tests/test_nitrate_oracle_golden.py, where both sides call the same pow
            assert rel_err(y, x) < (1e-10 if day < 10 else 1e-3), (day, n, rel_err(y, x))


@pytest.mark.parametrize("name,ndays,tight", [("sas_rk4_stats", 8, 8), ("sas_power_lbc", 5, 5), ("sas_nitrate_crop", 10, 10), ("sas_bromide_crop", 11, 4),
                                              ("sas_nitrate_crop_stats", 10, 4), ("sas_bromide_crop_stats", 11, 4),
                                              ("sas_virtualtracer", 20, 4), ("sas_virtualtracer_crop", 11, 4), ("sas_euler_virtualtracer_stats", 8, 4)])
def test_transport_driver_variants_reproduce_reference_step(name, ndays, tight):
    """RK4 solver (with age statistics); groundwater boundary (svat_lbc_transport_model_deterministic: same transport, balance errors
    left at 0); crop variant (redistribution after root growth / root loss) for nitrate and bromide.  The reference runs on NumPy's
    SVML `power` in this process (see above): 1e-10 over the `tight` days before a rounding-decided clamp, the reference's own libm
    envelope (1e-3 for the bromide exponents 0.1) afterwards -- up to the harvest pulse that empties a root zone, after which that
    cell's concentrations are quotients of rounding residues; the bit-for-bit pins are the oracle tests on the fixtures."""
    import make_sas_golden as mg
    import ref_harness as rh
    from oracle.engine import OracleSasEngine, lib
    from roger_b200 import sas
    from roger_b200.integration import CudaTransportDriver
    from tests.util_sas import rel_err

    lib().roger_oracle_set_pow(0)
    case = mg.build(name)
    solute = case.get("solute")
    Setup = {"nitrate": rh.make_nitrate_setup, "bromide": rh.make_bromide_setup, "virtualtracer": rh.make_virtualtracer_setup}.get(solute, rh.make_o18_setup)(case)
    if case.get("crop"):
        Setup = rh.with_crop_redistribution(Setup, case)
    ref, dut = Setup(), Setup()
    for m in (ref, dut):
        m.setup()
        m.warmup()
    driver = CudaTransportDriver(dut, engine_cls=OracleSasEngine)
    a, b = ref.state.variables, dut.state.variables
    meta = ref.state.var_meta
    names = list(sas.STATE) + [n for n in sas.OUTPUTS if n in meta and meta[n].active and n != "C_evap_soil"]
    if solute:
        names += ["M_transp", "M_q_rz", "M_q_ss", "M_cpr_rz", "M_rz", "M_ss", "M_s"]
    if solute == "nitrate":
        names += ["Nmin_rz", "Nmin_ss", "Nmin_s", "nit_s", "denit_s", "min_s"]
    if solute == "virtualtracer":  # settings.enable_virtualtracer: the soil evaporation carries the tracer
        names += ["C_evap_soil", "M_evap_soil"]
    if case.get("crop"):
        names += list(sas.REDISTRIBUTION)
    if case["age_statistics"]:
        names += [n for n in sas.STATISTICS if n != "sanity_ok"]
    moved = 0.0
    for day in range(ndays):
        ref.step(ref.state)
        driver.step()
        assert int(a.itt) == int(b.itt) and int(a.time) == int(b.time)
        for n in names:
            x, y = np.asarray(getattr(a, n))[2:-2, 2:-2], np.asarray(getattr(b, n))[2:-2, 2:-2]
            assert x.shape == y.shape
            if "num_error" in n:  # rounding residues (1e-14 mm): compared absolutely
                assert np.nanmax(np.abs(x - y)) < 1e-9, (day, n)
                continue
            lim = (1e-9 if n.startswith("C_iso") else 1e-10) if day < tight else 1e-3
            assert rel_err(y, x) < lim, (day, n, rel_err(y, x))
        if case.get("crop"):
            moved += float(np.asarray(b.re_rg).sum() + np.asarray(b.re_rl).sum())  # water moved between the layers
    if case.get("crop"):
        assert moved > 0
    if name == "sas_virtualtracer":
        assert float(np.nanmax(np.asarray(b.M_evap_soil))) > 0  # the pulse has arrived and evaporates
    if name == "sas_power_lbc":  # the reference leaves the balance errors at 0 under the groundwater boundary, and so does the driver
        assert float(np.abs(np.asarray(b.dS_num_error)).max()) == 0.0 and float(np.abs(np.asarray(a.dS_num_error)).max()) == 0.0


def test_transport_driver_refuses_what_the_reference_has_no_working_branch_for():
    import make_sas_golden as mg
    import ref_harness as rh
    from oracle.engine import OracleSasEngine
    from roger_b200.integration import CudaTransportDriver

    case = mg.build("sas_bromide")
    case["sas_solver"] = "RK4"
    m = rh.make_bromide_setup(case)()
    m.setup()
    with pytest.raises(NotImplementedError, match="isotope branch only"):
        CudaTransportDriver(m, engine_cls=OracleSasEngine)
    # the crop variant is driven for anion / nitrate / virtual-tracer solutes with the deterministic solver: the isotope crop model
    # (its redistribution kernel doubles the bin concentrations per day, crop.py:2405-2425) and the Euler crop model (reads variables the
    # reference does not declare, transport.py:3012) are refused loudly
    case = mg.build("sas_power")
    case["RE_RG"] = np.zeros((case["nx"], case["ny"], case["ndays"] + 1))
    case["RE_RL"] = np.zeros((case["nx"], case["ny"], case["ndays"] + 1))
    case["sas"] = dict(case["sas"], re_rg=np.zeros((case["nx"], case["ny"], 8)), re_rl=np.zeros((case["nx"], case["ny"], 8)))
    m = rh.with_crop_redistribution(rh.make_o18_setup(case), case)()
    m.setup()
    with pytest.raises(NotImplementedError, match="enable_crop_phenology"):
        CudaTransportDriver(m, engine_cls=OracleSasEngine)
    case = mg.build("sas_bromide_crop")
    case["sas_solver"] = "Euler"
    m = rh.with_crop_redistribution(rh.make_bromide_setup(case), case)()
    m.setup()
    with pytest.raises(NotImplementedError, match="enable_crop_phenology"):
        CudaTransportDriver(m, engine_cls=OracleSasEngine)


def test_driver_writes_the_files_the_setups_diagnostics_ask_for(tmp_path):
    """`CudaStepDriver.enable_output`: the setup's own `state.diagnostics` configuration (rate / average / collect with their
    output variables, frequency and path) produces files whose records equal what the reference's diagnostics hold in memory at
    the same output times (the reference itself cannot write here: h5netcdf is absent, it runs diskless)."""
    import ref_harness as rh
    from oracle.engine import OracleSvatEngine
    from roger_b200.integration import CudaStepDriver
    from scipy.io import netcdf_file
    from tests.util import normalised_max_error

    case = _case("svat_rain")
    Setup = rh.make_svat_setup(case)
    # maximum / minimum: theta_ss and S carry a time-step dimension in the reference (their record is the extreme of the last two
    # time levels, maximum.py:60-75), inf_mat and aet_soil are plain (x, y) variables (running extremes)
    want = {"rate": ["q_ss", "aet", "prec"], "average": ["theta_rz"], "collect": ["S_s"], "maximum": ["theta_ss", "inf_mat"],
            "minimum": ["S", "aet_soil"]}

    def configure(model):
        for name, vars_ in want.items():
            d = model.state.diagnostics[name]
            d.output_variables = list(vars_)
            d.output_frequency = 24 * 60 * 60
            d.sampling_frequency = 1
            d.initialize(model.state)

    ref = Setup()
    ref.setup()
    configure(ref)
    dut = Setup()
    dut.setup()
    configure(dut)
    driver = CudaStepDriver(dut, engine_cls=OracleSvatEngine)
    driver.enable_output(base_path=str(tmp_path), force_overwrite=True)
    # Maximum / Minimum zero their arrays right after writing (maximum.py:94-113): catch what they hand to write_output
    written = {}
    for name in ("maximum", "minimum"):
        d = ref.state.diagnostics[name]

        def spy(state, _d=d, _orig=d.write_output):
            for v in _d.output_variables:
                written[v] = np.asarray(getattr(_d.variables, v))[2:-2, 2:-2].copy()
            return _orig(state)

        d.write_output = spy
    records = []
    for step in range(400):
        ref.step(ref.state)
        driver.step()
        t = int(ref.state.variables.time)
        if t % 86400 == 0:
            rec = {}
            for name, vars_ in want.items():
                dv = ref.state.diagnostics[name].variables
                for v in vars_:
                    if name in ("maximum", "minimum"):
                        rec[v] = written[v]
                        continue
                    a = np.asarray(getattr(dv, v))
                    rec[v] = a[2:-2, 2:-2].copy() if a.ndim == 2 else a[2:-2, 2:-2, int(ref.state.variables.tau)].copy()
            records.append((t, rec))
            if len(records) == 2:
                break
    driver.close_output()
    assert len(records) == 2
    ident = dut.state.settings.identifier
    mask = np.asarray(dut.state.variables.maskCatch)[2:-2, 2:-2].astype(bool)
    for name, vars_ in want.items():
        with netcdf_file(str(tmp_path / f"{ident}.{name}.nc"), "r", mmap=False) as f:
            times = f.variables["Time"][:]
            assert np.array_equal(times, [t / 86400.0 for t, _ in records])
            for k, v in [(k, v) for k in range(len(records)) for v in vars_]:
                got = f.variables[v][k].T
                ref_v = records[k][1][v]
                # (Average.output has already divided the in-memory sums by the step count, average.py:93-97)
                assert v in ("prec", "inf_mat", "aet_soil") or k == 0 or np.abs(ref_v[mask]).sum() > 0, (name, v)  # (the second day of this case is dry)
                assert normalised_max_error(np.where(mask, ref_v, 0), np.where(mask, got, 0)) < 1e-10, (name, v)


def test_transport_driver_writes_the_files_the_setups_diagnostics_ask_for(tmp_path):
    """`CudaTransportDriver.enable_output`: daily records of the per-cell transport results equal what the reference's collect /
    average diagnostics hold in memory after the same step."""
    import make_sas_golden as mg
    import ref_harness as rh
    from oracle.engine import OracleSasEngine, lib
    from roger_b200.integration import CudaTransportDriver
    from scipy.io import netcdf_file
    from tests.util_sas import rel_err

    lib().roger_oracle_set_pow(0)
    case = mg.build("sas_families")
    Setup = rh.make_o18_setup(case)
    want = {"collect": ["C_iso_q_ss", "C_iso_transp", "C_iso_s"], "average": ["C_rz"]}

    def configure(model):
        for name, vars_ in want.items():
            d = model.state.diagnostics[name]
            d.output_variables = list(vars_)
            d.output_frequency = 24 * 60 * 60
            d.sampling_frequency = 1
            d.initialize(model.state)

    ref, dut = Setup(), Setup()
    for m in (ref, dut):
        m.setup()
        m.warmup()
        configure(m)
    driver = CudaTransportDriver(dut, engine_cls=OracleSasEngine)
    driver.enable_output(base_path=str(tmp_path), force_overwrite=True)
    records = []
    for day in range(6):
        ref.step(ref.state)
        driver.step()
        rec = {}
        for name, vars_ in want.items():
            dv = ref.state.diagnostics[name].variables
            for v in vars_:
                rec[v] = np.asarray(getattr(dv, v))[2:-2, 2:-2].copy()
        records.append((int(ref.state.variables.time), rec))
    driver.close_output()
    ident = dut.state.settings.identifier
    for name, vars_ in want.items():
        with netcdf_file(str(tmp_path / f"{ident}.{name}.nc"), "r", mmap=False) as f:
            assert np.array_equal(f.variables["Time"][:], [t / 86400.0 for t, _ in records])
            for v in vars_:
                for k, (_, rec) in enumerate(records):
                    got = f.variables[v][k].T
                    want_v = np.where(np.isnan(rec[v]), np.nan, rec[v])
                    assert rel_err(np.where(got == -9999, np.nan, got) if np.isnan(want_v).any() else got, want_v) < 1e-9, (name, v, k)


@pytest.mark.parametrize("which", ["deuterium", "chloride"])
def test_transport_driver_runs_the_sibling_solutes(which):
    """settings.enable_deuterium runs the oxygen-18 kernels with its own standard ratio and plausibility window (transport.py:315-340:
    VSMOW_conc2H, d2H_min / d2H_max), settings.enable_chloride the bromide kernels (the reference groups the two in every dispatch; only
    the setup-time initial solute mass differs, soil.py:1508-1525, which is the reference's own code on both sides here)."""
    import make_sas_golden as mg
    import ref_harness as rh
    from oracle.engine import OracleSasEngine, lib
    from roger_b200 import sas
    from roger_b200.integration import CudaTransportDriver
    from tests.util_sas import rel_err

    lib().roger_oracle_set_pow(0)
    rh.import_reference()
    from roger import roger_routine

    if which == "deuterium":
        case = mg.build("sas_power")
        case["C_ISO_IN"] = 8.0 * case["C_ISO_IN"] + 10.0  # the meteoric water line
        Base, on, off = rh.make_o18_setup(case), "enable_deuterium", "enable_oxygen18"
    else:
        case = mg.build("sas_bromide")
        Base, on, off = rh.make_bromide_setup(case), "enable_chloride", "enable_bromide"

    class Setup(Base):
        @roger_routine
        def set_settings(self, state):
            super().set_settings(state)
            setattr(state.settings, off, False)
            setattr(state.settings, on, True)

    ref, dut = Setup(), Setup()
    for m in (ref, dut):
        m.setup()
        m.warmup()
    driver = CudaTransportDriver(dut, engine_cls=OracleSasEngine)
    if which == "deuterium":
        st = dut.state.settings
        assert driver.settings.VSMOW_conc18O == float(st.VSMOW_conc2H) and driver.settings.d18O_min == float(st.d2H_min)
    a, b = ref.state.variables, dut.state.variables
    meta = ref.state.var_meta
    names = list(sas.STATE) + [n for n in sas.OUTPUTS if n in meta and meta[n].active and n != "C_evap_soil"]
    if which == "chloride":
        names = [n for n in names if not n.startswith("C_iso")] + ["M_transp", "M_q_rz", "M_q_ss", "M_cpr_rz", "M_rz", "M_ss", "M_s"]
    for day in range(5):  # the days before the first rounding-decided clamp of the power-law solver (SVML `power` on the reference side)
        ref.step(ref.state)
        driver.step()
        for n in names:
            x, y = np.asarray(getattr(a, n))[2:-2, 2:-2], np.asarray(getattr(b, n))[2:-2, 2:-2]
            lim = 1e-8 if n.startswith("C_iso") else 1e-10  # delta values amplify concentration rounding
            assert rel_err(y, x) < lim, (day, n, rel_err(y, x))
    if which == "deuterium":  # concentrations on deuterium's scale (VSMOW 2H/1H = 1.5576e-4), not oxygen-18's (2.0052e-3)
        assert float(np.nanmax(np.asarray(b.C_rz)[2:-2, 2:-2])) < 1e-3 < float(sas.delta_to_conc(-10.0))
