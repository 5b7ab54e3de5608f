"""GPU parity: the fused sm_100a SVAT / oneD step (through the C ABI) against (a) the committed reference
fixtures and (b) the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

# float64 state within 1e-10 relative (north star); the normalised-max idiom of test/restart_test.py:7-17
TOL = 1e-10


def _engine():
    from roger_b200.engine import CudaSvatEngine

    return CudaSvatEngine


@pytest.mark.parametrize("case", util.SVAT_CASES + util.ONED_CASES)
def test_cuda_matches_reference_fixture(case):
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, _engine(), from_reference_setup=True)
    worst = util.replay_and_compare(fx, model, TOL, write_fluxes=True)
    print(case, sorted(worst.items(), key=lambda kv: -kv[1][0])[:4])


@pytest.mark.parametrize("case", util.SVAT_CASES + util.ONED_CASES)
def test_cuda_reproduces_reference_fixture_bit_for_bit(case):
    """The kernel evaluates the C library's pow (roger_pow.h returns its bits, tests/test_pow.py) and otherwise the reference's own
    operation order: every recorded plane of every fixture -- and, for the two 256-cell fixtures, the digest of every field after
    EVERY step -- equals the numbers the unmodified reference's NumPy backend produced, bit for bit."""
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, _engine(), from_reference_setup=True)
    util.replay_and_compare(fx, model, 0.0, write_fluxes=True, exact=True)


@pytest.mark.parametrize("lateral,doy,dist,nsteps", [(True, 121, False, 70), (False, 40, False, 90), (False, 150, True, 60), (True, 330, True, 50)])
def test_cuda_equals_oracle_bit_for_bit(lateral, doy, dist, nsteps):
    """Kernel and oracle evaluate the same pow (roger_pow.h): every parameter, state and flux plane of 1024 heterogeneous
    cells -- rain events, snow, per-cell forcing, lateral flow -- must be identical in every bit after every step."""
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import abi, synth
    from roger_b200.svat import SvatModel

    shape = (32, 32)
    forcing = synth.mixed_forcing(45, seed=42, start_doy=doy)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lateral=lateral, lu_choices=(10, 11, 12, 8, 5, 9, 6), masked_fraction=0.05)
    weights = synth.distributed_weights(shape, seed=44) if dist else None
    lib().roger_oracle_set_pow(1)
    try:
        models = [SvatModel(e, shape, forcing, prim, th_rz, th_ss, lateral=lateral, weights=weights) for e in (_engine(), OracleSvatEngine)]
        for m in models:
            m.setup()
        models[0].engine.set_block_stepping(False)
        for step in range(nsteps):
            a, b = models[0].step(True), models[1].step(True)
            for k in ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt", "sanity_ok"):
                assert getattr(a, k) == getattr(b, k), (step, k)
            if step % 3 == 0 or step == nsteps - 1:
                for n in abi.FIELD_NAMES:
                    assert util.same_bits(models[0].get(n), models[1].get(n)), (step, n)
    finally:
        lib().roger_oracle_set_pow(0)


@pytest.mark.parametrize("lateral,doy,nsteps", [(False, 100, 80), (True, 300, 60)])
def test_groundwater_boundary_cuda_equals_oracle_bit_for_bit(lateral, doy, nsteps):
    """settings.enable_groundwater_boundary on 1024 heterogeneous cells: a water table that moves every day -- in a third of the
    cells up into the subsoil and back --, capillary rise into the subsoil, recharge through the vadose-zone reservoir; every plane
    every third step, state-only steps in between (the production mode)"""
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import abi, synth
    from roger_b200.svat import SvatModel

    shape = (32, 32)
    ndays = 30
    forcing = synth.mixed_forcing(ndays, seed=52, start_doy=doy)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=53, lateral=lateral, masked_fraction=0.05)
    rng = np.random.default_rng(54)
    prim["n0"] = rng.uniform(0.05, 0.3, shape)
    base = np.asarray(prim["z_soil"]) / 1000 + np.where(rng.random(shape) < 0.35, rng.uniform(0.02, 0.4, shape), rng.uniform(0.4, 12.0, shape))
    z_gw = base[:, :, None] - rng.uniform(0.1, 0.6, shape)[:, :, None] * np.sin(np.arange(ndays + 1) / 3.0)[None, None, :]
    z_gw[:, :, 0] = base
    lib().roger_oracle_set_pow(1)
    try:
        models = [SvatModel(e, shape, forcing, prim, th_rz, th_ss, lateral=lateral, z_gw_daily=z_gw.reshape(-1, ndays + 1))
                  for e in (_engine(), OracleSvatEngine)]
        for m in models:
            m.setup()
        moved = 0.0
        for step in range(nsteps):
            full = step % 3 == 0 or step == nsteps - 1
            a, b = models[0].step(full), models[1].step(True)
            for k in ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt", "sanity_ok"):
                assert getattr(a, k) == getattr(b, k), (step, k)
            kinds = dict(zip(abi.FIELD_NAMES, abi.FIELD_KINDS))
            for n in abi.FIELD_NAMES:
                if full or (kinds[n] != "F" and n not in ("k", "h")):
                    assert util.same_bits(models[0].get(n), models[1].get(n)), (step, n)
            if full:
                moved += float(models[0].get("cpr_ss").sum() + models[0].get("q_re").sum())
        assert moved > 0
    finally:
        lib().roger_oracle_set_pow(0)


@pytest.mark.parametrize("case", ["svat_rain", "svat_mixed_dist", "oned_mixed"])
def test_cuda_state_only_mode_matches(case):
    """write_fluxes=0 (the production mode) must give the same state as the diagnostics mode."""
    fx = util.load_case(case)
    model = util.model_from_fixture(fx, _engine(), from_reference_setup=True)
    util.replay_and_compare(fx, model, TOL, kinds=("S",), write_fluxes=False)


def test_async_download_is_a_snapshot():
    """roger_svat_download_async returns the plane as of the call, even when more steps are issued before the wait."""
    import torch

    fx = util.load_case("svat_mixed_dist")
    model = util.model_from_fixture(fx, _engine(), from_reference_setup=True)
    eng = model.engine
    for _ in range(5):
        model.step(False)
    want = {n: eng.download(n) for n in ("theta_rz", "S")}
    pins = {n: torch.empty(eng.n, dtype=torch.float64).pin_memory().numpy() for n in want}
    for n in want:
        eng.download_async(n, pins[n])
    for _ in range(30):
        model.step(False)
    eng.download_wait()
    for n in want:
        assert np.array_equal(pins[n], want[n], equal_nan=True)
    assert eng.get_scalars().itt == 35


@pytest.mark.parametrize("lateral,doy,ndays", [(True, 121, 25), (False, 121, 25), (False, 40, 30), (True, 330, 20)])
def test_day_blocked_stepping_equals_single_steps(lateral, doy, ndays):
    """roger_svat_run_steps lays out snow-free stretches in advance and runs them in one launch; every plane and every
    scalar must be bit-identical to stepping one launch at a time (doy 40 / 330: snow periods, mixed paths)."""
    from roger_b200 import abi, synth
    from roger_b200.svat import SvatModel

    shape = (24, 16)
    forcing = synth.mixed_forcing(ndays + 1, seed=42, start_doy=doy)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lateral=lateral, lu_choices=(10, 11, 12, 8, 5, 9))
    models = [SvatModel(_engine(), shape, forcing, prim, th_rz, th_ss, lateral=lateral) for _ in range(3)]
    for m in models:
        m.setup()
    models[1].engine.set_block_stepping(0)  # one launch and one flag read-back per step
    models[2].engine.set_block_stepping(2)  # queued single-step launches, one read-back per batch (the default of large grids)
    kinds = {n: k for n, k in zip(abi.FIELD_NAMES, abi.FIELD_KINDS)}
    total = 0
    for day in range(ndays):
        n0 = models[0].run_steps(144)
        n1 = models[1].run_steps(144)
        n2 = models[2].run_steps(144)
        assert n0 == n1 == n2
        total += n0
        b = models[1].engine.get_scalars()
        for m in (models[0], models[2]):
            a = m.engine.get_scalars()
            for k in ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt", "sanity_ok"):
                assert getattr(a, k) == getattr(b, k), (day, k)
            assert list(a.event_id) == list(b.event_id)
            for n in abi.FIELD_NAMES:
                if kinds[n] == "S" or n == "ks_ss":
                    x, y = m.get(n), models[1].get(n)
                    assert np.array_equal(x, y, equal_nan=True), (day, n)
    launches = [m.engine.launches for m in models]
    assert launches[1] >= total and launches[2] >= total
    if doy == 121:
        assert launches[0] < launches[1] / 3  # the blocked path really ran


@pytest.mark.parametrize("lateral,doy,dist", [(False, 45, True), (True, 121, False)])
def test_drop_in_driver_on_the_cuda_engine(lateral, doy, dist):
    """roger_b200.integration.CudaStepDriver driving the CUDA engine through a setup object with the
    reference's conventions (ghost cells, tau / taum1 levels, hooks; tests/mini_setup.py) against the plain host-side
    driver on the oracle: every variable the driver writes back, both time levels, and the scalars -- across a month
    change (land-use parameter refresh) and, for doy 45, snow."""
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import abi, synth
    from roger_b200.integration import CudaStepDriver
    from roger_b200.svat import SvatModel
    from tests.mini_setup import TWO_LEVEL, MiniSvatSetup

    shape = (8, 6)
    nsteps = 220
    forcing = synth.mixed_forcing(60, seed=42, start_doy=doy)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lateral=lateral, lu_choices=(10, 11, 12, 8, 5, 9), masked_fraction=0.1)
    weights = synth.distributed_weights(shape, seed=44) if dist else None
    lib().roger_oracle_set_pow(1)
    try:
        setup = MiniSvatSetup(shape, forcing, prim, th_rz, th_ss, lateral=lateral, weights=weights)
        setup.setup()
        driver = CudaStepDriver(setup, engine_cls=_engine())
        ref = SvatModel(OracleSvatEngine, shape, forcing, prim, th_rz, th_ss, lateral=lateral, weights=weights)
        ref.setup()
        vs = setup.state.variables
        kinds = dict(zip(abi.FIELD_NAMES, abi.FIELD_KINDS))
        months = set()
        for step in range(nsteps):
            prev = {n: np.array(getattr(vs, n))[2:-2, 2:-2, 1].copy() for n in ("S", "theta_rz", "swe")}
            sc = driver.step()
            rs = ref.step(True)
            months.add(int(vs.month[1]))
            for k in ("time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt"):
                assert float(np.asarray(getattr(vs, k))) == float(getattr(rs, k)), (step, k)
            assert [int(v) for v in vs.event_id] == [int(rs.event_id[0]), int(rs.event_id[1])]
            if step % 10 == 0 or step == nsteps - 1:
                for n in abi.FIELD_NAMES:
                    if kinds[n] == "P" and n != "ks_ss" or n == "maskRiverLake":
                        continue
                    a = np.asarray(getattr(vs, n))
                    want = ref.get(n).reshape(shape)
                    got = a[2:-2, 2:-2, 1] if a.ndim == 3 else a[2:-2, 2:-2]
                    assert util.same_bits(got, want), (step, n)
                    if a.ndim == 3 and n in TWO_LEVEL:
                        assert util.same_bits(a[2:-2, 2:-2, 0], want), (step, n, "taum1")
                    assert np.all(a[:2] == 0) and np.all(a[-2:] == 0)  # ghost cells untouched
            for n, before in prev.items():
                pass
        assert len(months) >= 2  # the monthly parameter refresh was exercised
    finally:
        lib().roger_oracle_set_pow(0)


def test_rate_diagnostics_accumulated_on_the_device():
    """RATE accumulators (roger/diagnostics/rate.py): daily sums of flux planes built on the device equal the sums the oracle
    builds step by step -- the daily hydrologic forcing of the offline transport model."""
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import sas_inputs, synth
    from roger_b200.svat import SvatModel

    shape = (16, 12)
    forcing = synth.mixed_forcing(8, seed=42, start_doy=121)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43)
    lib().roger_oracle_set_pow(1)
    try:
        daily = []
        for e in (_engine(), OracleSvatEngine):
            m = SvatModel(e, shape, forcing, prim, th_rz, th_ss)
            m.setup()
            daily.append(sas_inputs.daily_fluxes_from_svat(m, 5))
        for k in daily[0]:
            assert util.same_bits(daily[0][k], daily[1][k]), k
        assert daily[0]["TRANSP"][:, 1:].sum() > 0 and daily[0]["PREC"][:, 1:].sum() > 0
    finally:
        lib().roger_oracle_set_pow(0)


@pytest.mark.parametrize("shape,masked", [((1, 1), 0.0), ((257, 1), 0.0), ((3, 5), 1.0), ((31, 9), 0.3)])
def test_edge_grids_bit_for_bit(shape, masked):
    """a single cell, one cell more than a thread block, a grid entirely outside the catchment (maskCatch = 0 everywhere),
    a heavily masked grid; per-cell forcing with a snow period"""
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import abi, synth
    from roger_b200.svat import SvatModel

    forcing = synth.mixed_forcing(30, seed=42, start_doy=50)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lu_choices=(10, 11, 12, 8, 5), masked_fraction=masked)
    weights = synth.distributed_weights(shape, seed=44)
    lib().roger_oracle_set_pow(1)
    try:
        models = [SvatModel(e, shape, forcing, prim, th_rz, th_ss, weights=weights) for e in (_engine(), OracleSvatEngine)]
        for m in models:
            m.setup()
        for step in range(60):
            a, b = models[0].step(True), models[1].step(True)
            assert (a.time, a.dt_secs, a.itt_day, a.event_id[1], a.time_event0) == (b.time, b.dt_secs, b.itt_day, b.event_id[1], b.time_event0), step
        for n in abi.FIELD_NAMES:
            assert util.same_bits(models[0].get(n), models[1].get(n)), n
    finally:
        lib().roger_oracle_set_pow(0)


def test_average_maximum_minimum_diagnostics_on_the_device():
    """AVERAGE / MAXIMUM / MINIMUM accumulators (roger/diagnostics/average.py, maximum.py, minimum.py) against their NumPy
    restatement on the oracle engine, over output intervals of one day"""
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import synth
    from roger_b200.svat import SvatModel

    shape = (12, 9)
    forcing = synth.mixed_forcing(8, seed=42, start_doy=121)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, masked_fraction=0.1)
    spec = [("theta_rz", "average"), ("q_ss", "maximum"), ("dS", "minimum"), ("aet", "rate"), ("q_sur", "maximum"), ("z_sat", "average"),
            ("S", "maximum_t"), ("theta_ss", "minimum_t")]  # (x, y, timesteps) variables of the reference: extreme of the last two levels
    lib().roger_oracle_set_pow(1)
    try:
        models = [SvatModel(e, shape, forcing, prim, th_rz, th_ss) for e in (_engine(), OracleSvatEngine)]
        for m in models:
            m.setup()
            m.engine.set_diagnostics(spec)
        for day in range(4):
            while True:
                a, b = models[0].step(True), models[1].step(True)
                if a.time % 86400 == 0:
                    break
            da, db = models[0].engine.download_rates(reset=True), models[1].engine.download_rates(reset=True)
            for (name, mode), x, y in zip(spec, da.values(), db.values()):
                pass
            xa = np.stack(list(models[0].engine.download_rates(reset=False).values()))
            assert np.all(xa == 0)  # reset
            for k in da:
                assert util.same_bits(da[k], db[k]), (day, k)
    finally:
        lib().roger_oracle_set_pow(0)


def test_lazy_soil_conductivity_and_potential_planes():
    """The state-only kernels skip k, h of the whole soil column (results nobody reads inside the step); a download, an async
    snapshot or a plane pointer brings them up to date from theta.  Bit-identical to the oracle, which always evaluates them,
    and to the flux-storing kernel; masked cells included."""
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import synth
    from roger_b200.svat import SvatModel

    lib().roger_oracle_set_pow(1)
    try:
        shape = (16, 16)
        forcing = synth.mixed_forcing(3, seed=3, start_doy=150)
        prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=5, lateral=True, masked_fraction=0.1)
        models = [SvatModel(e, shape, forcing, prim, th_rz, th_ss, lateral=True) for e in (_engine(), OracleSvatEngine)]
        for m in models:
            m.setup()
        for step in range(40):
            for m in models:
                m.step(write_fluxes=(step % 10 == 9))
            if step in (0, 7, 9, 10, 39):
                for n in ("theta", "k", "h"):
                    assert util.same_bits(models[0].get(n), models[1].get(n)), (step, n)
        for m in models:  # day-blocked launches leave the planes stale as well
            m.run_steps(30)
        eng = models[0].engine
        pinned = np.empty(shape[0] * shape[1])
        eng.download_async("k", pinned)
        eng.download_wait()
        assert util.same_bits(pinned, models[1].get("k").reshape(-1))
        assert util.same_bits(models[0].get("h"), models[1].get("h"))
    finally:
        lib().roger_oracle_set_pow(0)


def test_async_diagnostic_output_file(tmp_path):
    """roger_b200.output.AsyncDiagnosticOutput on the CUDA engine (device accumulators, snapshot + copy stream + writer thread):
    the file holds the same values as the one written from the oracle engine, record by record (rate / average / maximum /
    minimum / collect), and every record is complete although steps keep running while it is copied and written."""
    from oracle.engine import OracleSvatEngine, lib
    from scipy.io import netcdf_file

    from tests.test_output import SPEC, run_with_output

    lib().roger_oracle_set_pow(1)
    try:
        pa, pb = str(tmp_path / "cuda.nc"), str(tmp_path / "oracle.nc")
        run_with_output(_engine(), pa, ndays=4, shape=(24, 20))
        run_with_output(OracleSvatEngine, pb, ndays=4, shape=(24, 20))
    finally:
        lib().roger_oracle_set_pow(0)
    with netcdf_file(pa, "r", mmap=False) as fa, netcdf_file(pb, "r", mmap=False) as fb:
        assert np.array_equal(fa.variables["Time"][:], fb.variables["Time"][:]) and fa.variables["Time"].shape == (4,)
        for name, _ in SPEC:
            assert util.same_bits(fa.variables[name][:], fb.variables[name][:]), name


def test_driver_output_files_on_the_cuda_engine(tmp_path):
    """CudaStepDriver.enable_output with the CUDA engine: the files of a setup's rate / average / collect diagnostics equal,
    bit for bit, those the same driver writes with the oracle engine plugged in (which tests/test_integration_reference.py
    checks against the reference's own diagnostics)."""
    import types

    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import synth
    from roger_b200.integration import CudaStepDriver
    from scipy.io import netcdf_file
    from tests.mini_setup import MiniSvatSetup

    shape = (8, 6)
    want = {"rate": ["q_ss", "aet", "prec"], "average": ["theta_rz"], "collect": ["S_s"]}
    forcing = synth.mixed_forcing(6, seed=42, start_doy=140)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, masked_fraction=0.1)
    lib().roger_oracle_set_pow(1)
    try:
        for tag, eng in (("cuda", _engine()), ("oracle", OracleSvatEngine)):
            setup = MiniSvatSetup(shape, forcing, prim, th_rz, th_ss)
            setup.setup()
            st = setup.state
            st.settings.identifier, st.settings.time_origin = "MINI", "2021-12-31 00:00:00"
            st.variables.x, st.variables.y = np.arange(shape[0] + 4, dtype=float), np.arange(shape[1] + 4, dtype=float)
            st.diagnostics = {n: types.SimpleNamespace(output_variables=v, output_frequency=86400, output_path="{identifier}." + n + ".nc",
                                                       base_output_path=None) for n, v in want.items()}
            out_dir = tmp_path / tag
            out_dir.mkdir()
            driver = CudaStepDriver(setup, engine_cls=eng)
            driver.enable_output(base_path=str(out_dir))
            while int(np.asarray(st.variables.time)) < 3 * 86400:
                driver.step()
            driver.close_output()
    finally:
        lib().roger_oracle_set_pow(0)
    for name, vars_ in want.items():
        with netcdf_file(str(tmp_path / "cuda" / f"MINI.{name}.nc"), "r", mmap=False) as fa, \
                netcdf_file(str(tmp_path / "oracle" / f"MINI.{name}.nc"), "r", mmap=False) as fb:
            assert fa.variables["Time"].shape[0] >= 2 and np.array_equal(fa.variables["Time"][:], fb.variables["Time"][:])
            for v in vars_:
                assert util.same_bits(fa.variables[v][:], fb.variables[v][:]), (name, v)


@pytest.mark.parametrize("lateral,dist", [(True, False), (False, True)])
def test_device_cell_order_changes_nothing_but_the_placement(lateral, dist):
    """roger_svat_optimize_cell_order sorts the cells on the device by infiltration capacity; uploads, downloads, per-cell
    forcing blocks, asynchronous plane hand-overs and the diagnostics accumulators keep the caller's cell numbering, and every
    plane stays identical in every bit to a run in the caller's own order (columns are independent)."""
    import torch
    from roger_b200 import abi, synth
    from roger_b200.svat import SvatModel

    shape = (37, 19)  # 703 cells: a ragged last tile
    forcing = synth.mixed_forcing(20, seed=42, start_doy=100)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=47, lateral=lateral, masked_fraction=0.05)
    weights = synth.distributed_weights(shape, seed=48) if dist else None
    models = [SvatModel(_engine(), shape, forcing, prim, th_rz, th_ss, lateral=lateral, weights=weights, cell_order=flag) for flag in (True, False)]
    for m in models:
        m.setup()
        m.engine.set_diagnostics([("q_ss", "rate"), ("theta_rz", "average"), ("S", "maximum")])
    # the order really is a non-trivial permutation: device position 0 holds the cell with the smallest ks * (1 - sealing)
    pins = [torch.empty(models[0].C, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
    for step in range(60):
        a, b = models[0].step(True), models[1].step(True)
        assert (a.time, a.dt_secs, a.itt_day, a.sanity_ok) == (b.time, b.dt_secs, b.itt_day, b.sanity_ok)
        if step % 20 == 19:
            for n in abi.FIELD_NAMES:
                assert util.same_bits(models[0].get(n), models[1].get(n)), (step, n)
            for m, pin in zip(models, pins):
                m.engine.download_async("S_fp_rz", pin)
                m.engine.download_wait()
            assert util.same_bits(pins[0], pins[1])
            ra, rb = models[0].engine.download_rates(reset=True), models[1].engine.download_rates(reset=True)
            for k in ra:
                assert util.same_bits(ra[k], rb[k]), (step, k)
    # a second call re-sorts from the current planes (composition of orders) and an upload after it lands in the right cells
    models[0].engine.optimize_cell_order()
    v = np.arange(models[0].C, dtype=np.float64)
    models[0].engine.upload("z_gw", v)
    assert util.same_bits(models[0].get("z_gw"), v)
    assert util.same_bits(models[0].get("S_lp_ss"), models[1].get("S_lp_ss"))


def test_lazy_mirrors_on_the_cuda_engine():
    """mirror="lazy": the driver downloads a plane only when `vs.<name>` is read; the values read -- after a month change with
    its land-use refresh (which itself reads vs.ta through the mirror) and a user edit of a state array -- equal the eager
    driver's, both on the CUDA engine."""
    from roger_b200 import synth
    from roger_b200.integration import CudaStepDriver
    from tests.mini_setup import MiniSvatSetup

    shape = (24, 16)
    forcing = synth.mixed_forcing(200, seed=42, start_doy=110)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lu_choices=(10, 11, 12, 8, 5, 9), masked_fraction=0.1)
    drivers = []
    for mode in ("eager", "lazy"):
        setup = MiniSvatSetup(shape, forcing, prim, th_rz, th_ss)
        setup.setup()
        drivers.append(CudaStepDriver(setup, engine_cls=_engine(), mirror=mode, fluxes=["q_ss"]))
    months = set()
    for step in range(400):
        if step == 150:  # a user hook that edits a state array on the host: picked up by the mirror, no announcement needed
            for d in drivers:
                vs = d.state.variables
                a = np.array(vs.S_dep)
                a[2:-2, 2:-2, 1] += 0.25
                vs.S_dep = a
        a, b = drivers[0].step(), drivers[1].step()
        assert (a.time, a.dt_secs, a.itt_day) == (b.time, b.dt_secs, b.itt_day)
        months.add(int(drivers[1].state.variables.month[1]))
        if step % 37 == 0 or step == 151:
            for n in ("S", "theta_rz", "q_ss", "S_dep", "S_int_top_tot"):
                x, y = np.asarray(getattr(drivers[0].state.variables, n)), np.asarray(getattr(drivers[1].state.variables, n))
                x = x[2:-2, 2:-2, 1] if x.ndim == 3 else x[2:-2, 2:-2]
                y = y[2:-2, 2:-2, 1] if y.ndim == 3 else y[2:-2, 2:-2]
                assert util.same_bits(x, y), (step, n)
    assert len(months) >= 2
    assert drivers[1].downloads < 0.02 * drivers[0].downloads
    with pytest.raises(RuntimeError, match="fluxes"):
        drivers[1].state.variables.q_hof
