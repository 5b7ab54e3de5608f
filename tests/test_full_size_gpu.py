"""BASELINE.json's full sizes on the GPU, checked through size-independent properties.

Columns are independent, so a grid built by repeating a small heterogeneous tile must reproduce the tile's own run in
every replica, bit for bit -- at 4096 x 4096 cells (oneD, configs[1]) and at 1024 x 1024 cells x 1500 age bins (SVAT +
oxygen-18 SAS, configs[2]) -- and the tile itself is compared with the CPU oracle.  Water balance closure is checked on
the full grid as well."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TILE = (32, 32)


def _free_gb():
    import torch

    free, _ = torch.cuda.mem_get_info()
    return free / 1e9


def test_oned_4096x4096_replicates_its_tile_and_closes_the_water_balance():
    if _free_gb() < 40:
        pytest.skip("needs 40 GB of free device memory")
    from oracle.engine import OracleSvatEngine, lib
    from roger_b200 import synth
    from roger_b200.engine import CudaSvatEngine
    from roger_b200.svat import SvatModel
    from tests.util import same_bits

    lib().roger_oracle_set_pow(1)  # kernel and oracle evaluate the same pow: bit-for-bit comparison

    side, nsteps = 4096, 30
    forcing = synth.mixed_forcing(12, seed=42, start_doy=121)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(TILE, seed=43, lateral=True)
    nt = TILE[0] * TILE[1]
    C = side * side
    idx = np.arange(C) % nt
    big_prim = {k: np.asarray(v).reshape(-1)[idx] for k, v in prim.items()}
    big = SvatModel(CudaSvatEngine, (side, side), forcing, big_prim, np.asarray(th_rz).reshape(-1)[idx], np.asarray(th_ss).reshape(-1)[idx],
                    lateral=True)
    tile = SvatModel(CudaSvatEngine, TILE, forcing, prim, th_rz, th_ss, lateral=True)
    ref = SvatModel(OracleSvatEngine, TILE, forcing, prim, th_rz, th_ss, lateral=True)
    for m in (big, tile, ref):
        m.setup()
    tile.engine.set_block_stepping(False)
    S0 = big.get("S")
    for _ in range(nsteps):
        a, b, c = big.step(False), tile.step(False), ref.step(False)
        assert (a.time, a.dt_secs, a.itt_day, a.event_id[1]) == (b.time, b.dt_secs, b.itt_day, b.event_id[1]) == (c.time, c.dt_secs, c.itt_day, c.event_id[1])
    for name in ("S_fp_rz", "S_lp_rz", "S_fp_ss", "S_lp_ss", "z_sat", "theta_rz", "S", "de", "z_wf", "k_ss"):
        x = big.get(name)
        t = tile.get(name)
        assert np.array_equal(x, t[idx], equal_nan=True), name  # every replica == the tile
        assert same_bits(ref.get(name), t), name  # the tile == the oracle
    assert big.engine.get_scalars().sanity_ok == 1
    # closure of the last step on the full grid: the kernel's own balance residual (numerics.py:303-346) stays at rounding level
    big.step(True)
    err = big.get("dS_num_error")
    assert float(np.nanmax(err)) < 1e-9
    assert np.isfinite(big.get("S")).all() and not np.array_equal(big.get("S"), S0)
    lib().roger_oracle_set_pow(0)


def test_sas_1024x1024x1500_replicates_its_tile():
    if _free_gb() < 90:
        pytest.skip("needs 90 GB of free device memory")
    from oracle.engine import OracleSasEngine, lib
    from roger_b200 import sas
    from roger_b200.sas_engine import CudaSasEngine
    from tests.test_sas_gpu import _random_case

    A, C, nt, ndays = 1500, 1024 * 1024, 256, 2
    daily, spec = _random_case(nt, A, ndays, seed=77)
    idx = np.arange(C) % nt
    c0 = float(sas.delta_to_conc(-10.0))
    big = CudaSasEngine(C, sas.default_settings(A))
    tile = CudaSasEngine(nt, sas.default_settings(A))
    lib().sas_oracle_set_pow(1)
    try:
        ref = OracleSasEngine(nt, sas.default_settings(A))
        for eng, sel in ((big, idx), (tile, slice(None))):
            eng.set_parameters(np.ones(len(idx) if eng is big else nt), sas.sas_parameter_block(len(idx) if eng is big else nt, spec))
            eng.init_linear(daily["S_RZ"][sel, 0], daily["S_SS"][sel, 0], c0)
        ref.set_parameters(np.ones(nt), sas.sas_parameter_block(nt, spec))
        for k in sas.STATE:
            ref.upload_state(k, tile.download_state(k))
        driver = sas.SasModel(lambda *a, **k: None, nt, A, daily, spec)  # host-side set_forcing only
        for day in range(ndays):
            driver.itt += 1
            f = driver._set_forcing()
            driver.C_snow[:, 0] = driver.C_snow[:, 1]
            big.step(np.ascontiguousarray(f[:, idx]))
            tile.step(f)
            ref.step(f)
            ob, ot, orf = big.download_outputs(), tile.download_outputs(), ref.download_outputs()
            same = (ob == ot[:, idx]) | (np.isnan(ob) & np.isnan(ot[:, idx]))
            assert same.all(), day
            assert ((ot == orf) | (np.isnan(ot) & np.isnan(orf))).all(), day
        # the age-ranked state of a sample of replicas, read back through the (cell, age) view of the C ABI
        st = tile.download_state("sa_ss")
        assert np.array_equal(st, ref.download_state("sa_ss"), equal_nan=True)
        mass_tile = st.sum(axis=1)
        out_big = big.download_outputs()
        assert np.isfinite(out_big[sas.OUTPUTS.index("C_rz")]).all()
        assert np.all(mass_tile > 0)
    finally:
        lib().sas_oracle_set_pow(0)
        big.close()
        tile.close()
