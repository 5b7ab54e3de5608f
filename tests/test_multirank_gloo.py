"""N > 1 path on CPU (gloo, world_size 2): contiguous cell-block partitioning + the flag OR collective.

Two processes each step their own block of cells (per-cell forcing, a snow period, heterogeneous parameters) and
exchange only the scheduler's flag words through roger_b200.parallel.make_flag_reducer; the result must equal the
single-process run of the whole grid exactly -- the property the multi-GPU run relies on.  The engine here is the CPU
oracle (same engine interface and reducer contract as the CUDA engine)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("S_fp_rz", "S_lp_rz", "S_fp_ss", "S_lp_ss", "swe", "swe_top", "S_snow", "theta_rz", "z_sat", "S", "de", "z_wf", "prec", "ta")


def _setup(shape=(6, 5), ndays=60):
    from roger_b200 import synth

    forcing = synth.mixed_forcing(ndays + 1, seed=42, start_doy=45)
    prim, th_rz, th_ss = synth.heterogeneous_parameters(shape, seed=43, lu_choices=(10, 11, 12, 8, 5))
    weights = synth.distributed_weights(shape, seed=44)
    return forcing, prim, th_rz, th_ss, weights


def _run(lo, hi, reducer, nsteps, q=None):
    from oracle.engine import OracleSvatEngine
    from roger_b200.svat import SvatModel

    forcing, prim, th_rz, th_ss, weights = _setup()
    sl = slice(lo, hi)
    n = hi - lo
    prim = {k: np.asarray(v).reshape(-1)[sl] for k, v in prim.items()}
    weights = {k: np.asarray(v).reshape(-1)[sl] for k, v in weights.items()}
    m = SvatModel(OracleSvatEngine, (n, 1), forcing, prim, np.asarray(th_rz).reshape(-1)[sl], np.asarray(th_ss).reshape(-1)[sl],
                  weights=weights, engine_kwargs={"flag_reducer": reducer})
    m.setup()
    scal = []
    for _ in range(nsteps):
        sc = m.step(False)
        scal.append((sc.time, sc.dt_secs, sc.itt_day, sc.event_id[1], sc.time_event0, sc.event_id_counter))
    return scal, {f: m.get(f) for f in FIELDS}


def _worker(rank, world, port, nsteps, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from roger_b200.parallel import cell_block, make_flag_reducer

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = cell_block(30, rank, world)
        scal, fields = _run(lo, hi, make_flag_reducer(dist), nsteps)
        np.savez(os.path.join(out, f"rank{rank}.npz"), scal=np.asarray(scal, dtype=np.float64), lo=lo, hi=hi, **fields)
    finally:
        dist.destroy_process_group()


def test_cell_block_partition():
    from roger_b200.parallel import cell_block

    for n, w in ((30, 2), (31, 4), (7, 8), (1 << 20, 8)):
        blocks = [cell_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank(tmp_path):
    import torch.multiprocessing as mp

    nsteps = 120
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, nsteps, str(tmp_path)), nprocs=2, join=True)
    scal1, fields1 = _run(0, 30, None, nsteps)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(2)]
    for p in parts:
        assert np.array_equal(p["scal"], np.asarray(scal1, dtype=np.float64))  # identical dt / event decisions on every rank
    for f in FIELDS:
        got = np.concatenate([p[f] for p in parts])
        assert np.array_equal(got, fields1[f], equal_nan=True), f
    # the partition matters: without the collective the ranks would take different time steps
    dts = np.asarray(scal1)[:, 1]
    assert len(set(dts.tolist())) > 1


@pytest.mark.parametrize("case", ["sas_power_stats", "sas_virtualtracer"])
def test_sas_cell_blocks_equal_the_whole_grid(case):
    """The offline transport shards with no exchange at all: every rank steps its contiguous block of cells,
    and the blocks put together are the whole grid bit for bit -- the reference fixture's cells, cut as `cell_block` cuts them."""
    from oracle.engine import OracleSasEngine, lib
    from roger_b200.parallel import cell_block
    from tests import util_sas as u

    lib().sas_oracle_set_pow(0)
    fx = u.load_case(case)
    C = int(fx["nx"]) * int(fx["ny"])
    per_cell = lambda v: isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == C  # noqa: E731

    def block(lo, hi):
        sub = {}
        for k, v in fx.items():
            if k.startswith("rec_") or k.startswith("simd_"):
                continue
            sub[k] = v[lo:hi] if per_cell(v) else v
        sub["nx"], sub["ny"] = np.asarray(hi - lo), np.asarray(1)
        return u.model_from_fixture(sub, OracleSasEngine)

    blocks = [cell_block(C, r, 2) for r in range(2)]
    models = [block(lo, hi) for lo, hi in blocks]
    for day in range(int(fx["ndays"])):
        snaps = []
        for m in models:
            m.step()
            snaps.append(u.snapshot(m))
        for k in snaps[0]:
            if "rec_" + k not in fx:
                continue
            got = np.concatenate([s[k] for s in snaps])
            assert u.same(got, fx["rec_" + k][day]), (day, k)
