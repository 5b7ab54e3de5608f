"""Generates tests/golden/svat_*.npz by running the UNMODIFIED reference NumPy backend
(/root/reference, read-only) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_svat_golden.py [case ...]

Each fixture holds: the inputs (forcing series, primary parameters, weights), the reference's
derived parameters / initial state right after `setup()`, and for a list of recorded steps the
value of every per-cell field of include/roger_b200_fields.h plus the global scalars, taken
after `RogerSetup.step` returned (i.e. after the user `after_timestep` hook).
"""
import os
import re
import sys

# The fixtures are produced with NumPy's AVX512 / AVX2 SIMD dispatch switched off (a documented NumPy runtime switch): float64
# `power` then calls the C library's pow -- the function the C oracle calls -- and oracle and reference agree bit for bit
# (tests/test_oracle_golden.py).  With NumPy's default SVML `power` the reference differs from itself in the last bit of
# 6 % of all pow calls, which the model's threshold tests occasionally amplify.
NO_SIMD = "AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2 FMA3"
if __name__ == "__main__" and os.environ.get("NPY_DISABLE_CPU_FEATURES") != NO_SIMD:
    os.environ["NPY_DISABLE_CPU_FEATURES"] = NO_SIMD  # must be set before numpy is imported
    os.execv(sys.executable, [sys.executable] + sys.argv)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402
from roger_b200 import synth  # noqa: E402
from tests import util as tutil  # noqa: E402


def field_names():
    txt = open(os.path.join(ROOT, "include", "roger_b200_fields.h")).read()
    return re.findall(r"X\(([PSF]), (\w+)\)", txt)


CASES = {
    # name: (event_type, ndays, shape, heterogeneous, distributed, lateral, start_doy, record_every)
    "svat_rain": dict(event="rain", ndays=10, shape=(3, 2), hetero=False, dist=False, lateral=False, doy=1, every=1),
    "svat_heavyrain": dict(event="heavyrain", ndays=4, shape=(4, 3), hetero=True, dist=False, lateral=False, doy=150, every=5,
                           masked=0.2),
    "svat_snowrain": dict(event="snow+rain", ndays=10, shape=(3, 3), hetero=True, dist=False, lateral=False, doy=20, every=1,
                          lu=(10, 11, 12, 8, 5)),
    "svat_snow": dict(event="snow", ndays=6, shape=(2, 3), hetero=True, dist=False, lateral=False, doy=20, every=1,
                      lu=(10, 11, 12, 8, 5)),
    "svat_mixed_dist": dict(event="mixed", ndays=45, shape=(4, 4), hetero=True, dist=True, lateral=False, doy=45, every=6,
                            masked=0.1),
    "oned_rain": dict(event="rain", ndays=10, shape=(3, 2), hetero=False, dist=False, lateral=True, doy=1, every=1),
    "oned_mixed": dict(event="mixed", ndays=30, shape=(3, 4), hetero=True, dist=False, lateral=True, doy=100, every=5),
    # 256 heterogeneous cells (a full tile of the CUDA layout, 8 warps with divergent cells), per-cell forcing, partly sealed
    # surfaces, masked cells; winter start: snow accumulation, rain on snow, melt.  Full planes are kept for a handful of
    # steps only; EVERY step is pinned through per-field digests of the bit patterns (`step_digest`, tests/util.py).
    "svat_hetero256": dict(event="mixed", ndays=30, shape=(16, 16), hetero=True, dist=True, lateral=False, doy=10, every=10 ** 9,
                           masked=0.05, sealed=0.3, lu=(5, 6, 7, 8, 9, 10, 11, 12, 13), full_records=(1, 2, 100, 200), digest=True),
    "oned_hetero256": dict(event="mixed", ndays=25, shape=(16, 16), hetero=True, dist=True, lateral=True, doy=95, every=10 ** 9,
                           masked=0.05, sealed=0.3, full_records=(1, 2, 100, 200), digest=True),
    # groundwater boundary condition (settings.enable_groundwater_boundary): a water table that moves with time -- in some cells up
    # into the subsoil and back -- capillary rise from it, recharge through the vadose-zone reservoir
    "svat_gwb": dict(event="mixed", ndays=24, shape=(4, 4), hetero=True, dist=True, lateral=False, doy=100, every=4, masked=0.1, gwb=True),
    "oned_gwb": dict(event="mixed", ndays=20, shape=(3, 4), hetero=True, dist=False, lateral=True, doy=95, every=4, gwb=True),
}


def build_case(name):
    c = CASES[name]
    nx, ny = c["shape"]
    if c["event"] == "mixed":
        forc = synth.mixed_forcing(c["ndays"], seed=42, start_doy=c["doy"])
    else:
        forc = synth.toy_forcing(c["event"], c["ndays"], seed=42, start_doy=c["doy"])
    if c["hetero"]:
        kw = {}
        if "lu" in c:
            kw["lu_choices"] = c["lu"]
        params, th_rz, th_ss = synth.heterogeneous_parameters((nx, ny), seed=43, lateral=c["lateral"],
                                                              masked_fraction=c.get("masked", 0.0), **kw)
    else:
        params, th_rz, th_ss = synth.benchmark_parameters((nx, ny), lateral=c["lateral"])
    if c.get("sealed"):  # sealing > 0 (infiltration.py:1563-1666 theta_d, :1767 z_sc, :1420 q_hof): a share of the cells
        rng = np.random.default_rng(46)
        params["sealing"] = np.where(rng.uniform(size=(nx, ny)) < c["sealed"], np.round(rng.uniform(0.05, 0.6, (nx, ny)), 2), 0.0)
    case = dict(nx=nx, ny=ny, ndays=c["ndays"], params=params, theta_rz0=th_rz, theta_ss0=th_ss, lateral=c["lateral"], **forc)
    if c.get("gwb"):
        rng = np.random.default_rng(47)
        params["n0"] = rng.uniform(0.05, 0.3, (nx, ny))
        params["kf"] = rng.uniform(5, 2500, (nx, ny))
        base = np.asarray(params["z_soil"]) / 1000 + np.where(rng.random((nx, ny)) < 0.5, rng.uniform(0.02, 0.4, (nx, ny)), rng.uniform(0.4, 4.0, (nx, ny)))
        base.reshape(-1)[0] += 11.0  # one cell with the water table more than 10 m below the soil
        t = np.arange(c["ndays"] + 1)
        amp = rng.uniform(0.1, 0.6, (nx, ny))
        z = base[:, :, None] - amp[:, :, None] * np.sin(t / 3.0)[None, None, :]
        z[:, :, 0] = base  # numerics.validate_parameters wants the initial water table below the soil
        case["Z_GW"] = z
    if c["dist"]:
        case.update(synth.distributed_weights((nx, ny), seed=44))
    return case, c


def flatten(snap, name):
    a = np.asarray(snap[name])
    if a.ndim == 3:
        a = a[:, :, 1]  # [tau]
    return a.reshape(-1).astype(np.float64)


def run(name):
    case, c = build_case(name)
    Setup = rh.make_svat_setup(case)
    model = Setup()
    model.setup()
    state = model.state
    vs = state.variables
    fields = field_names()
    ref_names = set(k for k, m in state.var_meta.items() if m.active)

    def record():
        snap = rh.snapshot(state)
        rows = []
        for kind, fname in fields:
            if fname == "maskRiverLake":
                rows.append((snap["maskRiver"] | snap["maskLake"]).reshape(-1).astype(np.float64))
            elif fname in ref_names and fname in snap:
                rows.append(flatten(snap, fname))
            else:
                rows.append(np.full(case["nx"] * case["ny"], np.nan))
        return np.stack(rows), snap

    init_fields, init_snap = record()
    # z_gw [tau], prec[taum1] are part of the state contract as well
    scal_names = ["time", "itt", "dt_secs", "itt_day", "event_id_counter", "time_event0", "dt"]

    def scalars():
        d = {k: float(np.asarray(getattr(vs, k))) for k in scal_names}
        d["event_id0"] = float(vs.event_id[0])
        d["event_id1"] = float(vs.event_id[1])
        d["month"] = float(vs.month[1])
        return d

    rec_steps, rec_fields, rec_scal = [], [], []
    all_scal = []
    digests = []
    step = 0
    runlen = state.settings.runlen
    import roger.core.numerics as numerics

    sanity = []
    while vs.time < runlen:
        model.step(state)
        step += 1
        s = scalars()
        all_scal.append([s[k] for k in sorted(s)])
        keep = (step <= 30 or step % c["every"] == 0) if "full_records" not in c else step in c["full_records"]
        if keep or c.get("digest"):
            fld, _ = record()
            if c.get("digest"):
                digests.append(np.array([tutil.field_digest(row) for row in fld], dtype=np.uint64))
            if keep:
                rec_steps.append(step)
                rec_fields.append(fld)
    fld, _ = record()
    if rec_steps[-1] != step:
        rec_steps.append(step)
        rec_fields.append(fld)
    scal_keys = sorted(scalars())

    out = dict(
        field_kinds=np.array([k for k, _ in fields]),
        field_names=np.array([n for _, n in fields]),
        init_fields=init_fields,
        rec_steps=np.array(rec_steps, dtype=np.int64),
        rec_fields=np.stack(rec_fields),
        scal_keys=np.array(scal_keys),
        all_scalars=np.array(all_scal),
        nx=case["nx"], ny=case["ny"], ndays=case["ndays"], lateral=int(case["lateral"]),
        PREC=case["PREC"], TA=case["TA"], PET=case["PET"], YEAR=case["YEAR"], MONTH=case["MONTH"], DOY=case["DOY"],
        theta_rz0=case["theta_rz0"], theta_ss0=case["theta_ss0"],
    )
    if digests:
        out["step_digest"] = np.stack(digests)  # (steps, fields) uint64
        out["digest_valid"] = ~np.isnan(init_fields).all(axis=1)
    for k, v in case["params"].items():
        out["param_" + k] = np.asarray(v)
    for k in ("prec_weight", "ta_offset", "pet_weight", "Z_GW"):
        if k in case:
            out[k] = case[k]
    # every derived per-cell variable after setup (for the host-side parameter derivation tests)
    fnames = set(n for _, n in fields)
    for k, v in init_snap.items():
        a = np.asarray(v)
        if a.ndim >= 2 and a.dtype != object and (not c.get("digest") or k in fnames or k in ("S_dep_tot", "ground_cover", "lai", "z_root_crop")):
            out["setup_" + k] = a
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "steps", step, "recorded", len(rec_steps), "size %.1f KB" % (os.path.getsize(path) / 1024),
          "max dS_num_error", float(np.nanmax(np.asarray(vs.dS_num_error))))


def dump_luts():
    """Numeric look-up tables (parameter DATA, roger/lookuptables.py) needed by the host-side parameter
    derivation; stored as a binary fixture next to the package."""
    rh.import_reference()
    import roger.lookuptables as lut

    path = os.path.join(ROOT, "roger_b200", "data", "lut.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(
        path,
        ILU=np.asarray(lut.ARR_ILU, dtype=np.float64),
        GC=np.asarray(lut.ARR_GC, dtype=np.float64),
        GCM=np.asarray(lut.ARR_GCM, dtype=np.float64),
        IS=np.asarray(lut.ARR_IS, dtype=np.float64),
        RDLU=np.asarray(lut.ARR_RDLU, dtype=np.float64),
        MLMS=np.asarray(lut.ARR_MLMS, dtype=np.float64),
    )
    print("lut", os.path.getsize(path) / 1024, "KB")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    if "lut" in names or not sys.argv[1:]:
        dump_luts()
    for n in names:
        if n != "lut":
            run(n)
