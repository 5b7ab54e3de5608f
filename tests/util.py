"""Shared helpers of the parity tests: fixture loading and replay through an engine."""
import os

import numpy as np

from roger_b200 import abi
from roger_b200.svat import SvatModel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SVAT_CASES = ["svat_rain", "svat_heavyrain", "svat_snowrain", "svat_snow", "svat_mixed_dist", "svat_hetero256", "svat_gwb"]
ONED_CASES = ["oned_rain", "oned_mixed", "oned_hetero256", "oned_gwb"]


def field_digest(row):
    """64-bit digest of one plane's bit patterns: sum of bits(x_i) * (2 i + 1) mod 2^64 with every NaN folded to one pattern
    and -0.0 to +0.0 (the equality the parity tests use, same_bits).  The large fixtures keep this per field and STEP instead
    of the planes themselves, so that every step of a 256-cell run is pinned to the reference in a few hundred kilobytes."""
    a = np.ascontiguousarray(np.asarray(row, dtype=np.float64).reshape(-1)) + 0.0  # -0.0 + 0.0 = +0.0
    bits = a.view(np.uint64).copy()
    bits[np.isnan(a)] = np.uint64(0x7FF8000000000000)
    w = (np.arange(bits.size, dtype=np.uint64) << np.uint64(1)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return np.uint64((bits * w).sum(dtype=np.uint64))


def load_case(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
        return {k: d[k] for k in d.files}


def model_from_fixture(fx, engine_cls, from_reference_setup=True, engine_kwargs=None):
    """Build a SvatModel on the fixture's inputs.  from_reference_setup=True starts every plane from
    the reference's own post-setup values (isolates the step); False uses roger_b200.params."""
    shape = (int(fx["nx"]), int(fx["ny"]))
    forcing = {k: fx[k] for k in ("PREC", "TA", "PET", "YEAR", "MONTH", "DOY")}
    primary = {k[len("param_"):]: fx[k] for k in fx if k.startswith("param_")}
    weights = None
    if "prec_weight" in fx:
        weights = {k: fx[k] for k in ("prec_weight", "ta_offset", "pet_weight")}
    z_gw = fx["Z_GW"].reshape(shape[0] * shape[1], -1) if "Z_GW" in fx else None  # groundwater boundary condition
    m = SvatModel(engine_cls, shape, forcing, primary, fx["theta_rz0"], fx["theta_ss0"], lateral=bool(fx["lateral"]),
                  weights=weights, engine_kwargs=engine_kwargs, z_gw_daily=z_gw)
    over = None
    if from_reference_setup:
        names = [str(n) for n in fx["field_names"]]
        over = {}
        for n, row in zip(names, fx["init_fields"]):
            if n in abi.FIELD_ID and not np.isnan(row).all():
                over[n] = row
    m.setup(override_planes=over)
    return m


ABS_FLOOR = 1e-12  # mm (or -): below this a value is cancellation residue, e.g. z_sat = 8.7e-19 in the reference


def normalised_max_error(a, b):
    """The reference's own parity idiom (test/restart_test.py:7-17): max |a-b| / max|a|, with an absolute
    floor: differences below ABS_FLOOR are rounding residue of an exact zero and count as equal."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    if (np.isnan(a) != np.isnan(b)).any():
        return np.inf
    a = np.where(both_nan, 0, a)
    b = np.where(both_nan, 0, b)
    scale = np.max(np.abs(a))
    diff = np.abs(a - b)
    diff = np.where(diff <= ABS_FLOOR, 0.0, diff)
    if scale == 0:
        return float(np.max(diff))
    return float(np.max(diff) / scale)


def same_bits(a, b):
    """value equality of every element (NaN == NaN, -0.0 == 0.0)"""
    a = np.asarray(a)
    b = np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def replay_and_compare(fx, model, tol, kinds=("S", "F"), skip=(), write_fluxes=True, collect=None, exact=False):
    names = [str(n) for n in fx["field_names"]]
    kinds_of = {str(n): str(k) for n, k in zip(fx["field_names"], fx["field_kinds"])}
    rec_steps = [int(s) for s in fx["rec_steps"]]
    rec = {s: i for i, s in enumerate(rec_steps)}
    scal_keys = [str(k) for k in fx["scal_keys"]]
    worst = {}
    nsteps = int(rec_steps[-1])
    for step in range(1, nsteps + 1):
        sc = model.step(write_fluxes)
        ref_sc = dict(zip(scal_keys, fx["all_scalars"][step - 1]))
        got = dict(time=sc.time, itt=sc.itt, dt_secs=sc.dt_secs, itt_day=sc.itt_day, event_id_counter=sc.event_id_counter,
                   time_event0=sc.time_event0, dt=sc.dt, event_id0=sc.event_id[0], event_id1=sc.event_id[1], month=sc.month)
        for k, v in ref_sc.items():
            assert float(got[k]) == float(v), f"step {step}: scalar {k}: got {got[k]} ref {v}"
        if exact and "step_digest" in fx:  # every step, every field the reference holds: digest of the bit patterns
            for j, n in enumerate(names):
                if kinds_of[n] in kinds and n not in skip and fx["digest_valid"][j]:
                    assert field_digest(model.get(n)) == fx["step_digest"][step - 1][j], \
                        f"step {step}: digest of {n} differs from the reference fixture"
        if step in rec:
            ref = fx["rec_fields"][rec[step]]
            for n, row in zip(names, ref):
                if kinds_of[n] not in kinds or n in skip or np.isnan(row).all():
                    continue
                if exact:
                    assert same_bits(row, model.get(n)), f"step {step}: {n} differs from the reference fixture"
                    continue
                if n.endswith("_num_error"):
                    # closure residuals are O(1e-14) rounding noise: compared absolutely (mm)
                    err = float(np.nanmax(np.abs(row - model.get(n)))) * 1e-1
                else:
                    err = normalised_max_error(row, model.get(n))
                if err > worst.get(n, (0, 0))[0]:
                    worst[n] = (err, step)
    bad = {n: v for n, v in worst.items() if v[0] > tol}
    if collect is not None:
        collect.update(worst)
    assert not bad, "fields above tolerance %g: %s" % (tol, sorted(bad.items(), key=lambda kv: -kv[1][0])[:12])
    return worst
