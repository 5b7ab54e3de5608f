"""Seeded synthetic forcing and parameter fields.

Restates the recipes of the reference's roger/tools/make_toy_data.py:18-173 (10-minute
`PREC`, daily `TA`/`PET` forward filled, `PET` = daily / 144) without its file IO, plus the
benchmark's homogeneous parameter set (benchmarks/SVAT_benchmark.py:95-121) and a
heterogeneous variant that respects numerics.validate_parameters_soil
(roger/core/numerics.py:32-105).  Used by the golden-fixture generator, the tests and bench.py.
"""
import numpy as np

SLOTS = 144


def _calendar(ndays, start_doy=1, year=2018):
    mdays = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    months, doys, years = [], [], []
    doy, y = start_doy, year
    for _ in range(ndays):
        d = doy
        mth = 1
        for k, n in enumerate(mdays):
            if d <= n:
                mth = k + 1
                break
            d -= n
        months.append(mth)
        doys.append(doy)
        years.append(y)
        doy += 1
        if doy > 365:
            doy, y = 1, y + 1
    rep = lambda a: np.repeat(np.asarray(a, dtype=np.int64), SLOTS)  # noqa: E731
    return rep(years), rep(months), rep(doys)


def toy_forcing(event_type="rain", ndays=10, seed=42, start_doy=1):
    """make_toy_data.py:31-160.  Returns dict(PREC, TA, PET, YEAR, MONTH, DOY) of 10-min series."""
    rng = np.random.default_rng(seed)
    n = ndays * SLOTS
    prec = np.zeros(n)
    if event_type in ("rain", "snow", "snow+rain"):
        r = rng.uniform(0, 1, 18)
        prec[12 : 12 + 18] = r
        prec[n // 2 : n // 2 + 18] = r
        if event_type == "rain":
            ta = rng.uniform(15, 20, ndays)
            pet = rng.uniform(2, 3, ndays)
        elif event_type == "snow":
            ta = rng.uniform(-3, -1, ndays)
            pet = rng.uniform(1, 2, ndays)
        else:
            ta = rng.uniform(0, 3, ndays)
            ta[:2] = -1
            pet = rng.uniform(1, 2, ndays)
    elif event_type == "heavyrain":
        r = rng.uniform(0.1, 6, 12 * 6)
        prec[12 : 12 + len(r)] = r
        prec[n // 2 : n // 2 + len(r)] = r
        ta = rng.uniform(15, 20, ndays)
        pet = rng.uniform(2, 3, ndays)
    elif event_type == "norain":
        ta = rng.uniform(15, 20, ndays)
        pet = rng.uniform(2, 3, ndays)
    elif event_type == "mixed":
        return mixed_forcing(ndays, seed, start_doy)
    else:
        raise ValueError(event_type)
    year, month, doy = _calendar(ndays, start_doy)
    return dict(PREC=prec, TA=np.repeat(ta, SLOTS), PET=np.repeat(pet, SLOTS) / 24 / 6, YEAR=year, MONTH=month, DOY=doy)


def mixed_forcing(ndays=365, seed=42, start_doy=1):
    """Year-long mix of make_toy_data.py:92-128 (its `rng.randint` calls are broken on a numpy
    Generator; `rng.integers` is used instead): ~100 events per year with Weibull/Poisson
    intensities, sinusoidal air temperature (-5..30 degC) and PET (1.5..5.5 mm/d)."""
    rng = np.random.default_rng(seed)
    n = ndays * SLOTS
    nev = max(int(100 * (ndays / 365)), 2)
    x1 = np.unique(rng.integers(0, n, size=nev))
    prec = np.zeros(n)
    for i, s in enumerate(x1):
        nxt = x1[i + 1] if i + 1 < len(x1) else n
        e = min(int(s + rng.integers(1, max(min(nxt - s, 72), 2))), n)
        lam = rng.weibull(1, 1)[0] * 1.5
        if lam > 5.5:
            prec[s:e] = rng.poisson(lam, e - s) * 0.5
        else:
            prec[s:e] = rng.poisson(lam, e - s) * rng.uniform(0.01, 0.1, e - s) * 4
    scale = np.sin(np.linspace(0, np.pi, 365))
    ta = np.zeros(ndays)
    pet = np.zeros(ndays)
    for i in range(ndays):
        ii = (i + start_doy - 1) % 365
        ta[i] = rng.uniform(-6 + scale[ii] * 35, -4 + scale[ii] * 35)
        pet[i] = rng.uniform(0.5 + scale[ii] * 4, 2.5 + scale[ii] * 4)
    year, month, doy = _calendar(ndays, start_doy)
    return dict(PREC=prec, TA=np.repeat(ta, SLOTS), PET=np.repeat(pet, SLOTS) / 24 / 6, YEAR=year, MONTH=month, DOY=doy)


def benchmark_parameters(shape, lateral=False):
    """Homogeneous parameter set of benchmarks/SVAT_benchmark.py:95-121 / oneD_benchmark.py:102-165."""
    full = lambda v: np.full(shape, v, dtype=np.float64)  # noqa: E731
    p = dict(
        lu_id=np.full(shape, 8, dtype=np.int64),
        z_soil=full(1000.0 if lateral else 2000.0),
        dmpv=full(50.0),
        lmpv=full(600.0 if lateral else 50.0),
        theta_ac=full(0.1),
        theta_ufc=full(0.1),
        theta_pwp=full(0.2),
        ks=full(5.0),
        kf=full(2500.0),
    )
    if lateral:
        p.update(sealing=full(0.0), slope=full(0.05), S_dep_tot=full(0.0), dmph=full(50.0))
    return p, full(0.3), full(0.3)


def heterogeneous_parameters(shape, seed=43, lateral=False, lu_choices=(5, 6, 7, 8, 9, 10, 11, 12, 13), masked_fraction=0.0):
    """Per-cell parameters: every draw respects validate_parameters_soil."""
    rng = np.random.default_rng(seed)
    lu = rng.choice(np.asarray(lu_choices, dtype=np.int64), size=shape)
    if masked_fraction > 0:
        holes = rng.uniform(size=shape) < masked_fraction
        lu = np.where(holes, rng.choice(np.asarray([14, 20, 999], dtype=np.int64), size=shape), lu)
    z_soil = rng.uniform(500, 2500, shape)
    theta_ac = rng.uniform(0.05, 0.25, shape)
    theta_ufc = rng.uniform(0.05, 0.25, shape)
    theta_pwp = rng.uniform(0.05, 0.25, shape)
    p = dict(
        lu_id=lu,
        z_soil=z_soil,
        dmpv=np.floor(rng.uniform(25, 150, shape)),
        lmpv=rng.uniform(100, 0.9 * z_soil),
        theta_ac=theta_ac,
        theta_ufc=theta_ufc,
        theta_pwp=theta_pwp,
        ks=np.exp(rng.uniform(np.log(0.5), np.log(100.0), shape)),
        kf=rng.uniform(5, 2500, shape),
    )
    if lateral:
        p.update(
            sealing=np.zeros(shape),
            slope=np.round(rng.uniform(0.01, 0.3, shape), 2),
            S_dep_tot=np.zeros(shape),
            dmph=np.floor(rng.uniform(10, 100, shape)),
        )
    th0 = theta_pwp + rng.uniform(0.2, 1.0, shape) * theta_ufc
    th1 = theta_pwp + rng.uniform(0.2, 1.0, shape) * theta_ufc
    return p, th0, th1


def distributed_weights(shape, seed=44):
    """Eberbaechle pattern (examples/catchment_scale/eberbaechle/svat_distributed/svat.py:169-186)."""
    rng = np.random.default_rng(seed)
    return dict(
        prec_weight=rng.uniform(0.8, 1.2, shape),
        ta_offset=rng.uniform(-2, 2, shape),
        pet_weight=rng.uniform(0.8, 1.2, shape),
    )
