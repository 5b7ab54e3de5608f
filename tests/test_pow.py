"""roger_pow / roger_exp (roger_b200/csrc/roger_pow.h), host build: accuracy against mpmath and the C library."""
import ctypes as C
import math

import numpy as np

from oracle.engine import lib

dp = C.POINTER(C.c_double)


def host_pow(x, y):
    L = lib()
    L.roger_pow_host_array.argtypes = [dp, dp, dp, C.c_int64]
    L.roger_pow_host_array.restype = None
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.broadcast_to(y, x.shape), dtype=np.float64)
    out = np.empty_like(x)
    L.roger_pow_host_array(x.ctypes.data_as(dp), y.ctypes.data_as(dp), out.ctypes.data_as(dp), x.size)
    return out


def test_bit_identical_to_the_c_librarys_pow_and_exp():
    """roger_pow.h restates the algorithm of the C library's pow / exp (glibc >= 2.28, the FMA build the loader selects on an
    FMA-capable x86-64 CPU) operation by operation with the published constants: for |y log x| < 511 (results between 2^-737
    and 2^737) the bits are the library's.  NumPy's float64 `power` / `exp` loops call exactly these functions when its SIMD
    dispatch is off, which is how the reference fixtures were produced -- so results computed with roger_pow can be compared
    with those fixtures bit for bit."""
    L = lib()
    L.roger_pow_compare_libm.restype = C.c_int64
    L.roger_pow_compare_libm.argtypes = [C.c_int64, C.c_uint64, C.POINTER(C.c_int64)]
    bade = C.c_int64(-1)
    bad = L.roger_pow_compare_libm(40_000_000, 12345, C.byref(bade))
    assert (bad, bade.value) == (0, 0), f"{bad} pow and {bade.value} exp results differ from the C library's (glibc version / non-FMA CPU?)"


def test_within_one_ulp_of_libm_and_mostly_identical():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 1, 100000), np.exp(rng.uniform(-200, 200, 100000))])
    y = np.concatenate([rng.choice([0.2, 0.5, 1.5, 1 / 3, 3.0, 0.01, 7.3], 100000), rng.uniform(-2, 2, 100000)])
    a = host_pow(x, y)
    b = np.array([math.pow(v, w) for v, w in zip(x, y)])
    ok = (b > 1e-300) & np.isfinite(b)
    ulp = np.abs(a[ok] - b[ok]) / np.spacing(b[ok])
    assert ulp.max() <= 1.0
    assert (ulp > 0).mean() < 2e-3


def test_error_against_exact_value_below_0_53_ulp():
    import mpmath as mp

    mp.mp.prec = 200
    rng = np.random.default_rng(2)
    x = rng.uniform(0, 1, 4000)
    y = rng.choice([0.2, 0.5, 1.5, 2.7], 4000)
    a = host_pow(x, y)
    worst = 0.0
    for v, k, r in zip(x, y, a):
        t = mp.mpf(float(v)) ** mp.mpf(float(k))
        worst = max(worst, float(abs((mp.mpf(float(r)) - t) / mp.mpf(float(np.spacing(r))))))
    assert worst < 0.53


def test_special_values():
    L = lib()
    L.roger_pow_host.restype = C.c_double
    L.roger_pow_host.argtypes = [C.c_double, C.c_double]
    f = L.roger_pow_host
    inf, nan = float("inf"), float("nan")
    assert f(0.0, 0.5) == 0.0 and f(0.0, -1.0) == inf and f(0.0, 0.0) == 1.0
    assert f(1.0, nan) == 1.0 and f(nan, 0.0) == 1.0 and math.isnan(f(nan, 1.0)) and math.isnan(f(-1.0, 0.5))
    assert f(2.0, 1024.0) == inf and f(2.0, -1080.0) == 0.0 and f(inf, -1.0) == 0.0 and f(0.5, inf) == 0.0
    assert f(5e-324, 0.5) == math.pow(5e-324, 0.5)
    assert f(4.0, 0.5) == 2.0 and f(0.25, 1.5) == 0.125 and f(3.0, 2.0) == 9.0
    L.roger_exp_host.restype = C.c_double
    L.roger_exp_host.argtypes = [C.c_double]
    assert L.roger_exp_host(0.0) == 1.0 and L.roger_exp_host(-1e4) == 0.0 and L.roger_exp_host(1e4) == inf
    rng = np.random.default_rng(3)
    x = rng.uniform(-50, 5, 20000)
    a = np.array([L.roger_exp_host(v) for v in x])
    b = np.exp(x)
    assert (np.abs(a - b) / np.spacing(b)).max() <= 1.0


def test_pow2_returns_the_bits_of_two_separate_powers():
    L = lib()
    L.roger_pow2_host.argtypes = [C.c_double, C.c_double, C.c_double, dp, dp]
    L.roger_pow2_host.restype = None
    L.roger_pow_host.restype = C.c_double
    L.roger_pow_host.argtypes = [C.c_double, C.c_double]
    rng = np.random.default_rng(4)
    xs = np.concatenate([rng.uniform(0, 2, 20000), [0.0, 1.0, np.inf, np.nan, -1.0, 5e-324, 1e-310]])
    for x in xs:
        y1, y2 = rng.uniform(-3, 3), rng.choice([0.0, 0.5, 2.0, np.inf, -np.inf, np.nan, 7.3])
        r1, r2 = C.c_double(), C.c_double()
        L.roger_pow2_host(float(x), float(y1), float(y2), C.byref(r1), C.byref(r2))
        for got, y in ((r1.value, y1), (r2.value, y2)):
            want = L.roger_pow_host(float(x), float(y))
            assert got == want or (math.isnan(got) and math.isnan(want)), (x, y, got, want)
