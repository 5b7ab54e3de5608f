#!/usr/bin/env python
"""Generates roger_b200/csrc/roger_pow_tables.h: the constants of roger_pow() (roger_b200/csrc/roger_pow.h).

roger_pow(x, y) = exp(y * log(x)) with log(x) carried as an unevaluated sum hi + lo (about 2^-66 relative),
table-driven in the style published for the ARM optimized routines / glibc 2.28+ `pow` (log: 128 sub-intervals
of [0.7055, 1.411) with 1/c on a 2^-7 / 2^-8 grid so that r = z/c - 1 is exact; exp: 2^(i/128) as scale bits +
relative tail).  The two tables are derived here with mpmath at 200 bits (they come out identical to the published ones,
entry for entry).  The polynomial coefficients and the split of ln2/N are the PUBLISHED constants of that algorithm (ARM
optimized routines, math/pow_log_data.c and math/exp_data.c, MIT licence; shipped as glibc's pow / exp since 2.28, here
glibc 2.39): with them, and with the operation order of roger_pow.h, roger_pow(x, y) and roger_exp(x) return the bits
of the C library's pow and exp -- the functions NumPy's float64 `power` / `exp` loops call when its SIMD dispatch is off,
i.e. what the reference fixtures were produced with.  tests/test_pow.py checks that equality on 4e7 arguments.
`--own-fit` writes coefficients fitted here instead (results then differ from the C library's in the last bit of ~0.06 %
of the evaluations; round 1 of this backend used those).

    python tools/gen_pow_tables.py            # rewrites the header
"""
import sys

# math/pow_log_data.c (poly[1..6] scaled as the code uses them: A1 = -2 * 0x1.555555555556p-2, ...) and math/exp_data.c
PUBLISHED_A = ["-0x1.555555555556p-1", "0x1.0000000000006p-1", "0x1.999999959554ep-1", "-0x1.555555529a47ap-1", "-0x1.2495b9b4845e9p+0",
               "0x1.0002b8b263fc3p+0"]
PUBLISHED_C = ["0x1.ffffffffffdbdp-2", "0x1.555555555543cp-3", "0x1.55555cf172b91p-5", "0x1.1111167a4d017p-7"]
PUBLISHED_NEGLN2HIN, PUBLISHED_NEGLN2LON = "-0x1.62e42fefa0000p-8", "-0x1.cf79abc9e3b3ap-47"
import os
import struct
from fractions import Fraction

import mpmath as mp

mp.mp.prec = 240
N = 128
OFF = 0x3FE6955500000000
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "roger_b200", "csrc", "roger_pow_tables.h")


def bits2d(b):
    return struct.unpack("<d", struct.pack("<Q", b & 0xFFFFFFFFFFFFFFFF))[0]


def d2bits(d):
    return struct.unpack("<Q", struct.pack("<d", d))[0]


def rnd(x):
    """mpf -> nearest double"""
    return float(mp.mpf(x))


def hexf(d):
    return float(d).hex()


def log_table():
    rows = []
    worst_r = Fraction(0)
    for i in range(N):
        lo = bits2d(OFF + (i << 45))
        hi = bits2d(OFF + ((i + 1) << 45) - 1)  # last double of the sub-interval
        if lo <= 1.0 <= hi:
            invc = Fraction(1)
        else:
            centre = (Fraction(lo) + Fraction(hi)) / 2
            den = 128 if centre < 1 else 256
            n = round(den / centre)
            invc = Fraction(n, den)
        # exactness of r = z * invc - 1 at the interval ends: multiple of the ulp grid and |r| < 2^-7
        for z in (lo, hi):
            r = Fraction(z) * invc - 1
            assert abs(r) < Fraction(1, 128), (i, float(r))
            assert Fraction(float(r)) == r, (i, "r not representable")
            worst_r = max(worst_r, abs(r))
        c = 1 / invc
        logc = mp.log(mp.mpf(c.numerator) / mp.mpf(c.denominator))
        logc_hi = float(mp.nint(logc * mp.mpf(2) ** 43) / mp.mpf(2) ** 43) if invc != 1 else 0.0
        logc_lo = rnd(logc - mp.mpf(logc_hi)) if invc != 1 else 0.0
        rows.append((float(invc), logc_hi, logc_lo))
    return rows, float(worst_r)


def cheb_fit(f, a, b, deg):
    """polynomial (monomial coefficients, low -> high) interpolating f at deg+1 Chebyshev nodes of [a, b]"""
    n = deg + 1
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    y = mp.matrix(n, 1)
    for r, x in enumerate(xs):
        for c in range(n):
            A[r, c] = x ** c
        y[r] = f(x)
    sol = mp.lu_solve(A, y)
    return [sol[k] for k in range(n)]


def main():
    rows, worst_r = log_table()
    rmax = mp.mpf(worst_r) * mp.mpf("1.0001")

    # log1p(r) = r - r^2/2 + r^3 * g(r);  g fitted with degree 5
    def g(r):
        if r == 0:
            return mp.mpf(1) / 3
        return (mp.log1p(r) - r + r * r / 2) / r ** 3

    c = cheb_fit(g, -rmax, rmax, 5)
    # p = ar3 * (A1 + r*A2 + ar2*(A3 + r*A4 + ar2*(A5 + r*A6))), ar2 = -r^2/2, ar3 = -r^3/2
    A = [-0.5, rnd(-2 * c[0]), rnd(-2 * c[1]), rnd(4 * c[2]), rnd(4 * c[3]), rnd(-8 * c[4]), rnd(-8 * c[5])]
    own = "--own-fit" in sys.argv
    if not own:
        A = [-0.5] + [float.fromhex(v) for v in PUBLISHED_A]
        c = [mp.mpf(A[1]) / -2, mp.mpf(A[2]) / -2, mp.mpf(A[3]) / 4, mp.mpf(A[4]) / 4, mp.mpf(A[5]) / -8, mp.mpf(A[6]) / -8]
    # error of the fitted log1p polynomial relative to log1p(r) itself
    worst = mp.mpf(0)
    for k in range(-400, 401):
        r = rmax * k / 400
        if r == 0:
            continue
        approx = r - r * r / 2 + r ** 3 * sum(c[j] * r ** j for j in range(6))
        worst = max(worst, abs(approx / mp.log1p(r) - 1))
    ln2 = mp.log(2)
    ln2hi = float(mp.nint(ln2 * mp.mpf(2) ** 42) / mp.mpf(2) ** 42)
    ln2lo = rnd(ln2 - mp.mpf(ln2hi))

    # exp: x = k*ln2/N + r;  2^(i/N) = double(scale_i) * (1 + tail_i)
    invln2N = rnd(N / ln2)
    q = ln2 / N
    negln2hiN = -float(mp.nint(q * mp.mpf(2) ** 41) / mp.mpf(2) ** 41)  # |kd| < 2^18 -> kd*hi exact
    negln2loN = -rnd(q + mp.mpf(negln2hiN))
    if not own:
        negln2hiN, negln2loN = float.fromhex(PUBLISHED_NEGLN2HIN), float.fromhex(PUBLISHED_NEGLN2LON)
    # e^r - 1 - r = r^2 * e2(r) on |r| <= ln2/(2N) * 1.01, degree 3
    rm = q / 2 * mp.mpf("1.01")

    def e2(r):
        if r == 0:
            return mp.mpf(1) / 2
        return (mp.expm1(r) - r) / (r * r)

    ce = cheb_fit(e2, -rm, rm, 3)
    C = [rnd(v) for v in ce]
    if not own:
        C = [float.fromhex(v) for v in PUBLISHED_C]
        ce = [mp.mpf(v) for v in C]
    worst_e = mp.mpf(0)
    for k in range(-200, 201):
        r = rm * k / 200
        approx = 1 + r + r * r * sum(ce[j] * r ** j for j in range(4))
        worst_e = max(worst_e, abs(approx / mp.exp(r) - 1))
    exptab = []
    for i in range(N):
        v = mp.mpf(2) ** (mp.mpf(i) / N)
        h = rnd(v)
        tail = rnd(v / mp.mpf(h) - 1)
        sbits = (d2bits(h) - (i << 45)) & 0xFFFFFFFFFFFFFFFF
        exptab.append((d2bits(tail), sbits))

    with open(OUT, "w") as f:
        w = f.write
        w("/* GENERATED by tools/gen_pow_tables.py%s -- do not edit.  Constants of roger_pow() (roger_pow.h).\n" % (" --own-fit" if own else ""))
        w(" * Tables derived with mpmath; polynomial coefficients and the ln2/N split %s.\n"
          % ("fitted by the generator" if own else "as published for the ARM optimized routines / glibc >= 2.28 pow and exp"))
        w(" * log: max |r| = %.6g, fitted log1p relative error %.3g; exp: fitted e^r relative error %.3g */\n"
          % (worst_r, float(worst), float(worst_e)))
        w("#ifndef ROGER_POW_TABLES_H\n#define ROGER_POW_TABLES_H\n")
        w("#define ROGER_POW_N 128\n")
        w("#define ROGER_POW_OFF 0x%016xULL\n" % OFF)
        w("#define ROGER_POW_LN2HI %s\n#define ROGER_POW_LN2LO %s\n" % (hexf(ln2hi), hexf(ln2lo)))
        for k in range(1, 7):
            w("#define ROGER_POW_A%d %s\n" % (k, hexf(A[k])))
        w("#define ROGER_POW_INVLN2N %s\n#define ROGER_POW_NEGLN2HIN %s\n#define ROGER_POW_NEGLN2LON %s\n"
          % (hexf(invln2N), hexf(negln2hiN), hexf(negln2loN)))
        for k in range(4):
            w("#define ROGER_POW_C%d %s\n" % (k + 2, hexf(C[k])))
        w("/* {1/c, log(c) rounded to 2^-43, log(c) - that} per sub-interval */\n")
        w("#define ROGER_POW_LOG_TABLE \\\n")
        w(" \\\n".join("    {%s, %s, %s}," % tuple(hexf(v) for v in row) for row in rows))
        w("\n/* {bits of the relative tail, bits of 2^(i/N) minus (i << 45)} */\n")
        w("#define ROGER_POW_EXP_TABLE \\\n")
        w(" \\\n".join("    {0x%016xULL, 0x%016xULL}," % row for row in exptab))
        w("\n#endif\n")
    print("wrote", OUT, "max|r| %.6g" % worst_r, "log1p fit %.3g" % float(worst), "exp fit %.3g" % float(worst_e))


if __name__ == "__main__":
    main()
