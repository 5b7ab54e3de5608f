/* roger_pow.h -- pow(x, y) as a fixed sequence of IEEE-754 binary64 operations (+, -, *, fma, integer ops and two
 * table look-ups), written once and compiled unchanged by nvcc (device) and gcc (host).
 *
 * Why it exists.  The StorAge-selection step raises (SA / S) to a per-cell exponent for every age bin
 * (roger/core/sas.py:194-240, :38-134) and then takes hard decisions on the results (`flux * tt > sa`, `SA > 0`,
 * roger/core/transport.py:431, :494; sas.py:232).  With power-law exponents below one those decisions amplify a
 * last-bit difference of pow into 1e-6 relative differences of the age distributions within days (DESIGN.md,
 * "SAS parity chain"), so the reference's results are only defined up to the libm in use (NumPy's SVML loops
 * and glibc already disagree with each other at that level).  A device pow that is bit-identical to a host pow
 * removes the libm from the comparison: the CUDA kernels and the CPU oracle (when told to) evaluate exactly the
 * same operation sequence and their results can be compared bit for bit.
 *
 * Algorithm: the table-driven pow published with the ARM optimized routines (math/pow.c, MIT licence) and shipped as the
 * C library's pow since glibc 2.28 (pinned here: glibc 2.39, x86-64, the FMA build the dynamic loader selects on any
 * FMA-capable CPU):
 *   log(x) = k ln2 + log(c_i) + log1p(r),  r = z / c_i - 1 exact by fma, carried as hi + lo (~2^-66 relative);
 *   exp(y hi + y lo) = 2^(k/N) e^r with 2^(i/N) from a table as scale bits + relative tail.
 * The operation sequence below -- which products are fused with the following sum and which are rounded separately -- is
 * that of the library's FMA build (read off its compiled code), the constants are the published ones
 * (tools/gen_pow_tables.py): roger_pow(x, y) and roger_exp(x) return the BITS of the C library's pow(x, y) and exp(x)
 * for results in the normal range (tests/test_pow.py: 4e7 arguments each, no difference).  That is the function NumPy's
 * float64 `power` / `exp` loops call when its SIMD dispatch is off -- what the reference fixtures were produced with --
 * so the CUDA kernels reproduce those fixtures bit for bit, not only the CPU oracle.
 * Error against the exactly rounded result: < 0.52 ulp.
 *
 * Domain: the general real power for x >= 0: x^0 = 1, 0^y = 0 / 1 / +inf, results above DBL_MAX -> +inf, results
 * below 2^-1022 are flushed to 0 (not gradual), negative or NaN base -> NaN (the SAS functions never raise a
 * negative base to a non-integer power without discarding the result).
 */
#ifndef ROGER_POW_H
#define ROGER_POW_H

#include <stdint.h>
#include "roger_pow_tables.h"

#if defined(__CUDACC__)
#define ROGER_POW_FN __host__ __device__ __forceinline__
#else
#include <math.h>
#include <string.h>
#define ROGER_POW_FN static inline
#endif

/* Polynomial and reduction constants.  The device build reads them from the constant bank: an FP64 instruction takes a
 * constant-bank operand directly, whereas a literal costs two moves to build the 64-bit immediate every time register
 * pressure evicts it (measured: 25 of 157 instructions per evaluation).  Same values either way. */
#if defined(__CUDACC__)
struct roger_pow_consts_t {
    double ln2hi, ln2lo, a1, a2, a3, a4, a5, a6, invln2n, negln2hin, negln2lon, c2, c3, c4, c5, shift, ovf, unf;
};
static __constant__ roger_pow_consts_t roger_pow_kc = {
    ROGER_POW_LN2HI, ROGER_POW_LN2LO, ROGER_POW_A1, ROGER_POW_A2, ROGER_POW_A3, ROGER_POW_A4, ROGER_POW_A5, ROGER_POW_A6,
    ROGER_POW_INVLN2N, ROGER_POW_NEGLN2HIN, ROGER_POW_NEGLN2LON, ROGER_POW_C2, ROGER_POW_C3, ROGER_POW_C4, ROGER_POW_C5,
    0x1.8p52, 0x1.62e42fefa39efp+9, -0x1.6232bdd7abcd2p+9};
#endif
#if defined(__CUDA_ARCH__)
#define RPK(field, literal) (roger_pow_kc.field)
#else
#define RPK(field, literal) (literal)
#endif

typedef struct { double invc, logc, logctail; } roger_pow_log_entry;
typedef struct { uint64_t tail, sbits; } roger_pow_exp_entry;

/* The tables are passed in by the caller: the CUDA kernels stage them in shared memory, the host keeps statics. */
typedef struct {
    const roger_pow_log_entry *log_tab;
    const roger_pow_exp_entry *exp_tab;
} roger_pow_tables;

#define ROGER_POW_DEFINE_HOST_TABLES(qual)                                                     \
    qual roger_pow_log_entry roger_pow_log_tab_h[ROGER_POW_N] = {ROGER_POW_LOG_TABLE};         \
    qual roger_pow_exp_entry roger_pow_exp_tab_h[ROGER_POW_N] = {ROGER_POW_EXP_TABLE};

ROGER_POW_FN uint64_t roger_pow_asu(double d)
{
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, sizeof u);
    return u;
#endif
}
ROGER_POW_FN double roger_pow_asd(uint64_t u)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, sizeof d);
    return d;
#endif
}
ROGER_POW_FN double roger_pow_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
/* every product / sum below is a separately rounded operation: the host build uses -ffp-contract=off, the
 * device build -fmad=false, and the explicit helpers keep that true under any flags on the device */
ROGER_POW_FN double roger_pow_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
ROGER_POW_FN double roger_pow_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

/* log(x) for positive normal x given as bits; returns hi, *tail = lo */
ROGER_POW_FN double roger_pow_log(uint64_t ix, double *tail, const roger_pow_log_entry *T)
{
    const uint64_t tmp = ix - ROGER_POW_OFF;
    const int i = (int)((tmp >> (52 - 7)) & (ROGER_POW_N - 1));
    const int k = (int)((int64_t)tmp >> 52);
    const uint64_t iz = ix - (tmp & (0xfffULL << 52));
    const double z = roger_pow_asd(iz);
    const double kd = (double)k;
    const double invc = T[i].invc, logc = T[i].logc, logctail = T[i].logctail;
    const double r = roger_pow_fma(z, invc, -1.0); /* exact */
    const double t1 = roger_pow_fma(kd, RPK(ln2hi, ROGER_POW_LN2HI), logc); /* exact: multiples of 2^-43 below 2^10 */
    const double t2 = roger_pow_add(t1, r);
    const double lo1 = roger_pow_fma(kd, RPK(ln2lo, ROGER_POW_LN2LO), logctail);
    const double lo2 = roger_pow_add(roger_pow_add(t1, -t2), r);
    const double ar = roger_pow_mul(-0.5, r);
    const double ar2 = roger_pow_mul(r, ar);
    const double ar3 = roger_pow_mul(r, ar2);
    const double hi = roger_pow_add(t2, ar2);
    const double lo3 = roger_pow_fma(ar, r, -ar2);
    const double lo4 = roger_pow_add(roger_pow_add(t2, -hi), ar2);
    /* p = log1p(r) - r + r^2/2 */
    const double q2 = roger_pow_fma(r, RPK(a6, ROGER_POW_A6), RPK(a5, ROGER_POW_A5));
    const double q1 = roger_pow_fma(ar2, q2, roger_pow_fma(r, RPK(a4, ROGER_POW_A4), RPK(a3, ROGER_POW_A3)));
    const double q0 = roger_pow_fma(ar2, q1, roger_pow_fma(r, RPK(a2, ROGER_POW_A2), RPK(a1, ROGER_POW_A1)));
    const double lo = roger_pow_fma(ar3, q0, roger_pow_add(roger_pow_add(roger_pow_add(lo1, lo2), lo3), lo4));
    const double y = roger_pow_add(hi, lo);
    *tail = roger_pow_add(roger_pow_add(hi, -y), lo);
    return y;
}

/* exp(x + xtail) for |x| below the overflow / flush thresholds */
ROGER_POW_FN double roger_pow_exp(double x, double xtail, const roger_pow_exp_entry *T)
{
    const double shift = RPK(shift, 0x1.8p52);
    double kd = roger_pow_fma(RPK(invln2n, ROGER_POW_INVLN2N), x, shift);
    const uint64_t ki = roger_pow_asu(kd);
    kd = roger_pow_add(kd, -shift);
    double r = roger_pow_fma(kd, RPK(negln2hin, ROGER_POW_NEGLN2HIN), x); /* exact product, one rounding of the sum */
    r = roger_pow_fma(kd, RPK(negln2lon, ROGER_POW_NEGLN2LON), r);
    r = roger_pow_add(r, xtail);
    const int idx = (int)(ki & (ROGER_POW_N - 1));
    const uint64_t top = ki << (52 - 7);
    const double tail = roger_pow_asd(T[idx].tail);
    const uint64_t sbits = T[idx].sbits + top;
    const double r2 = roger_pow_mul(r, r);
    const double pa = roger_pow_fma(r, RPK(c3, ROGER_POW_C3), RPK(c2, ROGER_POW_C2));
    const double pb = roger_pow_fma(r, RPK(c5, ROGER_POW_C5), RPK(c4, ROGER_POW_C4));
    double tmp = roger_pow_add(tail, r);
    tmp = roger_pow_fma(r2, pa, tmp);
    tmp = roger_pow_fma(roger_pow_mul(r2, r2), pb, tmp);
    const double scale = roger_pow_asd(sbits);
    return roger_pow_fma(scale, tmp, scale);
}

/* x^y for a positive NORMAL x given as bits and a finite y, without branches (range handling by selects): the form
 * the CUDA kernels inline several times per loop iteration so that independent evaluations interleave.
 * roger_pow() below is this function behind the special cases; both return the same bits on the common domain. */
/* second half of the evaluation: exp(y * (hi + lo)) with the range handling by selects */
ROGER_POW_FN double roger_pow_from_log(double hi, double lo, double y, const roger_pow_tables tabs);

ROGER_POW_FN double roger_pow_core(uint64_t ix, double y, const roger_pow_tables tabs)
{
    double lo;
    const double hi = roger_pow_log(ix, &lo, tabs.log_tab);
    return roger_pow_from_log(hi, lo, y, tabs);
}

ROGER_POW_FN double roger_pow_from_log(double hi, double lo, double y, const roger_pow_tables tabs)
{
    const double ehi = roger_pow_mul(y, hi);
    const double elo = roger_pow_fma(y, lo, roger_pow_fma(y, hi, -ehi));
    const int ovf = !(ehi < RPK(ovf, 0x1.62e42fefa39efp+9)); /* >= 1024 ln2: overflow              */
    const int unf = ehi < RPK(unf, -0x1.6232bdd7abcd2p+9); /* result below 2^-1022: flushed to 0 */
#if defined(ROGER_POW_CLEAN_INPUTS) && ROGER_POW_CLEAN_INPUTS
    /* (the form the SVAT step kernel keeps: same bits, and the register allocation its 128-register budget was tuned with) */
    const int out = ovf | unf;
    const double res = roger_pow_exp(out ? 0.0 : ehi, out ? 0.0 : elo, tabs.exp_tab);
#else
    /* out of range (or NaN) the reduction below runs on whatever ehi holds -- integer and floating-point operations only, the
     * table index is masked -- and its result is discarded by the selects: four selects fewer than cleaning the inputs first */
    const double res = roger_pow_exp(ehi, elo, tabs.exp_tab);
#endif
    return ovf ? roger_pow_asd(0x7ff0000000000000ULL) : (unf ? 0.0 : res);
}

/* true when x is a positive normal number (the domain of roger_pow_core) */
ROGER_POW_FN int roger_pow_is_pos_normal(uint64_t ix) { return ((ix >> 52) - 1ULL) < 0x7feULL; }

/* everything outside "positive normal base, finite exponent": zeros, subnormals, infinities, NaNs, negative bases */
#if defined(__CUDACC__)
static __host__ __device__ __noinline__
#else
static
#endif
double roger_pow_special(double x, double y, const roger_pow_tables tabs)
{
    uint64_t ix = roger_pow_asu(x);
    const uint64_t iy = roger_pow_asu(y);
    if ((iy << 1) == 0) return 1.0;               /* x^0 = 1 for every x, NaN included (C99 / NumPy)  */
    if (x == 1.0) return 1.0;                     /* 1^y = 1, NaN exponent included */
    if (x != x || y != y) return x + y;           /* NaN */
    if ((ix >> 63) != 0 && (ix << 1) != 0) return roger_pow_asd(0x7ff8000000000000ULL); /* negative base */
    if ((ix << 1) == 0) return (iy >> 63) ? roger_pow_asd(0x7ff0000000000000ULL) : 0.0;  /* 0^y */
    if ((ix << 1) == (0x7ff0000000000000ULL << 1)) return (iy >> 63) ? 0.0 : x;          /* inf^y */
    if ((iy << 1) == (0x7ff0000000000000ULL << 1))                                         /* x^(+-inf) */
        return ((x < 1.0) == ((iy >> 63) != 0)) ? roger_pow_asd(0x7ff0000000000000ULL) : 0.0;
    if ((ix >> 52) == 0) { /* subnormal base: normalise */
        ix = roger_pow_asu(roger_pow_mul(x, 0x1p52));
        ix -= 52ULL << 52;
    }
    return roger_pow_core(ix, y, tabs);
}

/* The common case -- positive normal base, finite exponent -- goes straight to the core (which returns exactly 1 for
 * x == 1 or y == 0); everything else takes the out-of-line special-case path.  Same bits as evaluating the special-case
 * chain first. */
ROGER_POW_FN double roger_pow(double x, double y, const roger_pow_tables tabs)
{
    const uint64_t ix = roger_pow_asu(x);
    const uint64_t iy = roger_pow_asu(y);
    if (roger_pow_is_pos_normal(ix) && ((iy >> 52) & 0x7ffULL) != 0x7ffULL) return roger_pow_core(ix, y, tabs);
    return roger_pow_special(x, y, tabs);
}

/* x^y1 and x^y2 with the logarithm of the common base evaluated once; each result has the bits roger_pow(x, y) returns */
ROGER_POW_FN void roger_pow2(double x, double y1, double y2, const roger_pow_tables tabs, double *r1, double *r2)
{
    const uint64_t ix = roger_pow_asu(x);
    const int plain = roger_pow_is_pos_normal(ix) && x != 1.0 && (y1 - y1 == 0.0) && (y2 - y2 == 0.0) && y1 != 0.0 && y2 != 0.0;
    if (!plain) {
        *r1 = roger_pow(x, y1, tabs);
        *r2 = roger_pow(x, y2, tabs);
        return;
    }
    double lo;
    const double hi = roger_pow_log(ix, &lo, tabs.log_tab);
    *r1 = roger_pow_from_log(hi, lo, y1, tabs);
    *r2 = roger_pow_from_log(hi, lo, y2, tabs);
}

/* exp(x) by the same reduction (tail 0): overflow -> +inf, results below 2^-1022 flushed to 0 */
ROGER_POW_FN double roger_exp(double x, const roger_pow_tables tabs)
{
    if (x != x) return x;
    if (!(x < 0x1.62e42fefa39efp+9)) return roger_pow_asd(0x7ff0000000000000ULL);
    if (x < -0x1.6232bdd7abcd2p+9) return 0.0;
    return roger_pow_exp(x, 0.0, tabs.exp_tab);
}

#endif /* ROGER_POW_H */
