// sas_device.cuh -- per-cell device code of the offline oxygen-18 StorAge-selection transport step
// (roger/core/transport.py:948-991 svat_transport_model_deterministic and the *_transport_iso_kernel of the
// process modules it calls).
//
// Execution model: ONE THREAD PER CELL, age axis walked sequentially, arrays stored AGE-MAJOR (one plane of
// C doubles per age bin, cell index fastest) so that the 32 threads of a warp -- 32 neighbouring cells at
// the same age bin -- read and write 256 contiguous bytes per access.
//
// Why sequential and not a warp-shuffle scan: the reference's cumulative sums are NumPy `cumsum`s, i.e.
// sequentially rounded, and the solver takes hard decisions on the results (`flux*tt > sa`, `SA > 0`,
// `sa <= 0`; transport.py:431, :494, :614, sas.py:232).  With the power-law exponents < 1 of the benchmark a
// last-bit difference in one cumulative sum flips such a decision and shows up as a 1e-6 relative difference
// within days ("SAS parity chain").  Walking the age axis in the reference's own order -- and
// summing with NumPy's pairwise tree, and calling the shared roger_pow() -- makes the kernel reproduce the
// CPU oracle bit for bit; a reassociating scan cannot.
//
// Ageing (transport.py:682-739) is a rotation of the plane ring (`base`), not a copy: bin j lives in plane
// (base + j) mod A, and the daily shift decrements base.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "roger_pow.h"
#include "roger_gamma.h"
#include "../../include/roger_b200_sas.h"

// Pointers that reach a __noinline__ function through CellCtx (a struct passed by reference) have lost their address space: the
// compiler then emits generic LD / ST with a descriptor it re-materialises through R2UR for every access (6 % of the day kernel's
// instructions in the ncu source view, and generic-load latency on the shared-memory pow tables).  These hints give it back.
#if defined(__CUDA_ARCH__) && !defined(ROGER_NO_HINTS)
#define ROGER_IS_GLOBAL(p) __builtin_assume(__isGlobal(p))
#define ROGER_IS_SHARED(p) __builtin_assume(__isShared(p))
#else
#define ROGER_IS_GLOBAL(p) ((void)0)
#define ROGER_IS_SHARED(p) ((void)0)
#endif

#ifndef ROGER_SAS_LOADGROUP
#define ROGER_SAS_LOADGROUP 2
#endif

namespace rsas {

constexpr int kMaxLeaves = 96;

// NumPy's pairwise summation of n contiguous doubles (np.add.reduce: blocks of <= 128 summed in 8 lanes, halves
// split at a multiple of 8) flattened into a left-to-right list of leaves; after leaf l, `merges[l]` pairs of
// partial sums on the stack are combined.
struct PwPlan {
    int n_leaves;
    short len[kMaxLeaves];
    unsigned char merges[kMaxLeaves];
};

struct SasParams {
    int A, substeps;
    int store_tt;
    int base;  // plane of age bin 0
    double vsmow, dmin, dmax;
    int64_t C;
};

struct Planes {  // age-major arrays
    double* sa[2];   // 0 rz, 1 ss
    double* msa[2];
    double* san;     // scratch: StorAge inside the sub-step loop
    double* TTn;     // scratch: accumulated cumulative travel-time distribution (index j holds TTn[j + 1])
    double* tt;      // optional: RSP_COUNT x A x C
    double* nmin[2]; // nitrate only: age-ranked mineral nitrogen of the root zone / subsoil (same ring as sa / msa)
    double* om;      // scratch of the split sub-step pass (ROGER_SAS_SPLIT): the SAS function value of every bin
    double* rev[2];  // SAS family 52 only: masked cumulative StorAge of the root zone / subsoil (bin-major), read backwards by omega()
    double* rk[6];   // RK4 solver only: the fourth stage array and the five accumulated travel-time distributions (bin-major like TTn)
};

// walks age bins 0..A-1 of one cell through the plane ring
struct Cursor {
    int64_t off;     // element offset of the current bin
    int64_t C;
    int64_t wrap_at; // offset at which the ring wraps
    int64_t span;    // A * C
    __device__ __forceinline__ Cursor(const SasParams& P, int64_t i) {
        C = P.C;
        span = (int64_t)P.A * P.C;
        off = (int64_t)P.base * P.C + i;
        wrap_at = span + i;
    }
    __device__ __forceinline__ void next() {
        off += C;
        if (off == wrap_at) off -= span;
    }
};

__device__ __forceinline__ bool isnan_d(double x) { return x != x; }

// running np.max: npx.max returns the first NaN it meets, else the maximum.  Which NaN is returned is unobservable
// (the result only enters comparisons and arithmetic), so a sticky flag replaces the early exit.
struct NpMax {
    double mx;
    bool nan;
    __device__ __forceinline__ void init(double a0) { mx = a0; nan = false; }
    __device__ __forceinline__ void push(double a) {
        nan = nan || isnan_d(a);
        mx = (a > mx) ? a : mx;
    }
    __device__ __forceinline__ double get() const { return nan ? __longlong_as_double(0x7ff8000000000000LL) : mx; }
};

// a / b for a divisor that is constant over a loop: rb = 1 / b (correctly rounded, computed once), then two
// fma-corrected quotient refinements.  With rb = RN(1/b) and a faithful quotient the last fma returns the correctly
// rounded a / b (Markstein's theorem) as long as nothing under- or overflows: div_range_ok() guards that, callers
// fall back to the division instruction otherwise.  tests/test_sas_gpu.py checks the equality on 1e8 random pairs.
__device__ __forceinline__ double div_by(double a, double b, double rb) {
    double q = __dmul_rn(a, rb);
    double r = __fma_rn(-b, q, a);
    q = __fma_rn(r, rb, q);
    r = __fma_rn(-b, q, a);
    return __fma_rn(r, rb, q);
}
__device__ __forceinline__ bool div_range_ok(double v) {  // |v| in [2^-500, 2^500]
    const unsigned e = (unsigned)((roger_pow_asu(v) >> 52) & 0x7ff);
    return (e - 523u) <= 1000u;
}

__device__ __forceinline__ double conc_to_delta(const SasParams& P, double conc) {  // transport.py:328-340
    const double d = 1000. * (conc / (P.vsmow * (1. - conc)) - 1.);
    return ((d < P.dmin) || (d > P.dmax)) ? __longlong_as_double(0x7ff8000000000000LL) : d;
}

// ---- SAS functions (roger/core/sas.py) --------------------------------------------------------------------
struct Family {
    double sp[8];      // parameter row, modified in place by the storage-dependent variants (the loop carry)
    double m;
    // per sub-step
    double S;          // max(SA) * m
    double uS, ulam, umask;
    bool is_uniform, is_dirac, is_kum, is_gamma, is_exp51, is_exp52, is_pow;
    bool fast;         // every inactive family provably contributes +0.0
    // id 52 (the exponential family read backwards along the age axis, sas.py:183): the masked cumulative StorAge of the layer,
    // written by the caller before the pass -- element k * rev_stride holds SA[k + 1]
    const double* rev;
    int64_t rev_stride;
};

__device__ __forceinline__ double s_rel(double S, const double* sp, double m) {
    double r = (S - sp[5]) / (sp[6] - sp[5]) * m;
    r = (r < 0) ? 0 : r;
    r = (r > 1) ? 1 : r;
    return r;
}

__device__ __forceinline__ void family_begin(Family& f, const double* sp_global, int64_t stride, double m) {
#pragma unroll
    for (int k = 0; k < 8; k++) f.sp[k] = sp_global[k * stride];
    f.m = m;
    f.rev = nullptr;
    f.rev_stride = 0;
}

// per sub-step preparation: everything that depends on max(SA) only
__device__ __forceinline__ void family_prepare(Family& f, double mx) {
    const double m = f.m;
    const double id = f.sp[0];
    f.is_uniform = (id == 1);
    f.umask = f.is_uniform ? 1.0 : 0.0;
    f.uS = mx * f.umask * m;              // sas.py:14-20
    f.ulam = 1 / f.uS * f.umask * m;
    f.is_dirac = (id == 2);
    f.S = mx * m;
    f.is_kum = (id == 3 || id == 31 || id == 32 || id == 33 || id == 34 || id == 35 || id == 36 || id == 37);
    {   // sas.py:51-108
        const double r = s_rel(f.S, f.sp, m);
        double* sp = f.sp;
        if (id == 31) { sp[1] = 1; sp[2] = sp[3] + (r * sp[4]); }
        if (id == 32) { sp[1] = sp[3] + ((1 - r) * sp[4]); sp[2] = 1; }
        if (id == 33) { sp[1] = 1; sp[2] = sp[3] + ((1 - r) * sp[4]); }
        if (id == 34) { sp[1] = sp[3] + (r * sp[4]); sp[2] = 1; }
        if (id == 35) { sp[1] = sp[3] + ((1 - r) * sp[4]); sp[2] = sp[3] + (r * sp[4]); }
        if (id == 36) { sp[1] = sp[3] + ((1 - r) * sp[4]); }
        if (id == 37) { sp[2] = sp[3] + (r * sp[4]); }
    }
    f.is_gamma = (id == 4);   // sp[3] carries Gamma(sp[1]) = exp(lgamma(sp[1])), put there by roger_sas_set_parameters
    f.is_exp51 = (id == 51);
    f.is_exp52 = (id == 52);
    f.is_pow = (id == 6 || id == 61 || id == 62);
    {   // sas.py:209-229
        const double r = s_rel(f.S, f.sp, m);
        if (id == 61) f.sp[1] = f.sp[3] + ((1 - r) * f.sp[4]);
        if (id == 62) f.sp[1] = f.sp[3] + (r * f.sp[4]);
    }
    // An inactive Kumaraswamy / power term is `value * 0`; it is +0.0 unless value is NaN or infinite, which needs a
    // negative or non-finite shape parameter -- or a non-finite S = max(SA) (inf / inf in the power law; a StorAge that a
    // diverged Euler / RK4 solution has driven to infinity poisons every family's sum there, and here).  With both shape
    // parameters finite and >= 0 and S finite (hence no SA above S, none NaN) the skipped terms are exactly +0.0 and skipping
    // them changes nothing.
    f.fast = (f.sp[1] >= 0) && (f.sp[2] >= 0) && (f.sp[1] <= 1.7e308) && (f.sp[2] <= 1.7e308) && (m == 1.0) && (f.S <= 1.7e308) &&
             (f.S >= -1.7e308);
}

// cumulative SAS distribution at age index jj (0..A) where the cumulative StorAge is SA: the sum of the five
// families in the reference's order (transport.py:399-428)
__device__ __forceinline__ double omega(const Family& f, const SasParams& P, const roger_pow_tables tabs, int jj, double SA) {
    const double m = f.m;
    double t = 0.0;
    if (f.is_uniform || !f.fast) {  // sas.py:5-35
        double o = ((SA < f.uS) ? ((SA > 0) ? f.ulam * SA : 0.) : 1.) * f.umask * m;
        if (jj == P.A) o = ((f.umask != 0) ? 1 : o) * m;
        o = ((f.uS <= 0) ? 0 : o) * m;
        t += o;
    }
    if (f.is_dirac || !f.fast) {  // sas.py:243-264
        double o = ((f.is_dirac && ((double)jj <= f.sp[1])) ? 0 : 1) * m;
        o = ((f.S <= 0) ? 0 : o) * (f.is_dirac ? 1.0 : 0.0) * m;
        t += o;
    }
    if (f.is_kum || !f.fast) {  // sas.py:38-134
        const double S = f.S;
        double v;
        const bool need_f = (S >= 0) ? ((SA > 0) && (SA < S)) : (SA > 0);
        if (need_f) v = 1 - roger_pow(1 - roger_pow(SA / S, f.sp[1], tabs), f.sp[2], tabs);
        else v = (S >= 0) ? ((SA > 0) ? 1. : 0.) : 0.;
        double o = v * (f.is_kum ? 1.0 : 0.0) * m;
        o = ((S <= 0) ? 0 : o) * m;
        t += o;
    }
    if (f.is_gamma) {  // sas.py:137-160; S carries the family mask there, so every other id selects the constant 0 (skipped: + 0.0)
        const double S = f.S;
        double v = 0.;
        if ((SA > 0) && (SA < S)) v = roger_gammainc(f.sp[1], f.sp[2] * SA / S, f.sp[3], tabs) / f.sp[3];
        double o = v * 1.0 * m;
        o = ((S <= 0) ? 0 : o) * m;
        t += o;
    }
    if (f.is_exp51 || f.is_exp52 || !f.fast) {  // sas.py:163-191
        const double S = f.S;
        double o = 0.0;
        if (f.is_exp51) o = (SA > 0) ? ((SA < S) ? 1 - roger_exp(f.sp[1] * (-1) * (SA / S), tabs) : 1.) : 0.;
        o = o * m;
        if (f.is_exp52) {  // [..., ::-1]: age index jj takes the value at nages - 1 - jj
            const int k = P.A - jj;
            const double SAr = (k == 0) ? 0.0 * m : f.rev[(int64_t)(k - 1) * f.rev_stride];
            o = (SAr > 0) ? ((SAr < S) ? 1 - roger_exp(f.sp[1] * (-1) * (SAr / S), tabs) : 1.) : 0.;
        }
        o = o * m;
        o = ((S <= 0) ? 0 : o) * m;
        t += o;
    }
    if (f.is_pow || !f.fast) {  // sas.py:194-240
        const double S = f.S;
        const bool in = (SA > 0) && (SA <= S);
        const double xr = SA / S;
        const uint64_t ixr = roger_pow_asu(xr);
        const double y = f.sp[1];
        // branch-free evaluation on the common domain (positive normal base, finite exponent) so that the evaluations of
        // neighbouring age bins interleave; anything else takes the general entry point (same bits where both apply)
        const bool easy = in && roger_pow_is_pos_normal(ixr) && (y - y == 0.0);
        double p = roger_pow_core(easy ? ixr : 0x3fe0000000000000ULL, easy ? y : 1.0, tabs);
        if (in && !easy) p = roger_pow(xr, y, tabs);
        double o = ((SA > 0) ? ((SA <= S) ? p : 1.) : 0.) * (f.is_pow ? 1.0 : 0.0) * m;
        o = ((S <= 0) ? 0 : o) * m;
        t += o;
    }
    return t;
}

// ---- streaming NumPy pairwise sums -------------------------------------------------------------------------
// Bins are visited once, in order.  load(j, L) reads the NL inputs of bin j (and may record addresses); work(j, L, v)
// computes the bin, performs its stores and fills the NS summands v.  Within a group of 4 bins all loads are issued
// before the first dependent instruction, so four bins' worth of memory latency overlap.
template <int NS, int NL, class Load, class Work>
__device__ __forceinline__ void stream_sums(const PwPlan* __restrict__ plan, Load load, Work work, double* out) {
    double st[NS][8];
    int sp = 0;
    int j = 0;
    const int nl = plan->n_leaves;
    for (int l = 0; l < nl; l++) {
        const int len = plan->len[l];
        double res[NS];
        if (len < 8) {
#pragma unroll
            for (int k = 0; k < NS; k++) res[k] = 0.;
            for (int i = 0; i < len; i++, j++) {
                double L[NL], v[NS];
                load(j, L);
                work(j, L, v);
#pragma unroll
                for (int k = 0; k < NS; k++) res[k] += v[k];
            }
        } else {
            double r[NS][8];
#pragma unroll
            for (int k = 0; k < NS; k++)
#pragma unroll
                for (int q = 0; q < 8; q++) r[k][q] = -0.0;  // x + (-0.0) == x: the first chunk initialises the lanes
            const int len8 = len - (len % 8);
            for (int i = 0; i < len8; i += 8, j += 8) {
#pragma unroll
                for (int g = 0; g < 8; g += ROGER_SAS_LOADGROUP) {  // this many bins' loads in flight at a time
                    double L[ROGER_SAS_LOADGROUP][NL];
#pragma unroll
                    for (int q = 0; q < ROGER_SAS_LOADGROUP; q++) load(j + g + q, L[q]);
#pragma unroll
                    for (int q = 0; q < ROGER_SAS_LOADGROUP; q++) {
                        double v[NS];
                        work(j + g + q, L[q], v);
#pragma unroll
                        for (int k = 0; k < NS; k++) r[k][g + q] += v[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NS; k++)
                res[k] = ((r[k][0] + r[k][1]) + (r[k][2] + r[k][3])) + ((r[k][4] + r[k][5]) + (r[k][6] + r[k][7]));
            for (int i = len8; i < len; i++, j++) {
                double L[NL], v[NS];
                load(j, L);
                work(j, L, v);
#pragma unroll
                for (int k = 0; k < NS; k++) res[k] += v[k];
            }
        }
#pragma unroll
        for (int k = 0; k < NS; k++) st[k][sp] = res[k];
        sp++;
        for (int c = plan->merges[l]; c > 0; c--) {
            sp--;
#pragma unroll
            for (int k = 0; k < NS; k++) st[k][sp - 1] = st[k][sp - 1] + st[k][sp];
        }
    }
#pragma unroll
    for (int k = 0; k < NS; k++) out[k] = 0.0 + st[k][0];
}

// The same sums with ONE copy of work() in the instruction stream (stream_sums<> unrolls it ten times, which is right for
// the light passes and thrashes the instruction cache for a heavy one).  The eight lanes stay in registers as a shift
// register: the lane a bin adds to is always lane 0, and the lanes rotate by one after every bin; a chunk is a multiple
// of 8 bins, so they are back in place when the tree sum is taken.  The next bin's loads are issued before this bin's work.
template <int NS, int NL, class Load, class Work>
__device__ __forceinline__ void stream_sums_rolled(const PwPlan* __restrict__ plan, int n_bins, Load load, Work work, double* out) {
    double st[NS][8];
    int sp = 0;
    int j = 0;
    const int nl = plan->n_leaves;
    double Lc[NL], Ln[NL];
    load(0, Lc);
    for (int l = 0; l < nl; l++) {
        const int len = plan->len[l];
        const int len8 = (len < 8) ? 0 : len - (len % 8);
        double res[NS], r[NS][8];
#pragma unroll
        for (int k = 0; k < NS; k++) {
            res[k] = 0.;
#pragma unroll
            for (int q = 0; q < 8; q++) r[k][q] = -0.0;
        }
        for (int i = 0; i < len; i++, j++) {
            if (i == len8 && len8 > 0) {
#pragma unroll
                for (int k = 0; k < NS; k++)
                    res[k] = ((r[k][0] + r[k][1]) + (r[k][2] + r[k][3])) + ((r[k][4] + r[k][5]) + (r[k][6] + r[k][7]));
            }
            if (j + 1 < n_bins) load(j + 1, Ln);
            double v[NS];
            work(j, Lc, v);
#pragma unroll
            for (int q = 0; q < NL; q++) Lc[q] = Ln[q];
            if (i < len8) {
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    const double t = r[k][0] + v[k];
#pragma unroll
                    for (int q = 0; q < 7; q++) r[k][q] = r[k][q + 1];
                    r[k][7] = t;
                }
            } else {
#pragma unroll
                for (int k = 0; k < NS; k++) res[k] += v[k];
            }
        }
        if (len8 == len && len8 > 0) {
#pragma unroll
            for (int k = 0; k < NS; k++)
                res[k] = ((r[k][0] + r[k][1]) + (r[k][2] + r[k][3])) + ((r[k][4] + r[k][5]) + (r[k][6] + r[k][7]));
        }
#pragma unroll
        for (int k = 0; k < NS; k++) st[k][sp] = res[k];
        sp++;
        for (int c = plan->merges[l]; c > 0; c--) {
            sp--;
#pragma unroll
            for (int k = 0; k < NS; k++) st[k][sp - 1] = st[k][sp - 1] + st[k][sp];
        }
    }
#pragma unroll
    for (int k = 0; k < NS; k++) out[k] = 0.0 + st[k][0];
}

// numpy.interp(p, cdf, 1..A) evaluated while the (non-decreasing) cdf streams by: the last index with cdf <= p and the
// value that follows it (calc_age_percentile, transport.py:9-56; numpy's bisection returns the same index for a
// non-decreasing array).
struct PctTracker {
    double p, x0, x1;
    int j;
    bool got1;
    __device__ __forceinline__ void init(double p_) { p = p_; j = -1; got1 = false; x0 = 0; x1 = 0; }
    __device__ __forceinline__ void push(int jbin, double v) {
        if (v <= p) { j = jbin; x0 = v; got1 = false; }
        else if (!got1) { x1 = v; got1 = true; }
    }
    __device__ __forceinline__ double result(int A, double cdf_max) const {
        double r;
        if (j == -1) r = 1.0;
        else if (j == A - 1) r = (double)A;
        else if (x0 == p) r = (double)(j + 1);
        else {
            const double slope = ((double)(j + 2) - (double)(j + 1)) / (x1 - x0);
            r = slope * (p - x0) + (double)(j + 1);
            if (isnan_d(r)) r = slope * (p - x1) + (double)(j + 2);
        }
        return (cdf_max <= 0) ? __longlong_as_double(0x7ff8000000000000LL) : r;
    }
};

struct FluxStats {   // diagnostics of one withdrawal (STATS kernels only)
    double wsum;     // sum(flux * tt), numerics.py:437-439
    double pct[5];   // 10 / 25 / 50 / 75 / 90 percentile ages of the travel-time distribution
    double avg;      // mean travel time
};

struct CellCtx {
    SasParams P;
    Planes pl;
    const PwPlan* plan;
    roger_pow_tables tabs;
    int64_t i;
    // element offset of age bin j of this cell in a rotated (state) array / in a scratch array
    __device__ __forceinline__ int64_t ring(int j) const {
        int p = P.base + j;
        if (p >= P.A) p -= P.A;
        return (int64_t)p * P.C + i;
    }
    __device__ __forceinline__ int64_t lin(int j) const { return (int64_t)j * P.C + i; }
    // address spaces of the pointers that are never null (see ROGER_IS_GLOBAL above); call at the top of a __noinline__ function
    __device__ __forceinline__ void assume_spaces() const {
        ROGER_IS_GLOBAL(pl.sa[0]); ROGER_IS_GLOBAL(pl.sa[1]); ROGER_IS_GLOBAL(pl.msa[0]); ROGER_IS_GLOBAL(pl.msa[1]);
        ROGER_IS_GLOBAL(pl.san); ROGER_IS_GLOBAL(pl.TTn); ROGER_IS_GLOBAL(plan);
        ROGER_IS_SHARED(tabs.log_tab); ROGER_IS_SHARED(tabs.exp_tab);
    }
};

#ifndef ROGER_SAS_UNROLL
#define ROGER_SAS_UNROLL 2
#endif
#ifndef ROGER_SAS_PF
#define ROGER_SAS_PF 0
#endif

// One sub-step of the deterministic solver for `n` consecutive bins starting at j0 (transport.py:388-470).  The n values
// of san / TTn were loaded by the caller (xb, tb); the cumulative SAS function of the n bins is evaluated first --
// independent evaluations the compiler interleaves -- then the sequentially dependent bookkeeping runs.
struct SubstepCarry {
    double acc_in, acc_new, acc_t, om_prev;
    NpMax mxn;
};
template <int U>
__device__ __forceinline__ void substep_block(const CellCtx& cx, const Family& fam, SubstepCarry& k, int j0, int n, const double* xb,
                                              const double* tb, bool first, double flux, double h, double fh, double m,
                                              double* __restrict__ san, double* __restrict__ TTn, double* __restrict__ rev_next = nullptr) {
    double om[U];
    double SAv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (u < n) {
            k.acc_in += xb[u];
            SAv[u] = first ? k.acc_in * m : k.acc_in;
        } else SAv[u] = 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; u++) om[u] = (u < n) ? omega(fam, cx.P, cx.tabs, j0 + u + 1, SAv[u]) : 0.0;
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (u < n) {
            const double x = xb[u];
            const double d = om[u] - k.om_prev;
            k.om_prev = om[u];
            double tti = (d >= 0) ? d : 0;
            const double q = flux * tti * h;
            const double ttqi = (q > x) ? x : q;
            tti = ttqi / fh;
            const double xn = x + -tti * flux * h;
            const int64_t o = cx.lin(j0 + u);
            san[o] = xn;
            k.acc_new += xn;
            k.mxn.push(k.acc_new);
            if (rev_next) rev_next[o] = k.acc_new;
            k.acc_t += tti;
            TTn[o] = (first ? 0.0 : tb[u]) + k.acc_t;
        }
    }
}

// The sub-step pass of the benchmark configuration -- power-law SAS function, cell inside the catchment (m == 1),
// S and flux * h in the range where div_by() is exact -- with everything that is constant over the age axis hoisted:
// same operations on the same values as substep_block<> / omega(), far fewer instructions.
struct PowPassCarry {
    double acc_in, acc_new, acc_t, om_prev, mxv;
    bool anynan;
};
struct PowPassConst {
    double S, rS, y, flux, h, fh, rfh;
};

// The two practically never-taken slow paths of the pass (a quotient outside the exact range of div_by()) are calls instead of inline
// IEEE division sequences / the general pow, which made a third of the loop body.  Same bits.  Measured on the same commands
// (gpurun_out/r3v_sas.log, r3w_bench_sas.json vs r3x_bench_sas.json): 88.1 -> 86.7 ms per day at 262 144 cells x 1000 bins (uniform
// cells), 651.5 -> 622.5 ms per day over 48 days of the BENCH workload (1024 x 1024 cells x 1500 bins).  ROGER_SAS_COLD_OUTLINE=0
// restores the inline form.
#ifndef ROGER_SAS_COLD_OUTLINE
#define ROGER_SAS_COLD_OUTLINE 1
#endif
__device__ __noinline__ double pow_pass_cold_pow(double SA, double S, double y, const roger_pow_tables tabs) {
    ROGER_IS_SHARED(tabs.log_tab); ROGER_IS_SHARED(tabs.exp_tab);
    return roger_pow(SA / S, y, tabs);
}
__device__ __noinline__ double pow_pass_cold_div(double a, double b) { return a / b; }

// N consecutive bins whose san / TTn values are already in registers (xb, tb)
template <int N, bool FIRST>
__device__ __forceinline__ void pow_pass_bins(const PowPassConst& k, PowPassCarry& c, const roger_pow_tables tabs, const double* xb,
                                              const double* tb, double* __restrict__ san_o, double* __restrict__ TTn_o, int64_t C) {
    double om[N];
#pragma unroll
    for (int u = 0; u < N; u++) {
        if (xb[u] == 0) {
            // an empty bin leaves the cumulative StorAge -- and therefore the SAS function value -- bit-identical to the
            // previous bin's: reuse it.  (Every dry day leaves an empty bin behind, and under spatially uniform forcing
            // whole warps skip together.)
            c.acc_in += xb[u];
            om[u] = (u == 0) ? c.om_prev : om[u - 1];
            continue;
        }
        c.acc_in += xb[u];
        const double SA = c.acc_in;
        const bool in = (SA > 0) && (SA <= k.S);
        // SA >= 2^-500 and S <= 2^500 (checked by the caller) keep the quotient normal and div_by() exact
        const bool easy = in && (SA >= 0x1p-500);
        const double xr = div_by(SA, k.S, k.rS);
        // outside `easy` the core runs on whatever bits xr has (arithmetic and a masked table index only) and its result is replaced
        double p = roger_pow_core(roger_pow_asu(xr), k.y, tabs);
#if ROGER_SAS_COLD_OUTLINE
        if (in && !easy) p = pow_pass_cold_pow(SA, k.S, k.y, tabs);
#else
        if (in && !easy) p = roger_pow(SA / k.S, k.y, tabs);
#endif
        om[u] = (SA > 0) ? ((SA <= k.S) ? p : 1.) : 0.;
    }
#pragma unroll
    for (int u = 0; u < N; u++) {
        const double x = xb[u];
        const double d = om[u] - c.om_prev;
        c.om_prev = om[u];
        double tti = (d >= 0) ? d : 0;
        const double q = k.flux * tti * k.h;
        const double ttqi = (q > x) ? x : q;
        // div_by() is exact for ttqi = 0 and for 2^-500 <= ttqi <= 2^500; ttqi <= flux * h * tti <= fh <= 2^500 (checked by the caller)
#if ROGER_SAS_COLD_OUTLINE
        if (ttqi == 0 || ttqi >= 0x1p-500) tti = div_by(ttqi, k.fh, k.rfh);
        else tti = pow_pass_cold_div(ttqi, k.fh);
#else
        tti = (ttqi == 0 || ttqi >= 0x1p-500) ? div_by(ttqi, k.fh, k.rfh) : ttqi / k.fh;
#endif
        const double xn = x + -tti * k.flux * k.h;
        san_o[u * C] = xn;
        c.acc_new += xn;
        c.mxv = (c.acc_new > c.mxv) ? c.acc_new : c.mxv;
        c.acc_t += tti;
        TTn_o[u * C] = (FIRST ? 0.0 : tb[u]) + c.acc_t;
    }
}

template <bool FIRST>
__device__ __noinline__ void pow_pass_tail(const PowPassConst& k, PowPassCarry& c, const roger_pow_tables tabs, double* __restrict__ san_o,
                                           double* __restrict__ TTn_o, int64_t C, int n) {
    ROGER_IS_GLOBAL(san_o); ROGER_IS_GLOBAL(TTn_o); ROGER_IS_SHARED(tabs.log_tab); ROGER_IS_SHARED(tabs.exp_tab);
    for (int u = 0; u < n; u++) {
        const double xb = san_o[u * C];
        const double tb = FIRST ? 0.0 : TTn_o[u * C];
        pow_pass_bins<1, FIRST>(k, c, tabs, &xb, &tb, san_o + u * C, TTn_o + u * C, C);
    }
}

// returns max(cumsum(san)) after the sub-step (np.max semantics)
template <int U, bool FIRST>
__device__ __forceinline__ double substep_pass_pow(const CellCtx& cx, double S, double y, double flux, double h, double fh,
                                                   double* __restrict__ san, double* __restrict__ TTn) {
    const int A = cx.P.A;
    const int64_t C = cx.P.C;
    PowPassConst k{S, 1.0 / S, y, flux, h, fh, 1.0 / fh};
    PowPassCarry c{0, 0, 0, 0.0, 0.0, false};  // om_prev = omega(SA[0] = 0) = 0, running max starts at SA[0] = 0
    const roger_pow_tables tabs = cx.tabs;
    ROGER_IS_SHARED(tabs.log_tab); ROGER_IS_SHARED(tabs.exp_tab);
    double* __restrict__ ps = san + cx.i;
    double* __restrict__ pt = TTn + cx.i;
    ROGER_IS_GLOBAL(ps); ROGER_IS_GLOBAL(pt);
    const int nfull = A / U;
    // register double buffer: the loads of block b + 1 are in flight while block b is computed.  The scratch arrays carry
    // U planes of padding, so the look-ahead of the last block needs no guard.
    double xn_[U], tn_[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        xn_[u] = ps[u * C];
        tn_[u] = FIRST ? 0.0 : pt[u * C];
    }
    for (int b = 0; b < nfull; b++) {
        double xb[U], tb[U];
#pragma unroll
        for (int u = 0; u < U; u++) { xb[u] = xn_[u]; tb[u] = tn_[u]; }
#pragma unroll
        for (int u = 0; u < U; u++) {
            xn_[u] = ps[(U + u) * C];
            tn_[u] = FIRST ? 0.0 : pt[(U + u) * C];
#if ROGER_SAS_PF  // experiment (measured 2.5 % SLOWER, profiles/r2_ncu_summary.md): L2 prefetch hint ROGER_SAS_PF bins beyond the register look-ahead
            static_assert(ROGER_SAS_PF + 2 * ROGER_SAS_UNROLL <= 8 + ROGER_SAS_UNROLL, "prefetch distance vs. scratch padding");
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + (size_t)(U + u + ROGER_SAS_PF) * C));
            if (!FIRST) asm volatile("prefetch.global.L2 [%0];" ::"l"(pt + (size_t)(U + u + ROGER_SAS_PF) * C));
#endif
        }
        pow_pass_bins<U, FIRST>(k, c, tabs, xb, tb, ps, pt, C);
        ps += U * C;
        pt += U * C;
    }
    if (nfull * U < A) pow_pass_tail<FIRST>(k, c, tabs, ps, pt, C, A - nfull * U);
    // np.max returns NaN when any cumulative sum is NaN; a NaN in the running sum stays (NaN + x = NaN), so its final value tells
    return isnan_d(c.acc_new) ? __longlong_as_double(0x7ff8000000000000LL) : c.mxv;
}

// The same pass split in two walks (ROGER_SAS_SPLIT): (1) cumulative StorAge -> SAS function value of every bin, stored to the
// scratch array `om`; (2) the sequential bookkeeping from the stored values.  Walk 1 carries one accumulator and no travel-time
// state, so four power evaluations fit into the register budget that holds two in the fused pass: the FP64 pipe sees twice the
// independent work per warp (the fused pass keeps it 40 % busy at 4 warps per scheduler).  Price: 16 B more traffic per bin
// and sub-step.  Same operations on the same values, hence the same bits.
#ifndef ROGER_SAS_SPLIT
#define ROGER_SAS_SPLIT 0
#endif
#ifndef ROGER_SAS_SPLIT_UNROLL
#define ROGER_SAS_SPLIT_UNROLL 4
#endif
template <bool FIRST>
__device__ __forceinline__ double substep_pass_pow_split(const CellCtx& cx, double S, double y, double flux, double h, double fh,
                                                         double* __restrict__ san, double* __restrict__ TTn, double* __restrict__ omv) {
    constexpr int U = ROGER_SAS_SPLIT_UNROLL;
    const int A = cx.P.A;
    const int64_t C = cx.P.C;
    const PowPassConst k{S, 1.0 / S, y, flux, h, fh, 1.0 / fh};
    const roger_pow_tables tabs = cx.tabs;
    {   // walk 1: omega
        const double* __restrict__ ps = san + cx.i;
        double* __restrict__ po = omv + cx.i;
        double acc_in = 0, om_prev = 0.0;
        const int nfull = A / U;
        double xn_[U];
#pragma unroll
        for (int u = 0; u < U; u++) xn_[u] = ps[u * C];
        for (int b = 0; b < nfull; b++) {
            double xb[U], SAv[U], om[U];
#pragma unroll
            for (int u = 0; u < U; u++) xb[u] = xn_[u];
#pragma unroll
            for (int u = 0; u < U; u++) xn_[u] = ps[(U + u) * C];  // look-ahead into the padding at the end
#pragma unroll
            for (int u = 0; u < U; u++) { acc_in += xb[u]; SAv[u] = acc_in; }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const double SA = SAv[u];
                const bool in = (SA > 0) && (SA <= k.S);
                const bool easy = in && (SA >= 0x1p-500);
                const double xr = div_by(SA, k.S, k.rS);
                double p = roger_pow_core(easy ? roger_pow_asu(xr) : 0x3fe0000000000000ULL, k.y, tabs);
                if (in && !easy) p = roger_pow(SA / k.S, k.y, tabs);
                om[u] = (SA > 0) ? ((SA <= k.S) ? p : 1.) : 0.;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {  // an empty bin repeats its predecessor's value bit for bit (same SA)
                if (xb[u] == 0) om[u] = (u == 0) ? om_prev : om[u - 1];
                po[u * C] = om[u];
            }
            om_prev = om[U - 1];
            ps += U * C;
            po += U * C;
        }
        for (int j = nfull * U; j < A; j++) {
            const double x = ps[0];
            acc_in += x;
            const double SA = acc_in;
            double o;
            if (x == 0) o = om_prev;
            else {
                const bool in = (SA > 0) && (SA <= k.S);
                const bool easy = in && (SA >= 0x1p-500);
                const double xr = div_by(SA, k.S, k.rS);
                double p = roger_pow_core(easy ? roger_pow_asu(xr) : 0x3fe0000000000000ULL, k.y, tabs);
                if (in && !easy) p = roger_pow(SA / k.S, k.y, tabs);
                o = (SA > 0) ? ((SA <= k.S) ? p : 1.) : 0.;
            }
            po[0] = o;
            om_prev = o;
            ps += C;
            po += C;
        }
    }
    // walk 2: travel-time bookkeeping
    double* __restrict__ ps = san + cx.i;
    double* __restrict__ pt = TTn + cx.i;
    const double* __restrict__ po = omv + cx.i;
    PowPassCarry c{0, 0, 0, 0.0, 0.0, false};
    constexpr int V = 4;
    int j = 0;
    for (; j + V <= A; j += V) {
        double xb[V], tb[V], om[V];
#pragma unroll
        for (int u = 0; u < V; u++) { xb[u] = ps[u * C]; tb[u] = FIRST ? 0.0 : pt[u * C]; om[u] = po[u * C]; }
#pragma unroll
        for (int u = 0; u < V; u++) {
            const double x = xb[u];
            const double d = om[u] - c.om_prev;
            c.om_prev = om[u];
            double tti = (d >= 0) ? d : 0;
            const double q = k.flux * tti * k.h;
            const double ttqi = (q > x) ? x : q;
            tti = (ttqi == 0 || div_range_ok(ttqi)) ? div_by(ttqi, k.fh, k.rfh) : ttqi / k.fh;
            const double xn = x + -tti * k.flux * k.h;
            ps[u * C] = xn;
            c.acc_new += xn;
            c.anynan = c.anynan || isnan_d(c.acc_new);
            c.mxv = (c.acc_new > c.mxv) ? c.acc_new : c.mxv;
            c.acc_t += tti;
            pt[u * C] = tb[u] + c.acc_t;
        }
        ps += V * C; pt += V * C; po += V * C;
    }
    for (; j < A; j++) {
        const double x = ps[0], tb = FIRST ? 0.0 : pt[0], o = po[0];
        const double d = o - c.om_prev;
        c.om_prev = o;
        double tti = (d >= 0) ? d : 0;
        const double q = k.flux * tti * k.h;
        const double ttqi = (q > x) ? x : q;
        tti = (ttqi == 0 || div_range_ok(ttqi)) ? div_by(ttqi, k.fh, k.rfh) : ttqi / k.fh;
        const double xn = x + -tti * k.flux * k.h;
        ps[0] = xn;
        c.acc_new += xn;
        c.anynan = c.anynan || isnan_d(c.acc_new);
        c.mxv = (c.acc_new > c.mxv) ? c.acc_new : c.mxv;
        c.acc_t += tti;
        pt[0] = tb + c.acc_t;
        ps += C; pt += C; po += C;
    }
    return c.anynan ? __longlong_as_double(0x7ff8000000000000LL) : c.mxv;
}

// ---- one SAS-controlled withdrawal from layer `src`, optionally received by layer `dst` (dst < 0: leaves the
// soil).  evapotranspiration.py:653-716, :831-895; subsurface_runoff.py:1531-1625, :1753-1815;
// capillary_rise.py:404-500.
// The sub-step solver of one flux (transport.py:362-470): leaves the accumulated cumulative travel-time distribution of
// the layer `src` in the scratch array TTn (index j holds TTn[j + 1]).  Returns false when flux * h <= 0: every sub-step
// then leaves StorAge and TTn untouched and TTn counts as zero.
__device__ __noinline__ bool solve_substeps(const CellCtx& cx, int src, double flux, const double* sp_global, int64_t sp_stride, double m) {
    cx.assume_spaces();
    constexpr int U = ROGER_SAS_UNROLL;
    const SasParams& P = cx.P;
    const int A = P.A;
    double* __restrict__ sa = cx.pl.sa[src];
    double* __restrict__ san = cx.pl.san;
    double* __restrict__ TTn = cx.pl.TTn;
    ROGER_IS_GLOBAL(sa); ROGER_IS_GLOBAL(san); ROGER_IS_GLOBAL(TTn); ROGER_IS_GLOBAL(sp_global); ROGER_IS_GLOBAL(cx.plan);
    ROGER_IS_SHARED(cx.tabs.log_tab); ROGER_IS_SHARED(cx.tabs.exp_tab);
    const double h = 1.0 / P.substeps;
    if (!(flux * h > 0)) return false;
    // cumulative StorAge of the layer (transport.py:343-359) -> S = max(SA); the layer's StorAge is copied into the
    // (unrotated) scratch array the sub-steps work on
    NpMax mx;
    const double SA0 = 0. * m;
    mx.init(SA0);
    // family 52 reads the cumulative StorAge backwards: two scratch planes take it in turns (this sub-step's, the next one's)
    const bool is52 = (sp_global[0] == 52);
    double* const rv[2] = {cx.pl.rev[0], cx.pl.rev[1]};
    {
        double acc = 0;
        int j = 0;
        for (; j + 8 <= A; j += 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = sa[cx.ring(j + q)];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                san[cx.lin(j + q)] = v[q];
                acc += v[q];
                mx.push(acc * m);
                if (is52) rv[0][cx.lin(j + q)] = acc * m;
            }
        }
        for (; j < A; j++) {
            const double v = sa[cx.ring(j)];
            san[cx.lin(j)] = v;
            acc += v;
            mx.push(acc * m);
            if (is52) rv[0][cx.lin(j)] = acc * m;
        }
    }
    Family fam;
    family_begin(fam, sp_global, sp_stride, m);
    const double fh = flux * h;
    for (int s = 0; s < P.substeps; s++) {  // transport.py:388-470
        double* rev_next = nullptr;
        if (is52) {
            fam.rev = rv[s & 1] + cx.i;
            fam.rev_stride = P.C;
            rev_next = rv[(s + 1) & 1];
        }
        family_prepare(fam, mx.get());
        const bool first = (s == 0);
        if (fam.fast && fam.is_pow && fam.S > 0 && div_range_ok(fam.S) && div_range_ok(fh)) {
#if ROGER_SAS_SPLIT
            const double S_next = first ? substep_pass_pow_split<true>(cx, fam.S, fam.sp[1], flux, h, fh, san, TTn, cx.pl.om)
                                        : substep_pass_pow_split<false>(cx, fam.S, fam.sp[1], flux, h, fh, san, TTn, cx.pl.om);
#else
            const double S_next = first ? substep_pass_pow<U, true>(cx, fam.S, fam.sp[1], flux, h, fh, san, TTn)
                                        : substep_pass_pow<U, false>(cx, fam.S, fam.sp[1], flux, h, fh, san, TTn);
#endif
            mx.init(S_next);
            continue;
        }
        SubstepCarry k;
        k.om_prev = omega(fam, P, cx.tabs, 0, SA0);
        k.acc_in = 0; k.acc_new = 0; k.acc_t = 0;
        k.mxn.init(SA0);
        // register double buffer: the loads of block b+1 are in flight while block b is computed
        double xn_[U], tn_[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            xn_[u] = (u < A) ? san[cx.lin(u)] : 0.0;
            tn_[u] = (!first && u < A) ? TTn[cx.lin(u)] : 0.0;
        }
        int j0 = 0;
        for (; j0 + U <= A; j0 += U) {
            double xb[U], tb[U];
#pragma unroll
            for (int u = 0; u < U; u++) { xb[u] = xn_[u]; tb[u] = tn_[u]; }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int jn = j0 + U + u;
                xn_[u] = (jn < A) ? san[cx.lin(jn)] : 0.0;
                tn_[u] = (!first && jn < A) ? TTn[cx.lin(jn)] : 0.0;
            }
            substep_block<U>(cx, fam, k, j0, U, xb, tb, first, flux, h, fh, m, san, TTn, rev_next);
        }
        if (j0 < A) substep_block<U>(cx, fam, k, j0, A - j0, xn_, tn_, first, flux, h, fh, m, san, TTn, rev_next);
        mx = k.mxn;
    }
    return true;
}

template <bool STATS>
__device__ __noinline__ void withdrawal(const CellCtx& cx, int src, int dst, double flux, const double* sp_global, int64_t sp_stride,
                                        double m, double* C_flux, double* C_iso_flux, int process, FluxStats* fs) {
    cx.assume_spaces();
    const SasParams& P = cx.P;
    const int A = P.A;
    double* __restrict__ sa = cx.pl.sa[src];
    double* __restrict__ msa = cx.pl.msa[src];
    double* __restrict__ TTn = cx.pl.TTn;
    ROGER_IS_GLOBAL(sa); ROGER_IS_GLOBAL(msa); ROGER_IS_GLOBAL(TTn); ROGER_IS_GLOBAL(cx.plan);
    const int64_t C = P.C;
    const bool active = solve_substeps(cx, src, flux, sp_global, sp_stride, m);

    // aggregate (transport.py:472-496), flux concentration (:512-535), isotope mixing into the receiving layer,
    // StorAge update (:599-619)
    double* __restrict__ sa_d = (dst >= 0) ? cx.pl.sa[dst] : sa;  // (never dereferenced when dst < 0)
    double* __restrict__ msa_d = (dst >= 0) ? cx.pl.msa[dst] : msa;
    double* __restrict__ tt_out = (P.store_tt && cx.pl.tt) ? cx.pl.tt + (size_t)process * (size_t)A * (size_t)C : nullptr;
    ROGER_IS_GLOBAL(sa_d); ROGER_IS_GLOBAL(msa_d);
    const double nsub = (double)P.substeps;
    double prev = 0. / nsub;
    constexpr int NS = STATS ? 4 : 2;
    double sums[NS];
    PctTracker trk[STATS ? 5 : 1];
    double cdf = 0;
    NpMax cdf_max;
    if (STATS) {
        const double pcts[5] = {0.1, 0.25, 0.5, 0.75, 0.9};
#pragma unroll
        for (int k = 0; k < 5; k++) trk[k].init(pcts[k]);
    }
    bool first_bin = true;
    stream_sums<NS, 5>(cx.plan,
        [&](int j, double* L) {
            const int64_t o = cx.ring(j);
            L[0] = active ? TTn[cx.lin(j)] : 0.;
            L[1] = sa[o];
            L[2] = msa[o];
            if (dst >= 0) { L[3] = sa_d[o]; L[4] = msa_d[o]; }
        },
        [&](int j, const double* L, double* v) {
            const int64_t o = cx.ring(j);
            const double cur = L[0] / nsub;
            double t = cur - prev;
            prev = cur;
            const double s_j = L[1];
            const double q = flux * t;
            const double ttq = (q > s_j) ? s_j : q;
            t = (flux > 0) ? ttq / flux : 0;
            t = t * m;
            const double ms_j = L[2];
            const double mtt = ((t > 0) ? ms_j : 0) * m;
            v[0] = t;
            v[1] = mtt * t;
            if (STATS) {
                v[2] = flux * t;
                v[3] = (double)(j + 1) * t;
                cdf += t;  // TT[1:] = cumsum(tt)
                if (first_bin) { cdf_max.init(cdf); first_bin = false; } else cdf_max.push(cdf);
#pragma unroll
                for (int k = 0; k < 5; k++) trk[k].push(j, cdf);
            }
            if (dst >= 0) {
                const double in = t * flux;
                const double sd = L[3];
                const double md = L[4];
                const double tot = in + sd;
                const double mixed = (tot > 0) ? md * (sd / tot) + mtt * (in / tot) : md;
                msa_d[o] = mixed * m;
                sa_d[o] = sd + t * flux * m;
            }
            double u = s_j + -flux * t;
            u = ((u > -1e-5) && (u < 0)) ? 0 : u;
            u = u * m;
            sa[o] = u;
            msa[o] = ((u <= 0) ? 0 : ms_j) * m;
            if (tt_out) tt_out[cx.lin(j)] = t;
        }, sums);
    if (STATS && fs) {
        fs->wsum = sums[2];
#pragma unroll
        for (int k = 0; k < 5; k++) fs->pct[k] = trk[k].result(A, cdf_max.get());
        fs->avg = (sums[0] > 0) ? sums[3] : __longlong_as_double(0x7ff8000000000000LL);
    }
    double conc = (sums[0] > 0) ? sums[1] / sums[0] : __longlong_as_double(0x7ff8000000000000LL);
    conc = (conc != 0) ? conc : __longlong_as_double(0x7ff8000000000000LL);
    *C_flux = conc * m;
    *C_iso_flux = conc_to_delta(P, *C_flux) * m;
}

// ================================================================================================================
// Conservative anions (bromide / chloride): msa holds the solute MASS of every age bin
// (infiltration.py:2350-2422, :2504-2560; evapotranspiration.py:620-651, :905-982; the anion kernels of
// subsurface_runoff.py and capillary_rise.py; transport.py:582-596 calc_mtt, :622-679 ageing).
// ================================================================================================================
// One flux out of layer `src`: water by the SAS solver, solute as mtt = clip(msa / sa * alpha * tt * flux, 0, msa); the
// receiving layer (dst >= 0) takes both.  with_solute = false is the water-only soil evaporation.  Returns sum(mtt).
template <bool STATS>
__device__ __noinline__ double anion_withdrawal(const CellCtx& cx, int src, int dst, double flux, const double* sp_global, int64_t sp_stride,
                                                double m, double alpha, bool with_solute, int process, FluxStats* fs) {
    cx.assume_spaces();
    const SasParams& P = cx.P;
    const int A = P.A;
    double* __restrict__ sa = cx.pl.sa[src];
    double* __restrict__ msa = cx.pl.msa[src];
    double* __restrict__ TTn = cx.pl.TTn;
    const int64_t C = P.C;
    const bool active = solve_substeps(cx, src, flux, sp_global, sp_stride, m);
    double* __restrict__ sa_d = (dst >= 0) ? cx.pl.sa[dst] : nullptr;
    double* __restrict__ msa_d = (dst >= 0) ? cx.pl.msa[dst] : nullptr;
    double* __restrict__ tt_out = (P.store_tt && cx.pl.tt && process >= 0) ? cx.pl.tt + (size_t)process * (size_t)A * (size_t)C : nullptr;
    const double nsub = (double)P.substeps;
    double prev = 0. / nsub;
    constexpr int NS = STATS ? 4 : 1;
    double sums[NS];
    PctTracker trk[STATS ? 5 : 1];
    double cdf = 0;
    NpMax cdf_max;
    if (STATS) {
        const double pcts[5] = {0.1, 0.25, 0.5, 0.75, 0.9};
#pragma unroll
        for (int k = 0; k < 5; k++) trk[k].init(pcts[k]);
    }
    bool first_bin = true;
    stream_sums<NS, 5>(cx.plan,
        [&](int j, double* L) {
            const int64_t o = cx.ring(j);
            L[0] = active ? TTn[cx.lin(j)] : 0.;
            L[1] = sa[o];
            if (with_solute) L[2] = msa[o];
            if (dst >= 0) { L[3] = sa_d[o]; L[4] = msa_d[o]; }
        },
        [&](int j, const double* L, double* v) {
            const int64_t o = cx.ring(j);
            const double cur = L[0] / nsub;
            double t = cur - prev;
            prev = cur;
            const double s_j = L[1];
            const double q = flux * t;
            const double ttq = (q > s_j) ? s_j : q;
            t = (flux > 0) ? ttq / flux : 0;
            t = t * m;
            double mtt = 0.0;
            if (with_solute) {
                const double ms_j = L[2];
                double w = ((s_j > 0) ? (ms_j / s_j) : 0) * alpha * t * flux;
                w = (w <= 0) ? 0 : w;
                w = (w > ms_j) ? ms_j : w;
                mtt = w * m;
                msa[o] = ms_j + -mtt * m;
            }
            v[0] = mtt;
            if (STATS) {
                v[1] = t;
                v[2] = flux * t;
                v[3] = (double)(j + 1) * t;
                cdf += t;
                if (first_bin) { cdf_max.init(cdf); first_bin = false; } else cdf_max.push(cdf);
#pragma unroll
                for (int k = 0; k < 5; k++) trk[k].push(j, cdf);
            }
            double u = s_j + -flux * t;
            u = ((u > -1e-5) && (u < 0)) ? 0 : u;
            sa[o] = u * m;
            if (dst >= 0) {
                msa_d[o] = L[4] + mtt * m;
                sa_d[o] = L[3] + t * flux * m;
            }
            if (tt_out) tt_out[cx.lin(j)] = t;
        }, sums);
    if (STATS && fs) {
        fs->wsum = sums[2];
#pragma unroll
        for (int k = 0; k < 5; k++) fs->pct[k] = trk[k].result(A, cdf_max.get());
        fs->avg = (sums[1] > 0) ? sums[3] : __longlong_as_double(0x7ff8000000000000LL);
    }
    return sums[0];
}

__device__ __noinline__ void storage_age_statistics(const CellCtx& cx, int which, double m, int n_pct, const double* pcts, double* pct_out,
                                                    double* avg_out);

struct AnionParams {  // per-cell planes (may be null: alpha = 1, no crops)
    const double* alpha_transp;
    const double* alpha_q;
    const double* lu_id;
    const double* S_sat_rz;
    int tracer = 0;  // settings.solute == 3, the virtual tracer: the soil evaporation carries the solute (evapotranspiration.py:723-791)
};

// ================================================================================================================
// Nitrate (settings.solute = 2): the soil nitrogen cycle on the age-ranked arrays, roger/core/nitrate.py:500-696
// calc_nitrogen_cycle_kernel with its process kernels (:68-173 denitrification, :176-257 nitrification, :260-302
// mineralisation, :305-346 fixation, :349-427 gaseous loss, :430-467 ammonium uptake), called between the capillary rise and
// the storage kernels of the anion day (transport.py:966-969).  Sums over the age axis are NumPy pairwise sums
// (stream_sums), the reversed cumulative sum of the denitrification is walked from the oldest bin down.  `msa_rz += ma_rz`
// is applied in the nitrification pass already: nothing reads msa_rz between the two statements of the reference.
// ================================================================================================================
struct CropParams {  // crop variant of the offline model: redistribution after root growth / root loss (null pointers: no crop model)
    const double* sp_rg;  // SAS parameter rows, [8][C]
    const double* sp_rl;
    const double* re;     // [2][C]: today's re_rg, re_rl
    double* out;          // [RCO_COUNT][C]
};

struct NitrateParams {
    const double* params;   // RNP_COUNT planes
    const double* forcing;  // RNF_COUNT planes of the day
    double* out;            // RNO_COUNT planes
    double dx, dy, zmax, dt;
};

__device__ __forceinline__ double temp_coeff(double T, double lo, double m) {
    double c = (((T >= lo) && (T <= 30)) ? T / (30 - lo) : 0) * m;
    c = ((T > 30) ? 1 : c) * m;
    return c;
}
// r = (wsum > 0) ? (w / wsum) * pot : 0, limited to nmin and to >= 0, each statement "* maskCatch"
__device__ __forceinline__ double distribute(double w, double wsum, double pot, double nmin, double m) {
    double v = ((wsum > 0) ? ((w / wsum) * pot) : 0.0) * m;
    v = ((v > nmin) ? nmin : v) * m;
    v = ((v < 0) ? 0 : v) * m;
    return v;
}
// denitrification of one layer (nitrate.py:68-173): the oldest solute is removed first; mr is left in `scratch` (unrotated)
__device__ __noinline__ void denit_layer(const CellCtx& cx, int layer, double msum, double Ssum, double km, double Dmax, double S_sat, double T,
                                         double m, const NitrateParams& np_, double* __restrict__ scratch) {
    cx.assume_spaces();
    const int A = cx.P.A;
    double* __restrict__ msa = cx.pl.msa[layer];
    const double S = Ssum * m;
    const double coeff = temp_coeff(T, 5, m);
    const double ms = msum * m;
    const double dt = np_.dt, dx = np_.dx, dy = np_.dy;
    double mr_pot = (Dmax * (dt / (365 * 24)) * dx * dy * 100 * (ms / (km * (dt / (365 * 24)) * dx * dy * 100 + ms))) * coeff * m;
    mr_pot = ((S >= 0.7 * S_sat) ? mr_pot : 0) * m;
    double acc = 0;
    for (int j = A - 1; j >= 0; j--) {
        const int64_t o = cx.ring(j);
        const double v = msa[o];
        acc = (j == A - 1) ? v : acc + v;
        double t = ((acc * m < mr_pot) ? 0 : v) * m;
        if (j == A - 1) t += -((t >= mr_pot) ? mr_pot : 0) * m;
        double mra = (v * m - t) * m;
        mra = ((mra < 0) ? 0 : mra) * m;
        const double mr = mra * m;
        msa[o] = v + -mr;
        scratch[cx.lin(j)] = mr;
    }
}

__device__ __noinline__ void nitrogen_cycle(const CellCtx& cx, const NitrateParams& np_, double transp, double m) {
    cx.assume_spaces();
    const SasParams& P = cx.P;
    const int64_t C = P.C, i = cx.i;
    double* __restrict__ s_rz = cx.pl.sa[0];
    double* __restrict__ s_ss = cx.pl.sa[1];
    double* __restrict__ m_rz = cx.pl.msa[0];
    double* __restrict__ m_ss = cx.pl.msa[1];
    double* __restrict__ n_rz = cx.pl.nmin[0];
    double* __restrict__ n_ss = cx.pl.nmin[1];
#define NP(k) np_.params[(int64_t)(k) * C + i]
#define NF(k) np_.forcing[(int64_t)(k) * C + i]
#define NO(k) np_.out[(int64_t)(k) * C + i]
    const double T = NF(RNF_temp_soil), z_root = NF(RNF_z_root), z_soil = NP(RNP_z_soil), lu = NF(RNF_lu_id);
    const double dt = np_.dt, dx = np_.dx, dy = np_.dy;
    double nfix = NP(RNP_kfix_rz) * (dt / (365 * 24)) * dx * dy * 100 * temp_coeff(T, 0, m) * (z_root / (np_.zmax * z_soil)) * m;
    {
        const bool fixer = lu == 541 || lu == 577 || lu == 578 || lu == 580 || lu == 581 || lu == 583 || lu == 584 || lu == 586 ||
                           lu == 587 || lu == 588;
        nfix = fixer ? nfix : 0;
    }
    nfix = nfix * m;
    NO(RNO_nfix_s) = nfix;
    const double min_rz = (NP(RNP_kmin_rz) * (dt / (365 * 24)) * dx * dy * 100 * temp_coeff(T, 0, m) * m) * m;
    const double min_ss = (NP(RNP_kmin_ss) * (dt / (365 * 24)) * dx * dy * 100 * temp_coeff(T, 0, m) * m) * m;
    NO(RNO_min_s) = min_rz + min_ss;
    {
        const int64_t o0 = cx.ring(0);
        double v = n_rz[o0];
        v += nfix;
        v += min_rz * m;
        n_rz[o0] = v;
        n_ss[o0] = n_ss[o0] + min_ss * m;
    }
    // sums the processes start from
    double s0[2];
    stream_sums<2, 2>(cx.plan,
        [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = s_rz[o]; L[1] = n_rz[o]; },
        [&](int, const double* L, double* v) { v[0] = L[0]; v[1] = L[1]; }, s0);
    const double Ssum = s0[0], S = Ssum * m;
    const double S_sat_rz = NP(RNP_S_sat_rz);
    // nitrification of the root zone (+ msa_rz += ma_rz)
    double sB[3];
    {
        const double nsum = s0[1];
        double pot = (NP(RNP_dmax_nit_rz) * (dt / (365 * 24)) * dx * dy * 100 *
                      (nsum / (NP(RNP_km_nit_rz) * (dt / (365 * 24)) * dx * dy * 100 + nsum))) * temp_coeff(T, 1, m) * m;
        pot = ((S < 0.9 * S_sat_rz) ? pot : 0) * m;
        stream_sums<3, 3>(cx.plan,
            [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = s_rz[o]; L[1] = n_rz[o]; L[2] = m_rz[o]; },
            [&](int j, const double* L, double* v) {
                const int64_t o = cx.ring(j);
                const double ma = distribute(L[0], Ssum, pot, L[1], m) * m;
                const double nn = L[1] + -ma;
                const double mm = L[2] + ma;
                n_rz[o] = nn;
                m_rz[o] = mm;
                v[0] = ma + 0.0;  // ma_s = ma_rz + ma_ss, and vs.ma_ss is never written by the reference (0)
                v[1] = nn;
                v[2] = mm;
            }, sB);
    }
    NO(RNO_nit_s) = sB[0];
    // gaseous loss of ammonium
    double sC[2];
    {
        double pot = (NP(RNP_kngl_rz) * (dt / (365 * 24)) * dx * dy * 100) * temp_coeff(T, 5, m) * m;
        pot = ((S < 0.9 * S_sat_rz) ? pot : 0) * m;
        const double nsum = sB[1];
        stream_sums<2, 1>(cx.plan,
            [&](int j, double* L) { L[0] = n_rz[cx.ring(j)]; },
            [&](int j, const double* L, double* v) {
                const double r = distribute(L[0], nsum, pot, L[0], m);
                const double nn = L[0] + -r * m;
                n_rz[cx.ring(j)] = nn;
                v[0] = r;
                v[1] = nn;
            }, sC);
    }
    NO(RNO_ngas_s) = sC[0];
    // ammonium uptake by plants
    double sD[2];
    {
        const double pot = ((transp > 0) ? NF(RNF_nup) * (z_root / (z_soil * np_.zmax)) : 0) * m;
        const double nsum = sC[1];
        stream_sums<2, 1>(cx.plan,
            [&](int j, double* L) { L[0] = n_rz[cx.ring(j)]; },
            [&](int j, const double* L, double* v) {
                const double r = distribute(L[0], nsum, pot, L[0], m);
                const double nn = L[0] + -r * m;
                n_rz[cx.ring(j)] = nn;
                v[0] = r;
                v[1] = nn;
            }, sD);
    }
    NO(RNO_nh4_up) = sD[0] * m;
    // denitrification of both layers; the subsoil's sums first
    double sF[3];
    stream_sums<3, 3>(cx.plan,
        [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = s_ss[o]; L[1] = m_ss[o]; L[2] = n_ss[o]; },
        [&](int, const double* L, double* v) { v[0] = L[0]; v[1] = L[1]; v[2] = L[2]; }, sF);
    denit_layer(cx, 0, sB[2], Ssum, NP(RNP_km_denit_rz), NP(RNP_dmax_denit_rz), S_sat_rz, T, m, np_, cx.pl.san);
    denit_layer(cx, 1, sF[1], sF[0], NP(RNP_km_denit_ss), NP(RNP_dmax_denit_ss), NP(RNP_S_sat_ss), T, m, np_, cx.pl.TTn);
    double sH[1];
    stream_sums<1, 2>(cx.plan,
        [&](int j, double* L) { const int64_t o = cx.lin(j); L[0] = cx.pl.san[o]; L[1] = cx.pl.TTn[o]; },
        [&](int, const double* L, double* v) { v[0] = L[0] + L[1]; }, sH);
    NO(RNO_denit_s) = sH[0];
    NO(RNO_Nmin_s) = sD[1] + sF[2] * m;
#undef NP
#undef NF
#undef NO
}

template <bool STATS, bool NITRATE = false>
__device__ __forceinline__ void anion_day(const CellCtx& cx, const double* __restrict__ forcing, double m, const double* __restrict__ sas_params,
                                          const AnionParams ap, double* __restrict__ out, double* __restrict__ stats, double* __restrict__ C_s_prev,
                                          const NitrateParams* nitrate = nullptr, const CropParams crop = CropParams{nullptr, nullptr, nullptr, nullptr}) {
    const SasParams& P = cx.P;
    const int A = P.A;
    const int64_t C = P.C, i = cx.i;
    double* __restrict__ s_rz = cx.pl.sa[0];
    double* __restrict__ s_ss = cx.pl.sa[1];
    double* __restrict__ m_rz = cx.pl.msa[0];
    double* __restrict__ m_ss = cx.pl.msa[1];
#define F(k) forcing[(int64_t)(k) * C + i]
#define O(k) out[(int64_t)(k) * C + i]
#define SP(p) (sas_params + (int64_t)(p) * 8 * C + i)
    const double C_in = F(RSF_C_in);
    const double alpha_q = ap.alpha_q ? ap.alpha_q[i] : 1.0;
    const int64_t o0 = cx.ring(0);
    double sum_sa_s_m1 = 0;
    if (STATS) {  // vs.sa_s[taum1] as this step's soil kernel recomputes it
        double sm[1];
        stream_sums<1, 2>(cx.plan,
            [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = s_rz[o]; L[1] = s_ss[o]; },
            [&](int, const double* L, double* v) { v[0] = L[0] + L[1] * m; }, sm);
        sum_sa_s_m1 = sm[0];
    }
    FluxStats fs_evap, fs_transp, fs_qss;
    double C_transp_v = 0, C_q_ss_v = 0, C_inf1 = 0, C_inf2 = 0, C_inf3 = 0;
    if (crop.re) {  // svat_crop_transport_model_deterministic (transport.py:1042-1051): redistribution opens the day
        // crop.py:2461-2558 (root growth: subsoil -> root zone), :2729-2828 (root loss: root zone -> subsoil): the flux is limited to the
        // water the source layer holds, the solute leaves with partition coefficient 1
        auto redistribute = [&](int src, int dst, double re_in, const double* sp, int k_re, int k_C, int k_M) {
            double sm[1];
            double* __restrict__ s_src = cx.pl.sa[src];
            stream_sums<1, 1>(cx.plan, [&](int j, double* L) { L[0] = s_src[cx.ring(j)]; }, [&](int, const double* L, double* v) { v[0] = L[0]; }, sm);
            const double re = ((re_in > sm[0]) ? sm[0] : re_in) * m;
            const double M = anion_withdrawal<false>(cx, src, dst, re, sp + i, C, m, 1.0, true, -1, nullptr);
            crop.out[(int64_t)k_re * C + i] = re;
            crop.out[(int64_t)k_C * C + i] = ((re > 0) ? M / re : 0) * m;
            crop.out[(int64_t)k_M * C + i] = M * m;
        };
        redistribute(1, 0, crop.re[i], crop.sp_rg, RCO_re_rg, RCO_C_re_rg, RCO_M_re_rg);
        redistribute(0, 1, crop.re[C + i], crop.sp_rl, RCO_re_rl, RCO_C_re_rl, RCO_M_re_rl);
    }
    {   // infiltration into the root zone
        const double c1 = ((F(RSF_inf_mat_rz) > 0) ? C_in : 0) * m, c2 = ((F(RSF_inf_pf_rz) > 0) ? C_in : 0) * m;
        const double M1 = c1 * F(RSF_inf_mat_rz) * m, M2 = c2 * F(RSF_inf_pf_rz) * m;
        O(RSO_C_inf_mat_rz) = c1; O(RSO_C_inf_pf_rz) = c2; O(RSO_C_iso_inf_mat_rz) = M1; O(RSO_C_iso_inf_pf_rz) = M2;
        C_inf1 = c1; C_inf2 = c2;
        s_rz[o0] = s_rz[o0] + (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) * m);
        m_rz[o0] = m_rz[o0] + (M1 + M2 * m);
    }
    double C_evap_v = 0;
    {   // soil evaporation: water only (evapotranspiration.py:620-651); the virtual tracer leaves with it, partition coefficient 1 (:723-791)
        const bool vt = ap.tracer != 0;
        const double M = anion_withdrawal<STATS>(cx, 0, -1, F(RSF_evap_soil), SP(RSP_evap_soil), C, m, vt ? 1.0 : 0.0, vt, RSP_evap_soil, &fs_evap);
        C_evap_v = vt ? ((F(RSF_evap_soil) > 0) ? M / F(RSF_evap_soil) : 0) * m : 0.0;
        O(RSO_C_evap_soil) = C_evap_v;
        O(RSO_C_iso_evap_soil) = vt ? M * m : 0.0;
    }
    {   // transpiration: crop nutrient uptake stops above 80 % saturation (evapotranspiration.py:931-938)
        const double lu = ap.lu_id ? ap.lu_id[i] : 0.0;
        bool crop_wet = false;
        if ((lu > 500) && (lu < 599)) {
            double sm[1];
            stream_sums<1, 1>(cx.plan, [&](int j, double* L) { L[0] = s_rz[cx.ring(j)]; }, [&](int, const double* L, double* v) { v[0] = L[0]; }, sm);
            const double sat = ap.S_sat_rz ? ap.S_sat_rz[i] : __longlong_as_double(0x7ff0000000000000LL);
            crop_wet = sm[0] >= 0.8 * sat;
        }
        const double alpha = (crop_wet ? 0 : (ap.alpha_transp ? ap.alpha_transp[i] : 1.0)) * m;
        const double M = anion_withdrawal<STATS>(cx, 0, -1, F(RSF_transp), SP(RSP_transp), C, m, alpha, true, RSP_transp, &fs_transp);
        C_transp_v = ((F(RSF_transp) > 0) ? M / F(RSF_transp) : 0) * m;
        O(RSO_C_transp) = C_transp_v;
        O(RSO_C_iso_transp) = M * m;
    }
    {
        const double M = anion_withdrawal<STATS>(cx, 0, 1, F(RSF_q_rz), SP(RSP_q_rz), C, m, alpha_q, true, RSP_q_rz, nullptr);
        O(RSO_C_q_rz) = ((F(RSF_q_rz) > 0) ? M / F(RSF_q_rz) : 0) * m;
        O(RSO_C_iso_q_rz) = M * m;
    }
    {   // infiltration into the subsoil
        const double c3 = ((F(RSF_inf_pf_ss) > 0) ? C_in : 0) * m;
        const double M3 = c3 * F(RSF_inf_pf_ss) * m;
        O(RSO_C_inf_pf_ss) = c3; O(RSO_C_iso_inf_pf_ss) = M3;
        C_inf3 = c3;
        s_ss[o0] = s_ss[o0] + F(RSF_inf_pf_ss) * m;
        m_ss[o0] = m_ss[o0] + M3 * m;
    }
    {
        const double M = anion_withdrawal<STATS>(cx, 1, -1, F(RSF_q_ss), SP(RSP_q_ss), C, m, alpha_q, true, RSP_q_ss, &fs_qss);
        C_q_ss_v = ((F(RSF_q_ss) > 0) ? M / F(RSF_q_ss) : 0) * m;
        O(RSO_C_q_ss) = C_q_ss_v;
        O(RSO_C_iso_q_ss) = M * m;
    }
    {
        const double M = anion_withdrawal<STATS>(cx, 1, 0, F(RSF_cpr_rz), SP(RSP_cpr_rz), C, m, alpha_q, true, RSP_cpr_rz, nullptr);
        O(RSO_C_cpr_rz) = ((F(RSF_cpr_rz) > 0) ? M / F(RSF_cpr_rz) : 0) * m;
        O(RSO_C_iso_cpr_rz) = M * m;
    }
    if (NITRATE) nitrogen_cycle(cx, *nitrate, F(RSF_transp), m);  // transport.py:966-969
    {   // storage kernels (root_zone.py:219-261, subsoil.py anion, soil.py anion): clip, solute of empty bins removed, masses
        double sums[6];
        stream_sums<6, 4>(cx.plan,
            [&](int j, double* L) {
                const int64_t o = cx.ring(j);
                L[0] = s_rz[o]; L[1] = s_ss[o]; L[2] = m_rz[o]; L[3] = m_ss[o];
            },
            [&](int j, const double* L, double* v) {
                const int64_t o = cx.ring(j);
                double a = L[0];
                a = (a < 1e-8) ? 0 : a;
                double b = L[1];
                b = (b < 1e-8) ? 0 : b;
                const double ma = (a <= 0) ? 0 : L[2], mb = (b <= 0) ? 0 : L[3];
                s_rz[o] = a; s_ss[o] = b; m_rz[o] = ma; m_ss[o] = mb;
                const double ms = ma + mb * m;
                v[0] = isnan_d(ma) ? 0 : ma;
                v[1] = a;
                v[2] = isnan_d(mb) ? 0 : mb;
                v[3] = b;
                v[4] = isnan_d(ms) ? 0 : ms;
                v[5] = a + b * m;
            }, sums);
        const double M_rz = sums[0] * m, M_ss = sums[2] * m, M_s = sums[4] * m;
        O(RSO_C_iso_rz) = M_rz; O(RSO_C_rz) = (sums[1] > 0) ? M_rz / sums[1] : 0;
        O(RSO_C_iso_ss) = M_ss; O(RSO_C_ss) = (sums[3] > 0) ? M_ss / sums[3] : 0;
        const double C_s = (sums[5] > 0) ? M_s / sums[5] : 0;
        O(RSO_C_iso_s) = M_s;   O(RSO_C_s) = C_s;
        if (STATS) {  // age statistics (water only) and balance errors (numerics.py:424-445, :640-673)
#define ST(k) stats[(int64_t)(k) * C + i]
            for (int k = 0; k < 5; k++) { ST(RSS_tt10_transp + k) = fs_transp.pct[k]; ST(RSS_tt10_q_ss + k) = fs_qss.pct[k]; }
            ST(RSS_ttavg_transp) = fs_transp.avg;
            ST(RSS_ttavg_q_ss) = fs_qss.avg;
            const double p5[5] = {0.1, 0.25, 0.5, 0.75, 0.9}, p3[3] = {0.25, 0.5, 0.75};
            double pc[5], avg;
            storage_age_statistics(cx, 0, m, 3, p3, pc, &avg);
            for (int k = 0; k < 3; k++) ST(RSS_rt25_rz + k) = pc[k];
            ST(RSS_rtavg_rz) = avg;
            storage_age_statistics(cx, 1, m, 3, p3, pc, &avg);
            for (int k = 0; k < 3; k++) ST(RSS_rt25_ss + k) = pc[k];
            ST(RSS_rtavg_ss) = avg;
            storage_age_statistics(cx, 2, m, 5, p5, pc, &avg);
            for (int k = 0; k < 5; k++) ST(RSS_rt10_s + k) = pc[k];
            ST(RSS_rtavg_s) = avg;
            auto nan0 = [](double x) { return (x != x) ? 0.0 : x; };
            const double Cp = C_s_prev[i];
            ST(RSS_dS_num_error) = fabs((sums[5] - sum_sa_s_m1) -
                                        (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - fs_evap.wsum - fs_transp.wsum - fs_qss.wsum) * 1.0);
            if (NITRATE)  // numerics.py:673-705: the denitrified and the nitrified mass of the day as well
                ST(RSS_dC_num_error) = fabs((sums[5] * C_s - sum_sa_s_m1 * Cp) -
                                            (F(RSF_inf_mat_rz) * nan0(C_inf1) + F(RSF_inf_pf_rz) * nan0(C_inf2) + F(RSF_inf_pf_ss) * nan0(C_inf3) -
                                             fs_transp.wsum * nan0(C_transp_v) - fs_qss.wsum * nan0(C_q_ss_v) -
                                             nitrate->out[(int64_t)RNO_denit_s * C + i] + nitrate->out[(int64_t)RNO_nit_s * C + i]) * 1.0);
            else if (ap.tracer)  // numerics.py:517-549: the evaporated tracer as well
                ST(RSS_dC_num_error) = fabs((sums[5] * C_s - sum_sa_s_m1 * Cp) -
                                            (F(RSF_inf_mat_rz) * nan0(C_inf1) + F(RSF_inf_pf_rz) * nan0(C_inf2) + F(RSF_inf_pf_ss) * nan0(C_inf3) -
                                             fs_evap.wsum * nan0(C_evap_v) - fs_transp.wsum * nan0(C_transp_v) - fs_qss.wsum * nan0(C_q_ss_v)) * 1.0);
            else
                ST(RSS_dC_num_error) = fabs((sums[5] * C_s - sum_sa_s_m1 * Cp) -
                                            (F(RSF_inf_mat_rz) * nan0(C_inf1) + F(RSF_inf_pf_rz) * nan0(C_inf2) + F(RSF_inf_pf_ss) * nan0(C_inf3) -
                                             fs_transp.wsum * nan0(C_transp_v) - fs_qss.wsum * nan0(C_q_ss_v)) * 1.0);
            ST(RSS_sanity_ok) = __longlong_as_double(0x7ff8000000000000LL);  // the anion branch of sanity_check is not evaluated
            C_s_prev[i] = C_s;
#undef ST
        }
    }
    {   // ageing (transport.py:622-679) as a rotation of the plane ring: the oldest two bins merge, bin 0 is emptied
        const int b_new = (P.base == 0) ? A - 1 : P.base - 1;
        const int b_last = (b_new == 0) ? A - 1 : b_new - 1;
        for (int L = 0; L < 2; L++) {
            double* __restrict__ sa = cx.pl.sa[L];
            double* __restrict__ msa = cx.pl.msa[L];
            double s_new = sa[(int64_t)b_last * C + i] + sa[(int64_t)b_new * C + i];
            s_new = (s_new < 1e-8) ? 0 : s_new;
            sa[(int64_t)b_last * C + i] = s_new;
            msa[(int64_t)b_last * C + i] = msa[(int64_t)b_last * C + i] + msa[(int64_t)b_new * C + i];
            sa[(int64_t)b_new * C + i] = 0;
            msa[(int64_t)b_new * C + i] = 0;
            if (NITRATE) {  // transport.py:809-824: the mineral nitrogen ages like the solute mass
                double* __restrict__ nm = cx.pl.nmin[L];
                nm[(int64_t)b_last * C + i] = nm[(int64_t)b_last * C + i] + nm[(int64_t)b_new * C + i];
                nm[(int64_t)b_new * C + i] = 0;
            }
        }
    }
#undef F
#undef O
#undef SP
}

// infiltration of `inf` with concentration C_in into bin 0 (infiltration.py:2218-2345, :2441-2500)
__device__ __forceinline__ void inflow(const CellCtx& cx, int layer, double inf, double C_in, double m, double* C_out, double* C_iso_out) {
    const SasParams& P = cx.P;
    double* __restrict__ sa = cx.pl.sa[layer];
    double* __restrict__ msa = cx.pl.msa[layer];
    *C_out = ((inf > 0) ? C_in : 0) * m;
    *C_iso_out = conc_to_delta(P, *C_out) * m;
    const double tt0 = ((inf > 0) ? 1 : 0) * m;
    const double mtt0 = ((inf > 0) ? C_in : 0) * m;
    Cursor c(P, cx.i);
    // bins j > 0 receive nothing: msa*(sa/sa) + 0 == msa, so with m == 1 only bin 0 changes
    const int n = (m == 1.0) ? 1 : P.A;
    for (int j = 0; j < n; j++, c.next()) {
        const double ttj = (j == 0) ? tt0 : 0.0, mttj = (j == 0) ? mtt0 : 0.0;
        const double s = sa[c.off], ms = msa[c.off];
        double v;
        if (inf * ttj + s > 0) v = ms * (s / (ttj * inf + s)) + mttj * ((ttj * inf) / (inf * ttj + s));
        else v = ms;
        msa[c.off] = v * m;
        if (j == 0) sa[c.off] = s + inf * m;
    }
}

// residence-time statistics of a storage (transport.py:161-312): RT = SA / max(SA) with SA = cumsum(sa) * m of the layer
// (which 0: root zone, 1: subsoil, 2: soil = root zone + subsoil * m), percentile ages and the mean residence time
__device__ __noinline__ void storage_age_statistics(const CellCtx& cx, int which, double m, int n_pct, const double* pcts, double* pct_out,
                                                    double* avg_out) {
    cx.assume_spaces();
    const SasParams& P = cx.P;
    const int A = P.A;
    const double* __restrict__ s_rz = cx.pl.sa[0];
    const double* __restrict__ s_ss = cx.pl.sa[1];
    auto value = [&](int j) {
        const int64_t o = cx.ring(j);
        if (which == 0) return s_rz[o];
        if (which == 1) return s_ss[o];
        return s_rz[o] + s_ss[o] * m;
    };
    NpMax mx;
    const double SA0 = 0. * m;
    mx.init(SA0);
    {
        double acc = 0;
        for (int j = 0; j < A; j++) { acc += value(j); mx.push(acc * m); }
    }
    const double mxv = mx.get();
    PctTracker trk[5];
    for (int k = 0; k < 5; k++) trk[k].init(k < n_pct ? pcts[k] : 0.5);
    double acc = 0;
    double RT_prev = (mxv > 0) ? SA0 / mxv : 0;
    NpMax rt_max;
    bool first_bin = true;
    double sums[2];
    stream_sums<2, 1>(cx.plan,
        [&](int j, double* L) { L[0] = value(j); },
        [&](int j, const double* L, double* v) {
            acc += L[0];
            const double RT = (mxv > 0) ? (acc * m) / mxv : 0;
            const double rt = RT - RT_prev;
            RT_prev = RT;
            v[0] = rt;
            v[1] = (double)(j + 1) * rt;
            if (first_bin) { rt_max.init(RT); first_bin = false; } else rt_max.push(RT);
            for (int k = 0; k < 5; k++) trk[k].push(j, RT);
        }, sums);
    for (int k = 0; k < n_pct; k++) pct_out[k] = trk[k].result(A, rt_max.get());
    *avg_out = (sums[0] > 0) ? sums[1] : __longlong_as_double(0x7ff8000000000000LL);
}

// storage kernels (root_zone.py:189-215, subsoil.py:153-180, soil.py:1036-1092) followed by ageing
// (transport.py:682-739) as a ring rotation
__device__ __forceinline__ void storage_concentrations(const CellCtx& cx, double m, double* out, int64_t oC, double* sum_sa_s, double* C_s_out,
                                                       double* sum_sa_rz, double* sum_sa_ss) {
    const SasParams& P = cx.P;
    const int A = P.A;
    double* __restrict__ s_rz = cx.pl.sa[0];
    double* __restrict__ s_ss = cx.pl.sa[1];
    double* __restrict__ m_rz = cx.pl.msa[0];
    double* __restrict__ m_ss = cx.pl.msa[1];
    double sums[6];
    stream_sums<6, 4>(cx.plan,
        [&](int j, double* L) {
            const int64_t o = cx.ring(j);
            L[0] = s_rz[o]; L[1] = s_ss[o]; L[2] = m_rz[o]; L[3] = m_ss[o];
        },
        [&](int j, const double* L, double* v) {
            const int64_t o = cx.ring(j);
            double a = L[0];
            a = (a < 1e-8) ? 0 : a;
            s_rz[o] = a;
            double b = L[1];
            b = (b < 1e-8) ? 0 : b;
            s_ss[o] = b;
            const double ma = L[2], mb = L[3];
            v[0] = a;
            v[1] = ma * a;
            v[2] = b;
            v[3] = mb * b;
            const double s = a + b * m;
            const double tot = a + b;
            double ms = (tot > 0) ? ma * (a / tot) + mb * (b / tot) : 0;
            ms = isnan_d(ms) ? 0 : ms;
            v[4] = s;
            v[5] = ms * s;
        }, sums);
    const double C_rz = ((sums[0] > 0) ? sums[1] / sums[0] : 0) * m;
    const double C_ss = ((sums[2] > 0) ? sums[3] / sums[2] : 0) * m;
    const double C_s = ((sums[4] > 0) ? sums[5] / sums[4] : 0) * m;
    const int64_t i = cx.i;
    out[RSO_C_rz * oC + i] = C_rz;
    out[RSO_C_ss * oC + i] = C_ss;
    out[RSO_C_s * oC + i] = C_s;
    out[RSO_C_iso_rz * oC + i] = conc_to_delta(P, C_rz) * m;
    out[RSO_C_iso_ss * oC + i] = conc_to_delta(P, C_ss) * m;
    out[RSO_C_iso_s * oC + i] = conc_to_delta(P, C_s) * m;
    *sum_sa_s = sums[4];
    *sum_sa_rz = sums[0];
    *sum_sa_ss = sums[2];
    *C_s_out = C_s;
}

__device__ __forceinline__ void ageing(const CellCtx& cx) {
    const SasParams& P = cx.P;
    const int A = P.A;
    const int64_t i = cx.i;
    // ageing: new bin j = old bin j-1; the oldest two bins merge; bin 0 is emptied.  With base' = base - 1 the old
    // bin A-1 becomes the new bin 0 and the old bin A-2 the new bin A-1.
    const int64_t C = P.C;
    const int b_new = (P.base == 0) ? A - 1 : P.base - 1;            // plane of old bin A-1 = new bin 0
    const int b_last = (b_new == 0) ? A - 1 : b_new - 1;             // plane of old bin A-2 = new bin A-1
    for (int L = 0; L < 2; L++) {
        double* __restrict__ sa = cx.pl.sa[L];
        double* __restrict__ msa = cx.pl.msa[L];
        const double sam1_last = sa[(int64_t)b_new * C + i], msam1_last = msa[(int64_t)b_new * C + i];
        const double s_l = sa[(int64_t)b_last * C + i], ms_l = msa[(int64_t)b_last * C + i];
        const double tot = s_l + sam1_last;
        double v = (tot > 0) ? msam1_last * (sam1_last / tot) + ms_l * (s_l / tot) : 0;
        v = isnan_d(v) ? 0 : v;
        double s_new = s_l + sam1_last;
        s_new = (s_new < 1e-8) ? 0 : s_new;
        v = (s_new <= 0) ? __longlong_as_double(0x7ff8000000000000LL) : v;
        sa[(int64_t)b_last * C + i] = s_new;
        msa[(int64_t)b_last * C + i] = v;
        sa[(int64_t)b_new * C + i] = 0;
        msa[(int64_t)b_new * C + i] = 0;
    }
}


// ================================================================================================================
// Explicit Euler solver (settings.sas_solver = "Euler"): svat_transport_model_euler (transport.py:2063-2413) with
// calc_TT_num (:859-907) and calc_TT_num_nonneg (:910-945), isotope branch.  `substeps` Euler steps per day
// (calculate_storage_selection, :3221-3258; settings.h = 1 / substeps); all five fluxes are evaluated from ONE
// StorAge snapshot per step, so a step is three walks over the age axis:
//   P1  S = max(cumsum(sa)) of both layers,
//   P2  the un-normalised non-negative pdf of every flux, summed with NumPy's pairwise tree (the normalisation constant),
//   P3  the same pdf again -> normalised -> cumulated -> differenced (tt), isotope mixing and StorAge update of both
//       layers, the storage kernels' `sa < 1e-8 -> 0` fused into the store.
// The flux concentrations -- and the travel-time distributions handed out -- are those of the day's LAST step (every
// earlier step's are overwritten in the reference as well), so only that step carries the ten concentration sums.
// ================================================================================================================
#ifndef ROGER_EULER_ROLLED
#define ROGER_EULER_ROLLED 1
#endif
#if ROGER_EULER_ROLLED
#define EULER_SUMS stream_sums_rolled
#define EULER_SUMS_ARG A,
#else
#define EULER_SUMS stream_sums
#define EULER_SUMS_ARG
#endif
struct EulerFlux {      // the hot per-flux scalars (registers); the Family descriptors of the general path live beside them
    double flux, fh;    // flux, flux * h
    bool active;        // calc_TT_num zeroes TTq where flux * h <= 0
    double om_prev;     // TTq of the previous age index
    double ssum;        // np.sum(ttq_nonneg)
    double accT, Tprev; // cumsum of the normalised pdf
    // exact shortcuts (same bits as the general path, see euler_q)
    bool sa_only;       // the SAS function depends on the cumulative StorAge only, not on the age index
    bool powfast;       // power law, cell inside the catchment, S in the range where div_by() is exact
    bool ssum_ok;       // ssum in that range as well
    double S, rS, y, rssum;  // S = max(SA) * m of the flux's layer, its reciprocal, the power-law exponent
};

// log(SA / S) of one layer's bin, shared by the layer's power-law fluxes (x^y1, x^y2, ... of one base: cf. roger_pow2)
struct EulerLog {
    double hi, lo;
    bool in, easy;
};
__device__ __forceinline__ EulerLog euler_log(const CellCtx& cx, double SA, double S, double rS) {
    EulerLog g;
    g.in = (SA > 0) && (SA <= S);
    g.easy = g.in && (SA >= 0x1p-500);  // keeps the quotient normal and div_by() exact
    const double xr = div_by(SA, S, rS);
    g.hi = roger_pow_log(g.easy ? roger_pow_asu(xr) : 0x3fe0000000000000ULL, &g.lo, cx.tabs.log_tab);
    return g;
}

// ttq_nonneg of bin j (calc_TT_num_nonneg :919-931) given the masked cumulative StorAge at its upper edge
__device__ __forceinline__ double euler_q(EulerFlux& e, const Family& fam, const CellCtx& cx, int j, double SA, double sa_d, const EulerLog& g) {
    double om;
    if (!e.active) om = 0.0;
    else if (e.sa_only && sa_d == 0) om = e.om_prev;  // an empty bin: same cumulative StorAge, same SAS function value
    else if (e.powfast) {                               // omega() for the power law with the division hoisted (cf. pow_pass_bins)
        const double S = e.S;
        double p = roger_pow_from_log(g.hi, g.lo, e.y, cx.tabs);
        if (g.in && !g.easy) p = roger_pow(SA / S, e.y, cx.tabs);
        om = (SA > 0) ? ((SA <= S) ? p : 1.) : 0.;
    } else om = omega(fam, cx.P, cx.tabs, j + 1, SA);
    double q = (om - e.om_prev) * e.fh;
    e.om_prev = om;
    q = (sa_d + q < 0) ? -sa_d : q;
    q = (q == 0) ? 0.0 : q;  // where(ttq == -0, 0, ttq)
    return q;
}

// the two-statement isotope mixing of transport.py:2115-2130 / :2278-2293 for one bin
__device__ __forceinline__ double euler_mix(double msa, double sa, double dsa1, double dmsa1, bool need_pos) {
    const double tot = dsa1 + sa;
    const bool c1 = need_pos ? ((tot > 0) && (msa > 0)) : (tot > 0);
    double v = (c1 ? msa * (sa / tot) : 0) + ((tot > 0) ? dmsa1 * (dsa1 / tot) : 0);
    v = ((dsa1 > 0) && (v <= 0)) ? dmsa1 : v;
    return v;
}

struct EulerStats {  // diagnostics of the day's last Euler step (STATS kernels): evap_soil, transp, q_ss
    double wsum[3];    // sum(flux * tt), numerics.py:437-439
    double pct[2][5];  // percentile ages of the transpiration / subsoil-percolation travel-time distributions
    double avg[2];     // mean travel times
};

// ANION: the conservative-solute branch (bromide / chloride; transport.py:2144-2156, :2320-2343, :2378-2407): msa holds solute
// mass, calc_mtt (:582-596) with the partition coefficients alpha_transp / alpha_q, no mixing statements, the anion storage
// kernels' `msa = 0 where sa <= 0` fused into the store.
template <bool LAST, bool STATS, bool ANION>
__device__ __noinline__ void euler_step(const CellCtx& cx, const double* __restrict__ forcing, double m, const double* __restrict__ sas_params,
                                        double* __restrict__ out, EulerStats* es, double alpha_t, double alpha_q) {
    const SasParams& P = cx.P;
    const int A = P.A;
    const int64_t C = P.C, i = cx.i;
    double* __restrict__ s_rz = cx.pl.sa[0];
    double* __restrict__ s_ss = cx.pl.sa[1];
    double* __restrict__ m_rz = cx.pl.msa[0];
    double* __restrict__ m_ss = cx.pl.msa[1];
    cx.assume_spaces();
    auto nan0 = [](double x) { return (x != x) ? 0.0 : x; };
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#define F(k) forcing[(int64_t)(k) * C + i]
#define O(k) out[(int64_t)(k) * C + i]
#define SP(p) (sas_params + (int64_t)(p) * 8 * C + i)
    const double h = 1.0 / (double)P.substeps;
    const double inf_mat = F(RSF_inf_mat_rz), inf_pf = F(RSF_inf_pf_rz), inf_ss = F(RSF_inf_pf_ss), C_in = F(RSF_C_in);
    {   // upper boundary condition :2071-2143: infiltration enters age bin 0 (bins j > 0: see clean_* below)
        const int64_t o = cx.ring(0);
        const double tt_im = ((inf_mat > 0) ? 1 : 0) * m, tt_ip = ((inf_pf > 0) ? 1 : 0) * m, tt_is = ((inf_ss > 0) ? 1 : 0) * m;
        const double mtt_im = ((inf_mat > 0) ? C_in : 0) * m, mtt_ip = ((inf_pf > 0) ? C_in : 0) * m, mtt_is = ((inf_ss > 0) ? C_in : 0) * m;
        const double dsa_rz = (inf_mat * tt_im + inf_pf * tt_ip) * h;
        const double dsa_ss = (inf_ss * tt_is) * h;
        const double dmsa_rz1 = nan0(mtt_im) * ((dsa_rz > 0) ? ((inf_mat * tt_im * h) / dsa_rz) : 0) +
                                nan0(mtt_ip) * ((dsa_rz > 0) ? ((inf_pf * tt_ip * h) / dsa_rz) : 0);
        const double dmsa_ss1 = nan0(mtt_is) * ((dsa_ss > 0) ? ((inf_ss * tt_is * h) / dsa_ss) : 0);
        double a = s_rz[o], b = s_ss[o];
        if (ANION) {  // solute enters as mtt_inf * h with mtt_inf[0] = C_in * mask in all three infiltration arrays (:2089-2101, :2150-2156)
            const double mi = nan0(C_in * m);
            s_rz[o] = a + dsa_rz;
            s_ss[o] = b + dsa_ss;
            m_rz[o] = m_rz[o] + (mi * h + mi * h);
            m_ss[o] = m_ss[o] + mi * h;
        } else {
            double ma = euler_mix(m_rz[o], a, dsa_rz, dmsa_rz1, true);
            double mb = euler_mix(m_ss[o], b, dsa_ss, dmsa_ss1, false);
            a += dsa_rz;
            b += dsa_ss;
            s_rz[o] = a;
            s_ss[o] = b;
            m_rz[o] = (a <= 0) ? 0 : ma;
            m_ss[o] = (b <= 0) ? 0 : mb;
        }
    }
    // what the boundary-condition statements do to a bin that receives nothing (dsa1 = 0, dmsa1 = 0):
    //   msa*(sa/(0 + sa)) + 0*(0/sa) under the statements' conditions, then msa = 0 where sa <= 0
    // (the rare sides of these shortcuts sit behind `asm volatile("")`: without it the compiler if-converts the selects and
    // executes the divisions anyway -- 5 of the 8 divisions per bin and step in the first capture)
    auto self_ratio = [](double x) {  // x / x: exactly 1 for a finite non-zero x
        double one = 1.0;
        if (!(x < 1.7976931348623157e308)) { asm volatile(""); one = x / x; }
        return one;
    };
    auto clean_rz = [&](double ms, double s) { return ((s > 0) && (ms > 0)) ? ms * self_ratio(s) + 0.0 : 0.0; };
    auto clean_ss = [&](double ms, double s) { return (s > 0) ? ms * self_ratio(s) + 0.0 : 0.0; };
    // P1: S of both layers (calc_SA :343-359 times maskCatch, np.max over nages points)
    NpMax mx_rz, mx_ss;
    mx_rz.init(0.0 * m);
    mx_ss.init(0.0 * m);
    {
        double acc_a = 0, acc_b = 0;
        Cursor c(P, i);
        // SAS family 52 (allocated by roger_sas_set_parameters when some cell uses it) reads the masked cumulative StorAge backwards:
        // a second copy of the walk leaves it in the scratch planes, so that the plain walk keeps its loads free of aliasing stores
        if (cx.pl.rev[0] != nullptr) {
            double* __restrict__ rev_a = cx.pl.rev[0];
            double* __restrict__ rev_b = cx.pl.rev[1];
            for (int j = 0; j < A; j++, c.next()) {
                acc_a += s_rz[c.off];
                acc_b += s_ss[c.off];
                mx_rz.push(acc_a * m);
                mx_ss.push(acc_b * m);
                rev_a[cx.lin(j)] = acc_a * m;
                rev_b[cx.lin(j)] = acc_b * m;
            }
        } else
        for (int j = 0; j < A; j++, c.next()) {
            acc_a += s_rz[c.off];
            acc_b += s_ss[c.off];
            mx_rz.push(acc_a * m);
            mx_ss.push(acc_b * m);
        }
    }
    EulerFlux fx[5];  // evap_soil, transp, q_rz (root zone); cpr_rz, q_ss (subsoil)
    Family fams[5];
    {
        const int procs[5] = {RSP_evap_soil, RSP_transp, RSP_q_rz, RSP_cpr_rz, RSP_q_ss};
        const double flx[5] = {F(RSF_evap_soil), F(RSF_transp), F(RSF_q_rz), F(RSF_cpr_rz), F(RSF_q_ss)};
        const double mxa = mx_rz.get(), mxb = mx_ss.get();
#pragma unroll
        for (int k = 0; k < 5; k++) {
            family_begin(fams[k], SP(procs[k]), C, m);
            if (fams[k].sp[0] == 52) {  // read backwards by omega(): the masked cumulative StorAge P1 left in the scratch planes
                fams[k].rev = cx.pl.rev[(k < 3) ? 0 : 1] + i;
                fams[k].rev_stride = C;
            }
            family_prepare(fams[k], (k < 3) ? mxa : mxb);
            fx[k].flux = flx[k];
            fx[k].fh = flx[k] * h;
            fx[k].active = !(fx[k].fh <= 0);
            const Family& f = fams[k];
            fx[k].sa_only = f.fast && !f.is_uniform && !f.is_dirac && !f.is_exp52;
            fx[k].powfast = f.fast && f.is_pow && (f.S > 0) && div_range_ok(f.S) && (f.sp[1] - f.sp[1] == 0.0);
            fx[k].S = f.S;
            fx[k].y = f.sp[1];
            fx[k].rS = 1.0 / f.S;
        }
    }
    // one logarithm per layer and bin serves every power-law flux of the layer (skipped for empty bins)
    const bool logs_a = (fx[0].active && fx[0].powfast) || (fx[1].active && fx[1].powfast) || (fx[2].active && fx[2].powfast);
    const bool logs_b = (fx[3].active && fx[3].powfast) || (fx[4].active && fx[4].powfast);
    const double S_a = fx[0].S, rS_a = fx[0].rS, S_b = fx[3].S, rS_b = fx[3].rS;
    auto layer_log = [&](int layer, double SA, double sa_d) {
        EulerLog g{0.0, 0.0, false, false};
        if ((layer == 0 ? logs_a : logs_b) && !(sa_d == 0)) g = euler_log(cx, SA, layer == 0 ? S_a : S_b, layer == 0 ? rS_a : rS_b);
        return g;
    };
    auto reset = [&]() {
#pragma unroll
        for (int k = 0; k < 5; k++) fx[k].om_prev = fx[k].active ? omega(fams[k], P, cx.tabs, 0, 0.0 * m) : 0.0;
    };
    // P2: normalisation constants
    {
        reset();
        double acc_a = 0, acc_b = 0, SAa = 0.0 * m, SAb = 0.0 * m;
        double sums[5];
        EULER_SUMS<5, 2>(cx.plan, EULER_SUMS_ARG
            [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = s_rz[o]; L[1] = s_ss[o]; },
            [&](int j, const double* L, double* v) {
                acc_a += L[0];
                acc_b += L[1];
                const double na = acc_a * m, nb = acc_b * m;
                const double da = na - SAa, db = nb - SAb;
                SAa = na;
                SAb = nb;
                const EulerLog ga = layer_log(0, na, da), gb = layer_log(1, nb, db);
#pragma unroll
                for (int k = 0; k < 5; k++) v[k] = euler_q(fx[k], fams[k], cx, j, (k < 3) ? na : nb, (k < 3) ? da : db, (k < 3) ? ga : gb);
            }, sums);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            fx[k].ssum = sums[k];
            fx[k].ssum_ok = (sums[k] > 0) && div_range_ok(sums[k]);
            fx[k].rssum = 1.0 / sums[k];
        }
    }
    // P3: travel-time distributions, mixing, update
    {
        reset();
#pragma unroll
        for (int k = 0; k < 5; k++) { fx[k].accT = 0; fx[k].Tprev = 0; }
        double acc_a = 0, acc_b = 0, SAa = 0.0 * m, SAb = 0.0 * m;
        const double evap = fx[0].flux, transp = fx[1].flux, q_rz = fx[2].flux, cpr = fx[3].flux, q_ss = fx[4].flux;
        double* __restrict__ tt_base = (LAST && P.store_tt && cx.pl.tt) ? cx.pl.tt : nullptr;
        PctTracker trk[(LAST && STATS) ? 10 : 1];
        double cdf_t = 0, cdf_s = 0;
        NpMax cmax_t, cmax_s;
        if (LAST && STATS) {
            const double pcts[5] = {0.1, 0.25, 0.5, 0.75, 0.9};
#pragma unroll
            for (int k = 0; k < 5; k++) { trk[k].init(pcts[k]); trk[5 + k].init(pcts[k]); }
        }
        auto bin = [&](int j, const double* L, double* v) {
            const int64_t o = cx.ring(j);
            const double a = L[0], b = L[1];
            const double ma = (ANION || j == 0) ? L[2] : clean_rz(L[2], a);
            const double mb = (ANION || j == 0) ? L[3] : clean_ss(L[3], b);
            acc_a += a;
            acc_b += b;
            const double na = acc_a * m, nb = acc_b * m;
            const double da = na - SAa, db = nb - SAb;
            SAa = na;
            SAb = nb;
            double tt[5];
            const EulerLog ga = layer_log(0, na, da), gb = layer_log(1, nb, db);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const double q = euler_q(fx[k], fams[k], cx, j, (k < 3) ? na : nb, (k < 3) ? da : db, (k < 3) ? ga : gb);
                const double p = (q > 0) ? ((fx[k].ssum_ok && div_range_ok(q)) ? div_by(q, fx[k].ssum, fx[k].rssum) : q / fx[k].ssum) : 0;
                fx[k].accT = (j == 0) ? p : fx[k].accT + p;  // np.cumsum starts from the first element
                const double d = fx[k].accT - fx[k].Tprev;
                fx[k].Tprev = fx[k].accT;
                tt[k] = (d >= 0) ? d : 0;
            }
            const double tt_e = tt[0], tt_t = tt[1], tt_q = tt[2], tt_c = tt[3], tt_s = tt[4];
            if (ANION) {
                auto amtt = [](double sa_, double msa_, double alpha, double t, double fh) {  // calc_mtt, transport.py:582-596
                    double w = ((sa_ > 0) ? (msa_ / sa_) : 0) * alpha * t * fh;
                    w = (w <= 0) ? 0 : w;
                    return (w > msa_) ? msa_ : w;
                };
                const double mt = amtt(a, ma, alpha_t, tt_t, fx[1].fh), mq = amtt(a, ma, alpha_q, tt_q, fx[2].fh);
                const double mc = amtt(b, mb, alpha_q, tt_c, fx[3].fh), msq = amtt(b, mb, alpha_q, tt_s, fx[4].fh);
                if (LAST) {
                    v[0] = mt; v[1] = mq; v[2] = mc; v[3] = msq;
                    if (STATS) {  // water-age statistics of the day's last step (transport.py:59-152), numerics.py:437-439
                        v[4] = evap * tt_e; v[5] = transp * tt_t; v[6] = q_ss * tt_s;
                        v[STATS ? 7 : 0] = tt_t; v[STATS ? 8 : 0] = tt_s;
                        v[STATS ? 9 : 0] = (double)(j + 1) * tt_t; v[STATS ? 10 : 0] = (double)(j + 1) * tt_s;
                        cdf_t += tt_t;
                        cdf_s += tt_s;
                        if (j == 0) { cmax_t.init(cdf_t); cmax_s.init(cdf_s); } else { cmax_t.push(cdf_t); cmax_s.push(cdf_s); }
#pragma unroll
                        for (int k = 0; k < 5; k++) { trk[k].push(j, cdf_t); trk[(STATS ? 5 : 0) + k].push(j, cdf_s); }
                    }
                    if (tt_base) {
                        const int64_t span = (int64_t)A * C, l = cx.lin(j);
                        tt_base[RSP_evap_soil * span + l] = tt_e;
                        tt_base[RSP_transp * span + l] = tt_t;
                        tt_base[RSP_q_rz * span + l] = tt_q;
                        tt_base[RSP_cpr_rz * span + l] = tt_c;
                        tt_base[RSP_q_ss * span + l] = tt_s;
                    }
                }
                double dsa_rz = (cpr * tt_c - evap * tt_e - transp * tt_t - q_rz * tt_q) * h;
                dsa_rz = (a + dsa_rz < 0) ? -a : dsa_rz;
                double dsa_ss = (q_rz * tt_q - cpr * tt_c - q_ss * tt_s) * h;
                dsa_ss = (b + dsa_ss < 0) ? -b : dsa_ss;
                double a2 = a + dsa_rz, b2 = b + dsa_ss;
                double dma = nan0(mc) - nan0(mt) - nan0(mq);
                double dmb = nan0(mq) - nan0(mc) - nan0(msq);
                dma = (ma + dma < 0) ? 0 : dma;
                dmb = (mb + dmb < 0) ? 0 : dmb;
                double ma2 = ma + dma, mb2 = mb + dmb;
                a2 = (a2 < 1e-8) ? 0 : a2;   // anion storage kernels (root_zone.py:219-240, subsoil.py), run after every Euler step
                b2 = (b2 < 1e-8) ? 0 : b2;
                ma2 = (a2 <= 0) ? 0 : ma2;
                mb2 = (b2 <= 0) ? 0 : mb2;
                s_rz[o] = a2;
                s_ss[o] = b2;
                m_rz[o] = ma2;
                m_ss[o] = mb2;
                return;
            }
            const double mtt_c = (tt_c > 0) ? mb : 0, mtt_q = (tt_q > 0) ? ma : 0;
            if (LAST) {
                v[0] = tt_e; v[1] = ((tt_e > 0) ? ma : 0) * tt_e;
                v[2] = tt_t; v[3] = ((tt_t > 0) ? ma : 0) * tt_t;
                v[4] = tt_q; v[5] = mtt_q * tt_q;
                v[6] = tt_c; v[7] = mtt_c * tt_c;
                v[8] = tt_s; v[9] = ((tt_s > 0) ? mb : 0) * tt_s;
                if (STATS) {  // transport.py:59-152 (TT[1:] = cumsum(tt)), numerics.py:437-439
                    v[10] = evap * tt_e; v[11] = transp * tt_t; v[12] = q_ss * tt_s;
                    v[13] = (double)(j + 1) * tt_t; v[14] = (double)(j + 1) * tt_s;
                    cdf_t += tt_t;
                    cdf_s += tt_s;
                    if (j == 0) { cmax_t.init(cdf_t); cmax_s.init(cdf_s); } else { cmax_t.push(cdf_t); cmax_s.push(cdf_s); }
#pragma unroll
                    for (int k = 0; k < 5; k++) { trk[k].push(j, cdf_t); trk[5 + k].push(j, cdf_s); }
                }
                if (tt_base) {
                    const int64_t span = (int64_t)A * C, l = cx.lin(j);
                    tt_base[RSP_evap_soil * span + l] = tt_e;
                    tt_base[RSP_transp * span + l] = tt_t;
                    tt_base[RSP_q_rz * span + l] = tt_q;
                    tt_base[RSP_cpr_rz * span + l] = tt_c;
                    tt_base[RSP_q_ss * span + l] = tt_s;
                }
            }
            double dsa_rz = (cpr * tt_c - evap * tt_e - transp * tt_t - q_rz * tt_q) * h;
            dsa_rz = (a + dsa_rz < 0) ? -a : dsa_rz;
            double dsa_ss = (q_rz * tt_q - cpr * tt_c - q_ss * tt_s) * h;
            dsa_ss = (b + dsa_ss < 0) ? -b : dsa_ss;
            const double dsa_rz1 = (cpr * tt_c) * h;
            // (cpr * tt_c * h) / dsa_rz1 divides a value by itself: exactly 1 for a finite positive value
            const double dmsa_rz1 = nan0(mtt_c) * ((dsa_rz1 > 0) ? self_ratio(dsa_rz1) : 0);
            const double dsa_ss1 = (q_rz * tt_q) * h;
            const double dmsa_ss1 = nan0(mtt_q) * ((dsa_ss1 > 0) ? self_ratio(dsa_ss1) : 0);
            // a bin that receives nothing keeps its concentration: the general expression reduces to clean_*()
            double ma2 = ma, mb2 = mb;
            if (!(dsa_rz1 == 0 && dmsa_rz1 == 0 && j > 0)) { asm volatile(""); ma2 = euler_mix(ma, a, dsa_rz1, dmsa_rz1, true); }
            if (!(dsa_ss1 == 0 && dmsa_ss1 == 0 && j > 0)) { asm volatile(""); mb2 = euler_mix(mb, b, dsa_ss1, dmsa_ss1, false); }
            double a2 = a + dsa_rz, b2 = b + dsa_ss;
            ma2 = (a2 <= 0) ? 0 : ma2;
            mb2 = (b2 <= 0) ? 0 : mb2;
            a2 = (a2 < 1e-8) ? 0 : a2;  // root_zone.py:196-199, subsoil.py:160-163 (run after every Euler step)
            b2 = (b2 < 1e-8) ? 0 : b2;
            s_rz[o] = a2;
            s_ss[o] = b2;
            m_rz[o] = ma2;
            m_ss[o] = mb2;
        };
        auto load = [&](int j, double* L) {
            const int64_t o = cx.ring(j);
            L[0] = s_rz[o]; L[1] = s_ss[o]; L[2] = m_rz[o]; L[3] = m_ss[o];
        };
        if (LAST) {
            constexpr int NS = ANION ? (STATS ? 11 : 4) : (STATS ? 15 : 10);
            double sums[NS];
            EULER_SUMS<NS, 4>(cx.plan, EULER_SUMS_ARG load, bin, sums);
            if (ANION && STATS && es) {
                const double qn = __longlong_as_double(0x7ff8000000000000LL);
                es->wsum[0] = sums[STATS ? 4 : 0]; es->wsum[1] = sums[STATS ? 5 : 0]; es->wsum[2] = sums[STATS ? 6 : 0];
#pragma unroll
                for (int k = 0; k < 5; k++) { es->pct[0][k] = trk[k].result(A, cmax_t.get()); es->pct[1][k] = trk[(STATS ? 5 : 0) + k].result(A, cmax_s.get()); }
                es->avg[0] = (sums[STATS ? 7 : 0] > 0) ? sums[STATS ? 9 : 0] : qn;
                es->avg[1] = (sums[STATS ? 8 : 0] > 0) ? sums[STATS ? 10 : 0] : qn;
            }
            if (ANION) {  // :2378-2407; C_evap_soil and the M_* flux planes are not touched by the Euler solver
                O(RSO_C_inf_mat_rz) = ((inf_mat * h > 0) ? C_in : 0) * m;
                O(RSO_C_inf_pf_rz) = ((inf_pf * h > 0) ? C_in : 0) * m;
                O(RSO_C_inf_pf_ss) = ((inf_ss * h > 0) ? C_in : 0) * m;
                O(RSO_C_transp) = ((transp > 0) ? sums[0] / fx[1].fh : 0) * m;
                O(RSO_C_q_rz) = ((q_rz > 0) ? sums[1] / fx[2].fh : 0) * m;
                O(RSO_C_cpr_rz) = ((cpr > 0) ? sums[2] / fx[3].fh : 0) * m;
                O(RSO_C_q_ss) = ((q_ss > 0) ? sums[3] / fx[4].fh : 0) * m;
                return;
            }
            if (STATS && es) {
                const double qn = __longlong_as_double(0x7ff8000000000000LL);
                es->wsum[0] = sums[NS - 5]; es->wsum[1] = sums[NS - 4]; es->wsum[2] = sums[NS - 3];
#pragma unroll
                for (int k = 0; k < 5; k++) { es->pct[0][k] = trk[k].result(A, cmax_t.get()); es->pct[1][k] = trk[(STATS ? 5 : 0) + k].result(A, cmax_s.get()); }
                es->avg[0] = (sums[2] > 0) ? sums[NS - 2] : qn;
                es->avg[1] = (sums[8] > 0) ? sums[NS - 1] : qn;
            }
            // calc_conc_iso_flux :512-535, delta_fluxes_svat :3659-3697
            const int oc[5] = {RSO_C_evap_soil, RSO_C_transp, RSO_C_q_rz, RSO_C_cpr_rz, RSO_C_q_ss};
            const int oi[5] = {RSO_C_iso_evap_soil, RSO_C_iso_transp, RSO_C_iso_q_rz, RSO_C_iso_cpr_rz, RSO_C_iso_q_ss};
#pragma unroll
            for (int k = 0; k < 5; k++) {
                double conc = (sums[2 * k] > 0) ? sums[2 * k + 1] / sums[2 * k] : qnan;
                conc = (conc != 0) ? conc : qnan;
                conc = conc * m;
                O(oc[k]) = conc;
                O(oi[k]) = conc_to_delta(P, conc) * m;
            }
            const double c1 = ((inf_mat > 0) ? C_in : qnan) * m, c2 = ((inf_pf > 0) ? C_in : qnan) * m, c3 = ((inf_ss > 0) ? C_in : qnan) * m;
            O(RSO_C_inf_mat_rz) = c1; O(RSO_C_iso_inf_mat_rz) = conc_to_delta(P, c1) * m;
            O(RSO_C_inf_pf_rz) = c2; O(RSO_C_iso_inf_pf_rz) = conc_to_delta(P, c2) * m;
            O(RSO_C_inf_pf_ss) = c3; O(RSO_C_iso_inf_pf_ss) = conc_to_delta(P, c3) * m;
        } else {
            double Lc[4], Ln[4], v[1];
            load(0, Lc);
            for (int j = 0; j < A; j++) {  // next bin's loads in flight while this bin computes
                if (j + 1 < A) load(j + 1, Ln);
                bin(j, Lc, v);
#pragma unroll
                for (int q = 0; q < 4; q++) Lc[q] = Ln[q];
            }
        }
    }
#undef F
#undef O
#undef SP
}

// The Runge-Kutta step (see the RK4 block below): the Euler step's building blocks as lambdas (`prepare`, `tt_of_bin`) so that the
// four stages can re-evaluate them.  A function of its own: folding it into euler_step<> cost the Euler solver 14 % (measured,
// 315 vs 276 ms per day at 262144 x 1000 -- the lambdas' captured state no longer stays in registers).
template <bool LAST, bool STATS, bool ANION, bool RK4 = true>
__device__ __noinline__ void rk4_step(const CellCtx& cx, const double* __restrict__ forcing, double m, const double* __restrict__ sas_params,
                                        double* __restrict__ out, EulerStats* es, double alpha_t, double alpha_q) {
    cx.assume_spaces();
    const SasParams& P = cx.P;
    const int A = P.A;
    const int64_t C = P.C, i = cx.i;
    double* __restrict__ s_rz = cx.pl.sa[0];
    double* __restrict__ s_ss = cx.pl.sa[1];
    double* __restrict__ m_rz = cx.pl.msa[0];
    double* __restrict__ m_ss = cx.pl.msa[1];
    ROGER_IS_GLOBAL(s_rz); ROGER_IS_GLOBAL(s_ss); ROGER_IS_GLOBAL(m_rz); ROGER_IS_GLOBAL(m_ss);
    ROGER_IS_GLOBAL(forcing); ROGER_IS_GLOBAL(sas_params); ROGER_IS_GLOBAL(out);
    auto nan0 = [](double x) { return (x != x) ? 0.0 : x; };
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#define F(k) forcing[(int64_t)(k) * C + i]
#define O(k) out[(int64_t)(k) * C + i]
#define SP(p) (sas_params + (int64_t)(p) * 8 * C + i)
    const double h = 1.0 / (double)P.substeps;
    const double inf_mat = F(RSF_inf_mat_rz), inf_pf = F(RSF_inf_pf_rz), inf_ss = F(RSF_inf_pf_ss), C_in = F(RSF_C_in);
    {   // upper boundary condition :2071-2143: infiltration enters age bin 0 (bins j > 0: see clean_* below)
        const int64_t o = cx.ring(0);
        const double tt_im = ((inf_mat > 0) ? 1 : 0) * m, tt_ip = ((inf_pf > 0) ? 1 : 0) * m, tt_is = ((inf_ss > 0) ? 1 : 0) * m;
        const double mtt_im = ((inf_mat > 0) ? C_in : 0) * m, mtt_ip = ((inf_pf > 0) ? C_in : 0) * m, mtt_is = ((inf_ss > 0) ? C_in : 0) * m;
        const double dsa_rz = (inf_mat * tt_im + inf_pf * tt_ip) * h;
        const double dsa_ss = (inf_ss * tt_is) * h;
        const double dmsa_rz1 = nan0(mtt_im) * ((dsa_rz > 0) ? ((inf_mat * tt_im * h) / dsa_rz) : 0) +
                                nan0(mtt_ip) * ((dsa_rz > 0) ? ((inf_pf * tt_ip * h) / dsa_rz) : 0);
        const double dmsa_ss1 = nan0(mtt_is) * ((dsa_ss > 0) ? ((inf_ss * tt_is * h) / dsa_ss) : 0);
        double a = s_rz[o], b = s_ss[o];
        if (ANION) {  // solute enters as mtt_inf * h with mtt_inf[0] = C_in * mask in all three infiltration arrays (:2089-2101, :2150-2156)
            const double mi = nan0(C_in * m);
            s_rz[o] = a + dsa_rz;
            s_ss[o] = b + dsa_ss;
            m_rz[o] = m_rz[o] + (mi * h + mi * h);
            m_ss[o] = m_ss[o] + mi * h;
        } else {
            double ma = euler_mix(m_rz[o], a, dsa_rz, dmsa_rz1, true);
            double mb = euler_mix(m_ss[o], b, dsa_ss, dmsa_ss1, false);
            a += dsa_rz;
            b += dsa_ss;
            s_rz[o] = a;
            s_ss[o] = b;
            m_rz[o] = (a <= 0) ? 0 : ma;
            m_ss[o] = (b <= 0) ? 0 : mb;
        }
    }
    // what the boundary-condition statements do to a bin that receives nothing (dsa1 = 0, dmsa1 = 0):
    //   msa*(sa/(0 + sa)) + 0*(0/sa) under the statements' conditions, then msa = 0 where sa <= 0
    // (the rare sides of these shortcuts sit behind `asm volatile("")`: without it the compiler if-converts the selects and
    // executes the divisions anyway -- 5 of the 8 divisions per bin and step in the first capture)
    auto self_ratio = [](double x) {  // x / x: exactly 1 for a finite non-zero x
        double one = 1.0;
        if (!(x < 1.7976931348623157e308)) { asm volatile(""); one = x / x; }
        return one;
    };
    auto clean_rz = [&](double ms, double s) { return ((s > 0) && (ms > 0)) ? ms * self_ratio(s) + 0.0 : 0.0; };
    auto clean_ss = [&](double ms, double s) { return (s > 0) ? ms * self_ratio(s) + 0.0 : 0.0; };
    EulerFlux fx[5];  // evap_soil, transp, q_rz (root zone); cpr_rz, q_ss (subsoil)
    Family fams[5];
    const double flx[5] = {F(RSF_evap_soil), F(RSF_transp), F(RSF_q_rz), F(RSF_cpr_rz), F(RSF_q_ss)};
    {
        const int procs[5] = {RSP_evap_soil, RSP_transp, RSP_q_rz, RSP_cpr_rz, RSP_q_ss};
#pragma unroll
        for (int k = 0; k < 5; k++) family_begin(fams[k], SP(procs[k]), C, m);
    }
    bool any52 = false;  // family 52 reads the layer's cumulative StorAge backwards: the S passes below leave it in scratch planes
#pragma unroll
    for (int k = 0; k < 5; k++) {
        if (fams[k].sp[0] == 52) {
            any52 = true;
            fams[k].rev = cx.pl.rev[(k < 3) ? 0 : 1] + i;
            fams[k].rev_stride = C;
        }
    }
    // one logarithm per layer and bin serves every power-law flux of the layer (skipped for empty bins)
    bool logs_a = false, logs_b = false;
    double S_a = 0, rS_a = 0, S_b = 0, rS_b = 0;
    // per-flux scalars from S = max(SA) of the two layers; `half`: the RK4 stages 2 and 3 advance by h / 2 (:1433-1449, :1583-1599)
    auto prepare = [&](double mxa, double mxb, bool half) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            family_prepare(fams[k], (k < 3) ? mxa : mxb);
            fx[k].flux = flx[k];
            fx[k].fh = half ? flx[k] * h / 2 : flx[k] * h;
            fx[k].active = !(fx[k].fh <= 0);
            const Family& f = fams[k];
            fx[k].sa_only = f.fast && !f.is_uniform && !f.is_dirac && !f.is_exp52;
            fx[k].powfast = f.fast && f.is_pow && (f.S > 0) && div_range_ok(f.S) && (f.sp[1] - f.sp[1] == 0.0);
            fx[k].S = f.S;
            fx[k].y = f.sp[1];
            fx[k].rS = 1.0 / f.S;
        }
        logs_a = (fx[0].active && fx[0].powfast) || (fx[1].active && fx[1].powfast) || (fx[2].active && fx[2].powfast);
        logs_b = (fx[3].active && fx[3].powfast) || (fx[4].active && fx[4].powfast);
        S_a = fx[0].S; rS_a = fx[0].rS; S_b = fx[3].S; rS_b = fx[3].rS;
    };
    auto layer_log = [&](int layer, double SA, double sa_d) {
        EulerLog g{0.0, 0.0, false, false};
        if ((layer == 0 ? logs_a : logs_b) && !(sa_d == 0)) g = euler_log(cx, SA, layer == 0 ? S_a : S_b, layer == 0 ? rS_a : rS_b);
        return g;
    };
    auto reset = [&]() {
#pragma unroll
        for (int k = 0; k < 5; k++) fx[k].om_prev = fx[k].active ? omega(fams[k], P, cx.tabs, 0, 0.0 * m) : 0.0;
    };
    auto set_sums = [&](const double* sums) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            fx[k].ssum = sums[k];
            fx[k].ssum_ok = (sums[k] > 0) && div_range_ok(sums[k]);
            fx[k].rssum = 1.0 / sums[k];
        }
    };
    // backward travel-time distributions of bin j (calc_TT_num / calc_tt_num, transport.py:859-945) from the masked cumulative StorAge
    // at the bin's upper edge (na, nb) and its increment over the bin (da, db)
    auto tt_of_bin = [&](int j, double na, double nb, double da, double db, double* tt) {
        const EulerLog ga = layer_log(0, na, da), gb = layer_log(1, nb, db);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const double q = euler_q(fx[k], fams[k], cx, j, (k < 3) ? na : nb, (k < 3) ? da : db, (k < 3) ? ga : gb);
            const double p = (q > 0) ? ((fx[k].ssum_ok && div_range_ok(q)) ? div_by(q, fx[k].ssum, fx[k].rssum) : q / fx[k].ssum) : 0;
            fx[k].accT = (j == 0) ? p : fx[k].accT + p;  // np.cumsum starts from the first element
            const double d = fx[k].accT - fx[k].Tprev;
            fx[k].Tprev = fx[k].accT;
            tt[k] = (d >= 0) ? d : 0;
        }
    };
    if (!RK4) {
        // P1: S of both layers (calc_SA :343-359 times maskCatch, np.max over nages points)
        NpMax mx_rz, mx_ss;
        mx_rz.init(0.0 * m);
        mx_ss.init(0.0 * m);
        {
            double acc_a = 0, acc_b = 0;
            Cursor c(P, i);
            for (int j = 0; j < A; j++, c.next()) {
                acc_a += s_rz[c.off];
                acc_b += s_ss[c.off];
                mx_rz.push(acc_a * m);
                mx_ss.push(acc_b * m);
                if (any52) {
                    cx.pl.rev[0][cx.lin(j)] = acc_a * m;
                    cx.pl.rev[1][cx.lin(j)] = acc_b * m;
                }
            }
        }
        prepare(mx_rz.get(), mx_ss.get(), false);
    }
    // P2: normalisation constants
    if (!RK4) {
        reset();
        double acc_a = 0, acc_b = 0, SAa = 0.0 * m, SAb = 0.0 * m;
        double sums[5];
        EULER_SUMS<5, 2>(cx.plan, EULER_SUMS_ARG
            [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = s_rz[o]; L[1] = s_ss[o]; },
            [&](int j, const double* L, double* v) {
                acc_a += L[0];
                acc_b += L[1];
                const double na = acc_a * m, nb = acc_b * m;
                const double da = na - SAa, db = nb - SAb;
                SAa = na;
                SAb = nb;
                const EulerLog ga = layer_log(0, na, da), gb = layer_log(1, nb, db);
#pragma unroll
                for (int k = 0; k < 5; k++) v[k] = euler_q(fx[k], fams[k], cx, j, (k < 3) ? na : nb, (k < 3) ? da : db, (k < 3) ? ga : gb);
            }, sums);
        set_sums(sums);
    }
    // Runge-Kutta solver (settings.sas_solver = "RK4", svat_transport_model_rk4 transport.py:1138-2047): four stages of the Euler
    // building blocks on stage copies of the StorAge (scratch planes, bin-major like TTn); stage 1 advances by h, stages 2 and 3 by
    // h / 2 with mass-like bookkeeping of the stage concentrations, stage 3 tests sarkn - dsarkn < 0, dsarkn_ss1 of stage 1 lacks
    // the flux -- all as written in the reference.  k1 + 2 k2 + 2 k3 + k4 accumulates in the reference's order in five more scratch
    // planes; P3 below divides by 6 and performs the Euler solver's update with the combined distributions (:1836-1945).
    if (RK4) {
        double* __restrict__ sr = cx.pl.san;
        double* __restrict__ ss = cx.pl.TTn;
        double* __restrict__ mr = cx.pl.om;
        double* __restrict__ mss = cx.pl.rk[0];
        double* __restrict__ k0 = cx.pl.rk[1];
        double* __restrict__ k1 = cx.pl.rk[2];
        double* __restrict__ k2 = cx.pl.rk[3];
        double* __restrict__ k3 = cx.pl.rk[4];
        double* __restrict__ k4 = cx.pl.rk[5];
        for (int j = 0; j < A; j++) {  // sarkn = sa, msarkn = msa after the boundary condition (:1224-1240)
            const int64_t o = cx.ring(j), l = cx.lin(j);
            const double a = s_rz[o], b = s_ss[o];
            sr[l] = a;
            ss[l] = b;
            mr[l] = (j == 0) ? m_rz[o] : clean_rz(m_rz[o], a);
            mss[l] = (j == 0) ? m_ss[o] : clean_ss(m_ss[o], b);
        }
        const double evap = flx[0], transp = flx[1], q_rz = flx[2], cpr = flx[3], q_ss = flx[4];
        for (int st = 0; st < 4; st++) {
            const bool half = (st == 1 || st == 2);
            // cumulative StorAge of the stage: calc_SA * maskCatch for the first stage, SArkn[1:] = cumsum(sarkn) afterwards
            double acc_a = 0, acc_b = 0, SAa = 0.0 * m, SAb = 0.0 * m;
            auto advance = [&](int j, double a, double b, double& na, double& nb) {
                if (st == 0) {
                    acc_a += a;
                    acc_b += b;
                    na = acc_a * m;
                    nb = acc_b * m;
                } else {
                    acc_a = (j == 0) ? a : acc_a + a;
                    acc_b = (j == 0) ? b : acc_b + b;
                    na = acc_a;
                    nb = acc_b;
                }
            };
            {
                NpMax mx_rz, mx_ss;
                mx_rz.init(0.0 * m);
                mx_ss.init(0.0 * m);
                for (int j = 0; j < A; j++) {
                    const int64_t l = cx.lin(j);
                    double na, nb;
                    advance(j, sr[l], ss[l], na, nb);
                    mx_rz.push(na);
                    mx_ss.push(nb);
                    if (any52) {
                        cx.pl.rev[0][l] = na;
                        cx.pl.rev[1][l] = nb;
                    }
                }
                prepare(mx_rz.get(), mx_ss.get(), half);
            }
            {   // normalisation constants of the stage
                reset();
                acc_a = 0; acc_b = 0;
                double sums[5];
                EULER_SUMS<5, 2>(cx.plan, EULER_SUMS_ARG
                    [&](int j, double* L) { const int64_t l = cx.lin(j); L[0] = sr[l]; L[1] = ss[l]; },
                    [&](int j, const double* L, double* v) {
                        double na, nb;
                        advance(j, L[0], L[1], na, nb);
                        const double da = na - SAa, db = nb - SAb;
                        SAa = na;
                        SAb = nb;
                        const EulerLog ga = layer_log(0, na, da), gb = layer_log(1, nb, db);
#pragma unroll
                        for (int k = 0; k < 5; k++) v[k] = euler_q(fx[k], fams[k], cx, j, (k < 3) ? na : nb, (k < 3) ? da : db, (k < 3) ? ga : gb);
                    }, sums);
                set_sums(sums);
            }
            reset();
#pragma unroll
            for (int k = 0; k < 5; k++) { fx[k].accT = 0; fx[k].Tprev = 0; }
            acc_a = 0; acc_b = 0; SAa = 0.0 * m; SAb = 0.0 * m;
            for (int j = 0; j < A; j++) {
                const int64_t l = cx.lin(j);
                double a = sr[l], b = ss[l], ma = mr[l], mb = mss[l];
                double na, nb;
                advance(j, a, b, na, nb);
                const double da = na - SAa, db = nb - SAb;
                SAa = na;
                SAb = nb;
                double tt[5];
                tt_of_bin(j, na, nb, da, db, tt);
                const double tt_e = tt[0], tt_t = tt[1], tt_q = tt[2], tt_c = tt[3], tt_s = tt[4];
                if (st == 0) {
                    k0[l] = tt_e; k1[l] = tt_t; k2[l] = tt_q; k3[l] = tt_c; k4[l] = tt_s;
                } else {
                    const double w = half ? 2. : 1.;
                    k0[l] = k0[l] + w * tt_e; k1[l] = k1[l] + w * tt_t; k2[l] = k2[l] + w * tt_q; k3[l] = k3[l] + w * tt_c;
                    k4[l] = k4[l] + w * tt_s;
                }
                // solute travel-time distributions of the stage: mtt = where(tt > 0, msarkn, 0) (calc_mtt, isotope branch :575-580)
                const double mtt_e = (tt_e > 0) ? ma : 0, mtt_t = (tt_t > 0) ? ma : 0, mtt_q = (tt_q > 0) ? ma : 0;
                const double mtt_c = (tt_c > 0) ? mb : 0, mtt_s = (tt_s > 0) ? mb : 0;
                if (st == 0) {  // :1388-1431
                    double dsa_rz = (cpr * tt_c - evap * tt_e - transp * tt_t - q_rz * tt_q) * h;
                    dsa_rz = (a + dsa_rz < 0) ? -a : dsa_rz;
                    double dsa_ss = (q_rz * tt_q - cpr * tt_c - q_ss * tt_s) * h;
                    dsa_ss = (b + dsa_ss < 0) ? -b : dsa_ss;
                    const double dsa_rz1 = (cpr * tt_c) * h;
                    const double dmsa_rz1 = nan0(mtt_c) * ((dsa_rz1 > 0) ? ((cpr * tt_c * h) / dsa_rz1) : 0);
                    const double dsa_ss1 = (tt_q) * h;
                    const double dmsa_ss1 = nan0(mtt_q) * ((dsa_ss1 > 0) ? ((q_rz * tt_q * h) / dsa_ss1) : 0);
                    ma = euler_mix(ma, a, dsa_rz1, dmsa_rz1, true);
                    mb = euler_mix(mb, b, dsa_ss1, dmsa_ss1, false);
                    a += dsa_rz;
                    b += dsa_ss;
                    ma = (a <= 0) ? 0 : ma;
                    mb = (b <= 0) ? 0 : mb;
                } else if (half) {  // :1551-1581, :1699-1729
                    double dsa_rz = (cpr * tt_c - evap * tt_e - transp * tt_t - q_rz * tt_q) * h / 2;
                    double dsa_ss = (q_rz * tt_q - cpr * tt_c - q_ss * tt_s) * h / 2;
                    if (st == 1) {
                        dsa_rz = (a + dsa_rz < 0) ? -a : dsa_rz;
                        dsa_ss = (b + dsa_ss < 0) ? -b : dsa_ss;
                    } else {
                        dsa_rz = (a - dsa_rz < 0) ? -a : dsa_rz;
                        dsa_ss = (b - dsa_ss < 0) ? -b : dsa_ss;
                    }
                    a += dsa_rz;
                    b += dsa_ss;
                    const double r1 = nan0(mtt_c);
                    const double r2 = nan0(mtt_e) + nan0(mtt_t) + nan0(mtt_q);
                    double s1 = nan0(mtt_q);
                    double s2 = nan0(mtt_c) + nan0(mtt_s);
                    if (st == 1) { s1 = s1 * h / 2; s2 = s2 * h / 2; }
                    double dr = ((r1 < 0) && (r2 >= 0)) ? r1 + r2 : r1 - r2;
                    double ds = ((s1 < 0) && (s2 >= 0)) ? s1 + s2 : s1 - s2;
                    dr = (dr > 0) ? ((ma + dr > 0) ? -ma : dr) : dr;
                    ds = (ds > 0) ? ((mb + ds > 0) ? -mb : ds) : ds;
                    ma += dr;
                    mb += ds;
                }
                if (st < 3) {
                    sr[l] = a;
                    ss[l] = b;
                    mr[l] = ma;
                    mss[l] = mb;
                }
            }
        }
    }
    // P3: travel-time distributions, mixing, update
    {
        reset();
#pragma unroll
        for (int k = 0; k < 5; k++) { fx[k].accT = 0; fx[k].Tprev = 0; }
        double acc_a = 0, acc_b = 0, SAa = 0.0 * m, SAb = 0.0 * m;
        const double evap = fx[0].flux, transp = fx[1].flux, q_rz = fx[2].flux, cpr = fx[3].flux, q_ss = fx[4].flux;
        double* __restrict__ tt_base = (LAST && P.store_tt && cx.pl.tt) ? cx.pl.tt : nullptr;
        PctTracker trk[(LAST && STATS) ? 10 : 1];
        double cdf_t = 0, cdf_s = 0;
        NpMax cmax_t, cmax_s;
        if (LAST && STATS) {
            const double pcts[5] = {0.1, 0.25, 0.5, 0.75, 0.9};
#pragma unroll
            for (int k = 0; k < 5; k++) { trk[k].init(pcts[k]); trk[5 + k].init(pcts[k]); }
        }
        auto bin = [&](int j, const double* L, double* v) {
            const int64_t o = cx.ring(j);
            const double a = L[0], b = L[1];
            const double ma = (ANION || j == 0) ? L[2] : clean_rz(L[2], a);
            const double mb = (ANION || j == 0) ? L[3] : clean_ss(L[3], b);
            double tt[5];
            if (RK4) {  // tt = (k1 + 2 k2 + 2 k3 + k4) / 6 (:1836-1856)
#pragma unroll
                for (int k = 0; k < 5; k++) tt[k] = L[4 + k] / 6.;
            } else {
                acc_a += a;
                acc_b += b;
                const double na = acc_a * m, nb = acc_b * m;
                const double da = na - SAa, db = nb - SAb;
                SAa = na;
                SAb = nb;
                tt_of_bin(j, na, nb, da, db, tt);
            }
            const double tt_e = tt[0], tt_t = tt[1], tt_q = tt[2], tt_c = tt[3], tt_s = tt[4];
            if (ANION) {
                auto amtt = [](double sa_, double msa_, double alpha, double t, double fh) {  // calc_mtt, transport.py:582-596
                    double w = ((sa_ > 0) ? (msa_ / sa_) : 0) * alpha * t * fh;
                    w = (w <= 0) ? 0 : w;
                    return (w > msa_) ? msa_ : w;
                };
                const double mt = amtt(a, ma, alpha_t, tt_t, fx[1].fh), mq = amtt(a, ma, alpha_q, tt_q, fx[2].fh);
                const double mc = amtt(b, mb, alpha_q, tt_c, fx[3].fh), msq = amtt(b, mb, alpha_q, tt_s, fx[4].fh);
                if (LAST) {
                    v[0] = mt; v[1] = mq; v[2] = mc; v[3] = msq;
                    if (STATS) {  // water-age statistics of the day's last step (transport.py:59-152), numerics.py:437-439
                        v[4] = evap * tt_e; v[5] = transp * tt_t; v[6] = q_ss * tt_s;
                        v[STATS ? 7 : 0] = tt_t; v[STATS ? 8 : 0] = tt_s;
                        v[STATS ? 9 : 0] = (double)(j + 1) * tt_t; v[STATS ? 10 : 0] = (double)(j + 1) * tt_s;
                        cdf_t += tt_t;
                        cdf_s += tt_s;
                        if (j == 0) { cmax_t.init(cdf_t); cmax_s.init(cdf_s); } else { cmax_t.push(cdf_t); cmax_s.push(cdf_s); }
#pragma unroll
                        for (int k = 0; k < 5; k++) { trk[k].push(j, cdf_t); trk[(STATS ? 5 : 0) + k].push(j, cdf_s); }
                    }
                    if (tt_base) {
                        const int64_t span = (int64_t)A * C, l = cx.lin(j);
                        tt_base[RSP_evap_soil * span + l] = tt_e;
                        tt_base[RSP_transp * span + l] = tt_t;
                        tt_base[RSP_q_rz * span + l] = tt_q;
                        tt_base[RSP_cpr_rz * span + l] = tt_c;
                        tt_base[RSP_q_ss * span + l] = tt_s;
                    }
                }
                double dsa_rz = (cpr * tt_c - evap * tt_e - transp * tt_t - q_rz * tt_q) * h;
                dsa_rz = (a + dsa_rz < 0) ? -a : dsa_rz;
                double dsa_ss = (q_rz * tt_q - cpr * tt_c - q_ss * tt_s) * h;
                dsa_ss = (b + dsa_ss < 0) ? -b : dsa_ss;
                double a2 = a + dsa_rz, b2 = b + dsa_ss;
                double dma = nan0(mc) - nan0(mt) - nan0(mq);
                double dmb = nan0(mq) - nan0(mc) - nan0(msq);
                dma = (ma + dma < 0) ? 0 : dma;
                dmb = (mb + dmb < 0) ? 0 : dmb;
                double ma2 = ma + dma, mb2 = mb + dmb;
                a2 = (a2 < 1e-8) ? 0 : a2;   // anion storage kernels (root_zone.py:219-240, subsoil.py), run after every Euler step
                b2 = (b2 < 1e-8) ? 0 : b2;
                ma2 = (a2 <= 0) ? 0 : ma2;
                mb2 = (b2 <= 0) ? 0 : mb2;
                s_rz[o] = a2;
                s_ss[o] = b2;
                m_rz[o] = ma2;
                m_ss[o] = mb2;
                return;
            }
            const double mtt_c = (tt_c > 0) ? mb : 0, mtt_q = (tt_q > 0) ? ma : 0;
            if (LAST) {
                v[0] = tt_e; v[1] = ((tt_e > 0) ? ma : 0) * tt_e;
                v[2] = tt_t; v[3] = ((tt_t > 0) ? ma : 0) * tt_t;
                v[4] = tt_q; v[5] = mtt_q * tt_q;
                v[6] = tt_c; v[7] = mtt_c * tt_c;
                v[8] = tt_s; v[9] = ((tt_s > 0) ? mb : 0) * tt_s;
                if (STATS) {  // transport.py:59-152 (TT[1:] = cumsum(tt)), numerics.py:437-439
                    v[10] = evap * tt_e; v[11] = transp * tt_t; v[12] = q_ss * tt_s;
                    v[13] = (double)(j + 1) * tt_t; v[14] = (double)(j + 1) * tt_s;
                    cdf_t += tt_t;
                    cdf_s += tt_s;
                    if (j == 0) { cmax_t.init(cdf_t); cmax_s.init(cdf_s); } else { cmax_t.push(cdf_t); cmax_s.push(cdf_s); }
#pragma unroll
                    for (int k = 0; k < 5; k++) { trk[k].push(j, cdf_t); trk[5 + k].push(j, cdf_s); }
                }
                if (tt_base) {
                    const int64_t span = (int64_t)A * C, l = cx.lin(j);
                    tt_base[RSP_evap_soil * span + l] = tt_e;
                    tt_base[RSP_transp * span + l] = tt_t;
                    tt_base[RSP_q_rz * span + l] = tt_q;
                    tt_base[RSP_cpr_rz * span + l] = tt_c;
                    tt_base[RSP_q_ss * span + l] = tt_s;
                }
            }
            double dsa_rz = (cpr * tt_c - evap * tt_e - transp * tt_t - q_rz * tt_q) * h;
            dsa_rz = (a + dsa_rz < 0) ? -a : dsa_rz;
            double dsa_ss = (q_rz * tt_q - cpr * tt_c - q_ss * tt_s) * h;
            dsa_ss = (b + dsa_ss < 0) ? -b : dsa_ss;
            const double dsa_rz1 = (cpr * tt_c) * h;
            // (cpr * tt_c * h) / dsa_rz1 divides a value by itself: exactly 1 for a finite positive value
            const double dmsa_rz1 = nan0(mtt_c) * ((dsa_rz1 > 0) ? self_ratio(dsa_rz1) : 0);
            const double dsa_ss1 = (q_rz * tt_q) * h;
            const double dmsa_ss1 = nan0(mtt_q) * ((dsa_ss1 > 0) ? self_ratio(dsa_ss1) : 0);
            // a bin that receives nothing keeps its concentration: the general expression reduces to clean_*()
            double ma2 = ma, mb2 = mb;
            if (!(dsa_rz1 == 0 && dmsa_rz1 == 0 && j > 0)) { asm volatile(""); ma2 = euler_mix(ma, a, dsa_rz1, dmsa_rz1, true); }
            if (!(dsa_ss1 == 0 && dmsa_ss1 == 0 && j > 0)) { asm volatile(""); mb2 = euler_mix(mb, b, dsa_ss1, dmsa_ss1, false); }
            double a2 = a + dsa_rz, b2 = b + dsa_ss;
            ma2 = (a2 <= 0) ? 0 : ma2;
            mb2 = (b2 <= 0) ? 0 : mb2;
            a2 = (a2 < 1e-8) ? 0 : a2;  // root_zone.py:196-199, subsoil.py:160-163 (run after every Euler step)
            b2 = (b2 < 1e-8) ? 0 : b2;
            s_rz[o] = a2;
            s_ss[o] = b2;
            m_rz[o] = ma2;
            m_ss[o] = mb2;
        };
        constexpr int NL = RK4 ? 9 : 4;
        auto load = [&](int j, double* L) {
            const int64_t o = cx.ring(j);
            L[0] = s_rz[o]; L[1] = s_ss[o]; L[2] = m_rz[o]; L[3] = m_ss[o];
            if (RK4) {
                const int64_t l = cx.lin(j);
#pragma unroll
                for (int k = 0; k < 5; k++) L[(RK4 ? 4 : 0) + k] = cx.pl.rk[1 + k][l];
            }
        };
        if (LAST) {
            constexpr int NS = ANION ? (STATS ? 11 : 4) : (STATS ? 15 : 10);
            double sums[NS];
            EULER_SUMS<NS, NL>(cx.plan, EULER_SUMS_ARG load, bin, sums);
            if (ANION && STATS && es) {
                const double qn = __longlong_as_double(0x7ff8000000000000LL);
                es->wsum[0] = sums[STATS ? 4 : 0]; es->wsum[1] = sums[STATS ? 5 : 0]; es->wsum[2] = sums[STATS ? 6 : 0];
#pragma unroll
                for (int k = 0; k < 5; k++) { es->pct[0][k] = trk[k].result(A, cmax_t.get()); es->pct[1][k] = trk[(STATS ? 5 : 0) + k].result(A, cmax_s.get()); }
                es->avg[0] = (sums[STATS ? 7 : 0] > 0) ? sums[STATS ? 9 : 0] : qn;
                es->avg[1] = (sums[STATS ? 8 : 0] > 0) ? sums[STATS ? 10 : 0] : qn;
            }
            if (ANION) {  // :2378-2407; C_evap_soil and the M_* flux planes are not touched by the Euler solver
                O(RSO_C_inf_mat_rz) = ((inf_mat * h > 0) ? C_in : 0) * m;
                O(RSO_C_inf_pf_rz) = ((inf_pf * h > 0) ? C_in : 0) * m;
                O(RSO_C_inf_pf_ss) = ((inf_ss * h > 0) ? C_in : 0) * m;
                O(RSO_C_transp) = ((transp > 0) ? sums[0] / fx[1].fh : 0) * m;
                O(RSO_C_q_rz) = ((q_rz > 0) ? sums[1] / fx[2].fh : 0) * m;
                O(RSO_C_cpr_rz) = ((cpr > 0) ? sums[2] / fx[3].fh : 0) * m;
                O(RSO_C_q_ss) = ((q_ss > 0) ? sums[3] / fx[4].fh : 0) * m;
                return;
            }
            if (STATS && es) {
                const double qn = __longlong_as_double(0x7ff8000000000000LL);
                es->wsum[0] = sums[NS - 5]; es->wsum[1] = sums[NS - 4]; es->wsum[2] = sums[NS - 3];
#pragma unroll
                for (int k = 0; k < 5; k++) { es->pct[0][k] = trk[k].result(A, cmax_t.get()); es->pct[1][k] = trk[(STATS ? 5 : 0) + k].result(A, cmax_s.get()); }
                es->avg[0] = (sums[2] > 0) ? sums[NS - 2] : qn;
                es->avg[1] = (sums[8] > 0) ? sums[NS - 1] : qn;
            }
            // calc_conc_iso_flux :512-535, delta_fluxes_svat :3659-3697
            const int oc[5] = {RSO_C_evap_soil, RSO_C_transp, RSO_C_q_rz, RSO_C_cpr_rz, RSO_C_q_ss};
            const int oi[5] = {RSO_C_iso_evap_soil, RSO_C_iso_transp, RSO_C_iso_q_rz, RSO_C_iso_cpr_rz, RSO_C_iso_q_ss};
#pragma unroll
            for (int k = 0; k < 5; k++) {
                double conc = (sums[2 * k] > 0) ? sums[2 * k + 1] / sums[2 * k] : qnan;
                conc = (conc != 0) ? conc : qnan;
                conc = conc * m;
                O(oc[k]) = conc;
                O(oi[k]) = conc_to_delta(P, conc) * m;
            }
            const double c1 = ((inf_mat > 0) ? C_in : qnan) * m, c2 = ((inf_pf > 0) ? C_in : qnan) * m, c3 = ((inf_ss > 0) ? C_in : qnan) * m;
            O(RSO_C_inf_mat_rz) = c1; O(RSO_C_iso_inf_mat_rz) = conc_to_delta(P, c1) * m;
            O(RSO_C_inf_pf_rz) = c2; O(RSO_C_iso_inf_pf_rz) = conc_to_delta(P, c2) * m;
            O(RSO_C_inf_pf_ss) = c3; O(RSO_C_iso_inf_pf_ss) = conc_to_delta(P, c3) * m;
        } else {
            double Lc[NL], Ln[NL], v[1];
            load(0, Lc);
            for (int j = 0; j < A; j++) {  // next bin's loads in flight while this bin computes
                if (j + 1 < A) load(j + 1, Ln);
                bin(j, Lc, v);
#pragma unroll
                for (int q = 0; q < NL; q++) Lc[q] = Ln[q];
            }
        }
    }
#undef F
#undef O
#undef SP
}

// anion storage kernels (root_zone.py:219-261, subsoil.py / soil.py anion kernels) and ageing (transport.py:622-679) as used
// by anion_day<>, for the Euler solver
__device__ __forceinline__ void anion_storage(const CellCtx& cx, double m, double* __restrict__ out, double* sum_sa_s = nullptr) {
    const int64_t C = cx.P.C, i = cx.i;
    double* __restrict__ s_rz = cx.pl.sa[0];
    double* __restrict__ s_ss = cx.pl.sa[1];
    double* __restrict__ m_rz = cx.pl.msa[0];
    double* __restrict__ m_ss = cx.pl.msa[1];
    double sums[6];
    stream_sums<6, 4>(cx.plan,
        [&](int j, double* L) {
            const int64_t o = cx.ring(j);
            L[0] = s_rz[o]; L[1] = s_ss[o]; L[2] = m_rz[o]; L[3] = m_ss[o];
        },
        [&](int j, const double* L, double* v) {
            const int64_t o = cx.ring(j);
            double a = L[0];
            a = (a < 1e-8) ? 0 : a;
            double b = L[1];
            b = (b < 1e-8) ? 0 : b;
            const double ma = (a <= 0) ? 0 : L[2], mb = (b <= 0) ? 0 : L[3];
            s_rz[o] = a; s_ss[o] = b; m_rz[o] = ma; m_ss[o] = mb;
            const double ms = ma + mb * m;
            v[0] = isnan_d(ma) ? 0 : ma;
            v[1] = a;
            v[2] = isnan_d(mb) ? 0 : mb;
            v[3] = b;
            v[4] = isnan_d(ms) ? 0 : ms;
            v[5] = a + b * m;
        }, sums);
#define O(k) out[(int64_t)(k) * C + i]
    const double M_rz = sums[0] * m, M_ss = sums[2] * m, M_s = sums[4] * m;
    O(RSO_C_iso_rz) = M_rz; O(RSO_C_rz) = (sums[1] > 0) ? M_rz / sums[1] : 0;
    O(RSO_C_iso_ss) = M_ss; O(RSO_C_ss) = (sums[3] > 0) ? M_ss / sums[3] : 0;
    O(RSO_C_iso_s) = M_s;   O(RSO_C_s) = (sums[5] > 0) ? M_s / sums[5] : 0;
    if (sum_sa_s) *sum_sa_s = sums[5];
#undef O
}
__device__ __forceinline__ void anion_ageing(const CellCtx& cx) {
    const SasParams& P = cx.P;
    const int A = P.A;
    const int64_t C = P.C, i = cx.i;
    const int b_new = (P.base == 0) ? A - 1 : P.base - 1;
    const int b_last = (b_new == 0) ? A - 1 : b_new - 1;
    for (int L = 0; L < 2; L++) {
        double* __restrict__ sa = cx.pl.sa[L];
        double* __restrict__ msa = cx.pl.msa[L];
        double s_new = sa[(int64_t)b_last * C + i] + sa[(int64_t)b_new * C + i];
        s_new = (s_new < 1e-8) ? 0 : s_new;
        sa[(int64_t)b_last * C + i] = s_new;
        msa[(int64_t)b_last * C + i] = msa[(int64_t)b_last * C + i] + msa[(int64_t)b_new * C + i];
        sa[(int64_t)b_new * C + i] = 0;
        msa[(int64_t)b_new * C + i] = 0;
    }
}

template <bool STATS>
__device__ __forceinline__ void euler_day_anion(const CellCtx& cx, const double* __restrict__ forcing, double m, const double* __restrict__ sas_params,
                                                const AnionParams ap, double* __restrict__ out, double* __restrict__ stats,
                                                double* __restrict__ C_s_prev) {
    const double alpha_t = ap.alpha_transp ? ap.alpha_transp[cx.i] : 1.0, alpha_q = ap.alpha_q ? ap.alpha_q[cx.i] : 1.0;
    const int n = cx.P.substeps;
    const int64_t C = cx.P.C, i = cx.i;
    for (int s = 0; s + 1 < n; s++) euler_step<false, false, true>(cx, forcing, m, sas_params, out, nullptr, alpha_t, alpha_q);
    double sum_sa_s_m1 = 0, Cp = 0;
    if (STATS) {  // [taum1] of the last Euler step: what its predecessor left (after_substep_anion)
        if (n > 1) {  // the storage kernels after that step (idempotent on the state: the step applied their thresholds itself)
            anion_storage(cx, m, out, &sum_sa_s_m1);
            Cp = out[(int64_t)RSO_C_s * C + i];
        } else {
            Cp = C_s_prev[i];
            double sm[1];
            stream_sums<1, 2>(cx.plan,
                [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = cx.pl.sa[0][o]; L[1] = cx.pl.sa[1][o]; },
                [&](int, const double* L, double* v) { v[0] = L[0] + L[1] * m; }, sm);
            sum_sa_s_m1 = sm[0];
        }
    }
    EulerStats es;
    euler_step<true, STATS, true>(cx, forcing, m, sas_params, out, &es, alpha_t, alpha_q);
    double sum_sa_s = 0;
    anion_storage(cx, m, out, &sum_sa_s);
    if (STATS) {  // age statistics of the water (transport.py:59-312), balance errors numerics.py:424-445 and :642-672 with settings.h
#define F(k) forcing[(int64_t)(k) * C + i]
#define O(k) out[(int64_t)(k) * C + i]
#define ST(k) stats[(int64_t)(k) * C + i]
        for (int k = 0; k < 5; k++) { ST(RSS_tt10_transp + k) = es.pct[0][k]; ST(RSS_tt10_q_ss + k) = es.pct[1][k]; }
        ST(RSS_ttavg_transp) = es.avg[0];
        ST(RSS_ttavg_q_ss) = es.avg[1];
        const double p5[5] = {0.1, 0.25, 0.5, 0.75, 0.9}, p3[3] = {0.25, 0.5, 0.75};
        double pc[5], avg;
        storage_age_statistics(cx, 0, m, 3, p3, pc, &avg);
        for (int k = 0; k < 3; k++) ST(RSS_rt25_rz + k) = pc[k];
        ST(RSS_rtavg_rz) = avg;
        storage_age_statistics(cx, 1, m, 3, p3, pc, &avg);
        for (int k = 0; k < 3; k++) ST(RSS_rt25_ss + k) = pc[k];
        ST(RSS_rtavg_ss) = avg;
        storage_age_statistics(cx, 2, m, 5, p5, pc, &avg);
        for (int k = 0; k < 5; k++) ST(RSS_rt10_s + k) = pc[k];
        ST(RSS_rtavg_s) = avg;
        auto nan0 = [](double x) { return (x != x) ? 0.0 : x; };
        const double h = 1.0 / (double)n;
        const double C_s = O(RSO_C_s);
        ST(RSS_dS_num_error) = fabs((sum_sa_s - sum_sa_s_m1) - (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - es.wsum[0] - es.wsum[1] - es.wsum[2]) * h);
        ST(RSS_dC_num_error) = fabs((sum_sa_s * C_s - sum_sa_s_m1 * Cp) -
                                    (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                     F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - es.wsum[1] * nan0(O(RSO_C_transp)) -
                                     es.wsum[2] * nan0(O(RSO_C_q_ss))) * h);
        ST(RSS_sanity_ok) = __longlong_as_double(0x7ff8000000000000LL);  // the anion branch of numerics.sanity_check is not restated
        C_s_prev[i] = C_s;
#undef F
#undef O
#undef ST
    }
    anion_ageing(cx);
}

template <bool STATS, bool RK4 = false>
__device__ __forceinline__ void euler_day(const CellCtx& cx, const double* __restrict__ forcing, double m, const double* __restrict__ sas_params,
                                          double* __restrict__ out, double* __restrict__ stats, double* __restrict__ C_s_prev,
                                          const double* __restrict__ bounds, double atol, double rtol) {
    const SasParams& P = cx.P;
    const int64_t C = P.C, i = cx.i;
    const int n = P.substeps;
    auto step = [&](auto last, EulerStats* es) {
        constexpr bool L = decltype(last)::value;
        if (RK4) rk4_step<L, L && STATS, false, true>(cx, forcing, m, sas_params, out, es, 1.0, 1.0);
        else euler_step<L, L && STATS, false>(cx, forcing, m, sas_params, out, es, 1.0, 1.0);
    };
    for (int s = 0; s + 2 < n; s++) step(std::false_type{}, nullptr);
    if (n > 1) step(std::false_type{}, nullptr);
    double sum_sa_s, C_s, sum_sa_rz, sum_sa_ss;
    double sum_sa_s_m1 = 0, Cp = 0;
    if (STATS) {  // [taum1] of the last Euler step: what its predecessor left (after_substep_iso, transport.py:3421-3520)
        if (n > 1) {
            storage_concentrations(cx, m, out, C, &sum_sa_s, &C_s, &sum_sa_rz, &sum_sa_ss);
            Cp = C_s;
        } else Cp = C_s_prev[i];
        double sm[1];
        stream_sums<1, 2>(cx.plan,
            [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = cx.pl.sa[0][o]; L[1] = cx.pl.sa[1][o]; },
            [&](int, const double* L, double* v) { v[0] = L[0] + L[1] * m; }, sm);
        sum_sa_s_m1 = sm[0];
    }
    EulerStats es;
    step(std::true_type{}, &es);
    storage_concentrations(cx, m, out, C, &sum_sa_s, &C_s, &sum_sa_rz, &sum_sa_ss);
    if (STATS) {
#define F(k) forcing[(int64_t)(k) * C + i]
#define O(k) out[(int64_t)(k) * C + i]
#define ST(k) stats[(int64_t)(k) * C + i]
        for (int k = 0; k < 5; k++) { ST(RSS_tt10_transp + k) = es.pct[0][k]; ST(RSS_tt10_q_ss + k) = es.pct[1][k]; }
        ST(RSS_ttavg_transp) = es.avg[0];
        ST(RSS_ttavg_q_ss) = es.avg[1];
        const double p5[5] = {0.1, 0.25, 0.5, 0.75, 0.9}, p3[3] = {0.25, 0.5, 0.75};
        double pc[5], avg;
        storage_age_statistics(cx, 0, m, 3, p3, pc, &avg);
        for (int k = 0; k < 3; k++) ST(RSS_rt25_rz + k) = pc[k];
        ST(RSS_rtavg_rz) = avg;
        storage_age_statistics(cx, 1, m, 3, p3, pc, &avg);
        for (int k = 0; k < 3; k++) ST(RSS_rt25_ss + k) = pc[k];
        ST(RSS_rtavg_ss) = avg;
        storage_age_statistics(cx, 2, m, 5, p5, pc, &avg);
        for (int k = 0; k < 5; k++) ST(RSS_rt10_s + k) = pc[k];
        ST(RSS_rtavg_s) = avg;
        // numerics.py:424-445, :476-516, :1051-1122: the right-hand sides carry settings.h = 1 / substeps
        auto nan0 = [](double x) { return (x != x) ? 0.0 : x; };
        const double h = 1.0 / (double)n;
        const double lhs1 = sum_sa_s - sum_sa_s_m1;
        const double rhs1 = (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - es.wsum[0] - es.wsum[1] - es.wsum[2]) * h;
        const double lhs2 = sum_sa_s * C_s - sum_sa_s_m1 * Cp;
        const double rhs2 = (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                             F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - es.wsum[0] * nan0(O(RSO_C_evap_soil)) -
                             es.wsum[1] * nan0(O(RSO_C_transp)) - es.wsum[2] * nan0(O(RSO_C_q_ss))) * h;
        ST(RSS_dS_num_error) = fabs(lhs1 - rhs1);
        ST(RSS_dC_num_error) = fabs(lhs2 - rhs2);
        auto isclose = [&](double x, double y) { return ((fabs(x - y) <= atol + rtol * fabs(y)) && isfinite(y)) || (x == y); };
        const bool c1 = (m != 0) ? isclose(lhs1, rhs1) : true;
        const bool c2 = (m != 0) ? isclose(lhs2, rhs2) : true;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        const double sat_rz = bounds ? bounds[i] : inf, sat_ss = bounds ? bounds[C + i] : inf;
        const double pwp_rz = bounds ? bounds[2 * C + i] : 0.0, pwp_ss = bounds ? bounds[3 * C + i] : 0.0;
        const bool b3 = (sum_sa_s <= (sat_rz + sat_ss) - (pwp_rz + pwp_ss)) && (sum_sa_s >= 0);
        const bool b4 = (sum_sa_rz <= sat_rz - pwp_rz) && (sum_sa_rz >= 0);
        const bool b5 = (sum_sa_ss <= sat_ss - pwp_ss) && (sum_sa_ss >= 0);
        ST(RSS_sanity_ok) = (c1 && c2 && b3 && b4 && b5) ? 1.0 : 0.0;
        C_s_prev[i] = C_s;
#undef F
#undef O
#undef ST
    }
    ageing(cx);
}

}  // namespace rsas
