/* sas_oracle.c -- CPU restatement of the reference's deterministic oxygen-18 StorAge-selection transport
 * step (roger/core/transport.py:948-991 and the *_transport_iso_kernel of every process module).
 *
 * TEST INFRASTRUCTURE (see svat_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's CPU legs use it.
 * Parity status: PINNED against outputs of the reference's own NumPy backend (tests/golden/sas_*.npz produced
 * by tests/golden/make_sas_golden.py).  SAS families restated: uniform (1), dirac (2), Kumaraswamy (3, 31-37),
 * gamma (4), exponential (51, 52), power (6, 61, 62).  The gamma family calls scipy.special.gammainc in the reference;
 * roger_gamma.h restates the published series / continued fraction it is built on (not scipy's Lanczos prefactor and
 * asymptotic branch) and is pinned by tolerance: a table of scipy outputs (tests/golden/gammainc_table.npz, 4e-15) and
 * the reference fixtures sas_gamma52 / sas_euler_gamma52 (1e-10).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/roger_b200_sas.h"
#include "../roger_b200/csrc/roger_pow.h"
#include "../roger_b200/csrc/roger_gamma.h"

/* Which pow the SAS functions call.  0: the C library's pow -- what NumPy's float64 `power` loop calls when its
 * AVX512 SVML dispatch is off; the mode the oracle is PINNED in against the reference fixtures.  1: roger_pow()
 * (roger_b200/csrc/roger_pow.h), the fixed IEEE operation sequence the CUDA kernels also evaluate, so that GPU
 * results can be compared with this oracle bit for bit.  Since round 2 roger_pow restates the C library's own algorithm
 * with the published constants: the two modes return the same bits (tests/test_pow.py), and both reproduce the fixtures. */
ROGER_POW_DEFINE_HOST_TABLES(static const)
int g_roger_oracle_pow_mode = 0; /* shared with svat_oracle.c */
#define g_pow_mode g_roger_oracle_pow_mode
void roger_oracle_set_pow(int mode) { g_pow_mode = mode; }
void sas_oracle_set_pow(int mode) { g_pow_mode = mode; }
double roger_pow_host(double x, double y)
{
    const roger_pow_tables t = {roger_pow_log_tab_h, roger_pow_exp_tab_h};
    return roger_pow(x, y, t);
}
void roger_pow2_host(double x, double y1, double y2, double *r1, double *r2)
{
    const roger_pow_tables t = {roger_pow_log_tab_h, roger_pow_exp_tab_h};
    roger_pow2(x, y1, y2, t, r1, r2);
}
void roger_pow_host_array(const double *x, const double *y, double *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) out[i] = roger_pow_host(x[i], y[i]);
}
static double sas_pow(double x, double y) { return g_pow_mode ? roger_pow_host(x, y) : pow(x, y); }
double roger_exp_host(double x)
{
    const roger_pow_tables t = {roger_pow_log_tab_h, roger_pow_exp_tab_h};
    return roger_exp(x, t);
}
static double sas_exp(double x) { return g_pow_mode ? roger_exp_host(x) : exp(x); }
/* roger_pow / roger_exp against the C library's pow / exp on n pseudo-random arguments (xorshift, four argument classes: the
 * SAS domain, tiny bases, bases next to 1, a broad class): returns the number of pow results whose bits differ,
 * *exp_mismatches that of exp.  tests/test_pow.py requires 0 / 0. */
int64_t roger_pow_compare_libm(int64_t n, uint64_t seed, int64_t *exp_mismatches)
{
    const roger_pow_tables t = {roger_pow_log_tab_h, roger_pow_exp_tab_h};
    uint64_t s = seed ? seed : 88172645463325252ULL;
    int64_t bad = 0, bade = 0;
#define RND() (s ^= s << 13, s ^= s >> 7, s ^= s << 17, (double)(s >> 11) * 0x1p-53)
    for (int64_t k = 0; k < n; k++) {
        double x, y, e;
        switch (k & 3) {
        case 0: x = RND(); y = 0.1 + 3 * RND(); e = -30 * RND(); break;
        case 1: x = exp(-40 * RND()); y = 0.05 + 2 * RND(); e = -700 * RND() * RND(); break;
        case 2: x = 1 - RND() * 1e-3; y = 5 * RND(); e = 1e-3 * (RND() - 0.5); break;
        default: x = exp(40 * (RND() - 0.5)); y = 20 * (RND() - 0.5); e = 1000 * (RND() - 0.5); break;
        }
        const double a = roger_pow(x, y, t), b = pow(x, y);
        if (fabs(y * log(x)) < 511 && memcmp(&a, &b, sizeof a) != 0) bad++;
        const double c = roger_exp(e, t), d = exp(e);
        if (fabs(e) < 511 && memcmp(&c, &d, sizeof c) != 0) bade++;
    }
#undef RND
    if (exp_mismatches) *exp_mismatches = bade;
    return bad;
}

/* numpy add.reduce over a contiguous axis (pairwise, 8 lanes, blocks of 128) */
static double pw_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return pw_sum(a, n2) + pw_sum(a + n2, n - n2);
    }
}
static double np_sum(const double *a, int n) { return 0.0 + pw_sum(a, n); }
static double np_max(const double *a, int n)
{ /* npx.max propagates NaN */
    double mx = a[0];
    for (int i = 1; i < n; i++) {
        if (isnan(a[i])) return a[i];
        if (a[i] > mx) mx = a[i];
    }
    return mx;
}

typedef struct {
    int A, substeps;
    double vsmow, dmin, dmax;
} sas_ctx;

/* transport.py:328-340 */
static double conc_to_delta(const sas_ctx *c, double conc)
{
    double d = 1000. * (conc / (c->vsmow * (1. - conc)) - 1.);
    return ((d < c->dmin) || (d > c->dmax)) ? NAN : d;
}

/* ---- SAS functions (roger/core/sas.py), Omega over nages = A + 1 points ---- */
static void sas_uniform(const sas_ctx *c, const double *SA, const double *sp, double m, double *Om)
{ /* sas.py:5-35 */
    const int N = c->A + 1;
    double mask = (sp[0] == 1) ? 1.0 : 0.0;
    double S = np_max(SA, N) * mask * m;
    double lam = 1 / S * mask * m;
    for (int j = 0; j < N; j++) Om[j] = ((SA[j] < S) ? ((SA[j] > 0) ? lam * SA[j] : 0.) : 1.) * mask * m;
    Om[N - 1] = ((mask != 0) ? 1 : Om[N - 1]) * m;
    for (int j = 0; j < N; j++) Om[j] = ((S <= 0) ? 0 : Om[j]) * m;
}
static void sas_dirac(const sas_ctx *c, const double *SA, const double *sp, double m, double *Om)
{ /* sas.py:243-264 */
    const int N = c->A + 1;
    int mask2 = (sp[0] == 2);
    double S = np_max(SA, N) * m;
    for (int j = 0; j < N; j++) Om[j] = ((mask2 && ((double)j <= sp[1])) ? 0 : 1) * m;
    for (int j = 0; j < N; j++) Om[j] = ((S <= 0) ? 0 : Om[j]) * (mask2 ? 1.0 : 0.0) * m;
}
static double s_rel(double S, const double *sp, double m)
{
    double r = (S - sp[5]) / (sp[6] - sp[5]) * m;
    r = (r < 0) ? 0 : r;
    r = (r > 1) ? 1 : r;
    return r;
}
static void sas_kumaraswami(const sas_ctx *c, const double *SA, double *sp, double m, double *Om)
{ /* sas.py:38-134 (modifies sp[1], sp[2] for the storage-dependent variants) */
    const int N = c->A + 1;
    double id = sp[0];
    int mask3 = (id == 3 || id == 31 || id == 32 || id == 33 || id == 34 || id == 35 || id == 36 || id == 37);
    double S = np_max(SA, N) * m;
    double r = s_rel(S, sp, m);
    if (id == 31) { sp[1] = 1; sp[2] = sp[3] + (r * sp[4]); }
    if (id == 32) { sp[1] = sp[3] + ((1 - r) * sp[4]); sp[2] = 1; }
    if (id == 33) { sp[1] = 1; sp[2] = sp[3] + ((1 - r) * sp[4]); }
    if (id == 34) { sp[1] = sp[3] + (r * sp[4]); sp[2] = 1; }
    if (id == 35) { sp[1] = sp[3] + ((1 - r) * sp[4]); sp[2] = sp[3] + (r * sp[4]); }
    if (id == 36) { sp[1] = sp[3] + ((1 - r) * sp[4]); }
    if (id == 37) { sp[2] = sp[3] + (r * sp[4]); }
    for (int j = 0; j < N; j++) {
        double f = 1 - sas_pow(1 - sas_pow(SA[j] / S, sp[1]), sp[2]);
        double v;
        if (S >= 0) v = (SA[j] > 0) ? ((SA[j] < S) ? f : 1.) : 0.;
        else v = (SA[j] > 0) ? f : 0.;
        Om[j] = v * (mask3 ? 1.0 : 0.0) * m;
    }
    for (int j = 0; j < N; j++) Om[j] = ((S <= 0) ? 0 : Om[j]) * m;
}
/* scipy.special.gammainc(a, x): argument rules of cephes igam, then series / continued fraction on the prefactor */
double sas_oracle_gammainc(double a, double x)
{
    const double G = exp(lgamma(a));
    if (g_pow_mode) {
        const roger_pow_tables t = {roger_pow_log_tab_h, roger_pow_exp_tab_h};
        return roger_gammainc(a, x, G, t);
    }
    if (!(x >= 0) || !(a >= 0)) return NAN;
    if (a == 0) return (x > 0) ? 1.0 : NAN;
    if (x == 0) return 0.0;
    if (isinf(a)) return isinf(x) ? NAN : 0.0;
    if (isinf(x)) return 1.0;
    return roger_gammainc_from_fac(a, x, pow(x, a) * exp(-x) / G);
}
static void sas_gamma(const sas_ctx *c, const double *SA, const double *sp, double m, double *Om)
{ /* sas.py:137-160: S carries the family mask, so every other id selects the constant 0; the quotient by
   * exp(gammaln(a)) on top of the already regularised gammainc is the reference's */
    const int N = c->A + 1;
    const double mask = (sp[0] == 4) ? 1.0 : 0.0;
    const double S = np_max(SA, N) * mask * m;
    const double G = (mask != 0) ? exp(lgamma(sp[1])) : 1.0;
    for (int j = 0; j < N; j++) {
        double v = 0.;
        if ((SA[j] > 0) && (SA[j] < S)) v = sas_oracle_gammainc(sp[1], sp[2] * SA[j] / S) / G;
        Om[j] = v * mask * m;
    }
    for (int j = 0; j < N; j++) Om[j] = ((S <= 0) ? 0 : Om[j]) * m;
}
static void sas_exponential(const sas_ctx *c, const double *SA, const double *sp, double m, double *Om)
{ /* sas.py:163-191 */
    const int N = c->A + 1;
    int m51 = (sp[0] == 51), m52 = (sp[0] == 52);
    double S = np_max(SA, N) * m;
    double *tmp = (double *)malloc(sizeof(double) * N);
    for (int j = 0; j < N; j++) tmp[j] = (SA[j] > 0) ? ((SA[j] < S) ? 1 - sas_exp(sp[1] * (-1) * (SA[j] / S)) : 1.) : 0.;
    for (int j = 0; j < N; j++) Om[j] = (m51 ? tmp[j] : 0.0) * m;
    for (int j = 0; j < N; j++) Om[j] = (m52 ? tmp[N - 1 - j] : Om[j]) * m;
    for (int j = 0; j < N; j++) Om[j] = ((S <= 0) ? 0 : Om[j]) * m;
    free(tmp);
}
static void sas_power(const sas_ctx *c, const double *SA, double *sp, double m, double *Om)
{ /* sas.py:194-240 */
    const int N = c->A + 1;
    double id = sp[0];
    int mask6 = (id == 6 || id == 61 || id == 62);
    double S = np_max(SA, N) * m;
    double r = s_rel(S, sp, m);
    if (id == 61) sp[1] = sp[3] + ((1 - r) * sp[4]);
    if (id == 62) sp[1] = sp[3] + (r * sp[4]);
    for (int j = 0; j < N; j++)
        Om[j] = ((SA[j] > 0) ? ((SA[j] <= S) ? sas_pow(SA[j] / S, sp[1]) : 1.) : 0.) * (mask6 ? 1.0 : 0.0) * m;
    for (int j = 0; j < N; j++) Om[j] = ((S <= 0) ? 0 : Om[j]) * m;
}

/* the sum of the SAS families in the reference's order (transport.py:399-428 and calc_TT_num :866-892) */
static void sas_sum_families(const sas_ctx *c, const double *SA, double *sp, double m, double *TT, double *Om)
{
    const int N = c->A + 1;
    for (int j = 0; j < N; j++) TT[j] = 0;
    sas_uniform(c, SA, sp, m, Om);
    for (int j = 0; j < N; j++) TT[j] += Om[j];
    sas_dirac(c, SA, sp, m, Om);
    for (int j = 0; j < N; j++) TT[j] += Om[j];
    sas_kumaraswami(c, SA, sp, m, Om);
    for (int j = 0; j < N; j++) TT[j] += Om[j];
    sas_gamma(c, SA, sp, m, Om);
    for (int j = 0; j < N; j++) TT[j] += Om[j];
    sas_exponential(c, SA, sp, m, Om);
    for (int j = 0; j < N; j++) TT[j] += Om[j];
    sas_power(c, SA, sp, m, Om);
    for (int j = 0; j < N; j++) TT[j] += Om[j];
}

/* transport.py:343-359 */
static void calc_SA(int A, const double *sa, double *SA)
{
    SA[0] = 0;
    double acc = 0;
    for (int j = 0; j < A; j++) { acc += sa[j]; SA[j + 1] = acc; }
    SA[0] = 0;
}

/* transport.py:362-509 backward travel-time distribution of one flux */
static void calc_tt(const sas_ctx *c, const double *SA_in, const double *sa, double flux, const double *sas_params, double m,
                    double *tt, double *work)
{
    const int A = c->A, N = A + 1;
    double *SAn = work, *san = SAn + N, *TTn = san + A, *TTi = TTn + N, *tti = TTi + N, *ttqi = tti + A, *Om = ttqi + A;
    double sp[8];
    memcpy(sp, sas_params, sizeof(sp));
    memcpy(SAn, SA_in, sizeof(double) * N);
    memcpy(san, sa, sizeof(double) * A);
    memset(TTn, 0, sizeof(double) * N);
    const double h = 1.0 / c->substeps;
    for (int s = 0; s < c->substeps; s++) {
        sas_sum_families(c, SAn, sp, m, TTi, Om);
        for (int j = 0; j < A; j++) {
            double d = TTi[j + 1] - TTi[j];
            tti[j] = (d >= 0) ? d : 0;
        }
        for (int j = 0; j < A; j++) {
            double q = flux * tti[j] * h;
            ttqi[j] = (q > san[j]) ? san[j] : q;
        }
        for (int j = 0; j < A; j++) tti[j] = (flux * h > 0) ? ttqi[j] / (flux * h) : 0;
        for (int j = 0; j < A; j++) san[j] += -tti[j] * flux * h;
        {
            double acc = 0;
            for (int j = 0; j < A; j++) { acc += san[j]; SAn[j + 1] = acc; }
        }
        {
            double acc = 0;
            for (int j = 0; j < A; j++) { acc += tti[j]; TTn[j + 1] += acc; }
        }
    }
    /* aggregate :480-496 */
    double prev = TTn[0] / c->substeps;
    for (int j = 0; j < A; j++) {
        double cur = TTn[j + 1] / c->substeps;
        tt[j] = cur - prev;
        prev = cur;
    }
    for (int j = 0; j < A; j++) {
        double q = flux * tt[j];
        double ttq = (q > sa[j]) ? sa[j] : q;
        tt[j] = (flux > 0) ? ttq / flux : 0;
    }
}

/* flux concentration transport.py:512-535 with calc_mtt :565-580 (isotope branch) */
static double flux_conc(const sas_ctx *c, const double *msa, const double *tt, double m, double *mtt, double *tmp)
{
    const int A = c->A;
    for (int j = 0; j < A; j++) mtt[j] = ((tt[j] > 0) ? msa[j] : 0) * m;
    for (int j = 0; j < A; j++) tmp[j] = mtt[j] * tt[j];
    double st = np_sum(tt, A);
    double conc = (st > 0) ? np_sum(tmp, A) / st : NAN;
    conc = (conc != 0) ? conc : NAN;
    return conc;
}

/* update_sa transport.py:599-619 */
static void update_sa(int A, double *sa, const double *tt, double flux, double m)
{
    for (int j = 0; j < A; j++) {
        double v = sa[j] + -flux * tt[j];
        v = ((v > -1e-5) && (v < 0)) ? 0 : v;
        sa[j] = v * m;
    }
}

/* a SAS-controlled withdrawal from one layer, optionally received by the other layer
 * (evapotranspiration.py:653-716, :831-895; subsurface_runoff.py:1531-1625, :1753-1815; capillary_rise.py:404-500) */
static void withdrawal(const sas_ctx *c, double *sa_src, double *msa_src, double *sa_dst, double *msa_dst, double flux,
                       const double *sp, double m, double *C_flux, double *C_iso_flux, double *tt_out, double *work)
{
    const int A = c->A, N = A + 1;
    double *SA = work, *tt = SA + N, *mtt = tt + A, *tmp = mtt + A, *w2 = tmp + A;
    calc_SA(A, sa_src, SA);
    for (int j = 0; j < N; j++) SA[j] = SA[j] * m;
    calc_tt(c, SA, sa_src, flux, sp, m, tt, w2);
    for (int j = 0; j < A; j++) tt[j] = tt[j] * m;
    double conc = flux_conc(c, msa_src, tt, m, mtt, tmp);
    *C_flux = conc * m;
    *C_iso_flux = conc_to_delta(c, *C_flux) * m;
    if (sa_dst) { /* isotope mixing into the receiving layer, bin by bin */
        for (int j = 0; j < A; j++) {
            double in = tt[j] * flux;
            double tot = in + sa_dst[j];
            double v = (tot > 0) ? msa_dst[j] * (sa_dst[j] / tot) + mtt[j] * (in / tot) : msa_dst[j];
            msa_dst[j] = v * m;
        }
    }
    update_sa(A, sa_src, tt, flux, m);
    if (sa_dst)
        for (int j = 0; j < A; j++) sa_dst[j] += tt[j] * flux * m;
    for (int j = 0; j < A; j++) msa_src[j] = ((sa_src[j] <= 0) ? 0 : msa_src[j]) * m;
    if (tt_out) memcpy(tt_out, tt, sizeof(double) * A);
}

/* infiltration of `inf` with concentration C_in into bin 0 (infiltration.py:2218-2345, :2441-2500) */
static void inflow(const sas_ctx *c, double *sa, double *msa, double inf, double C_in, double m, double *C_out, double *C_iso_out)
{
    const int A = c->A;
    *C_out = ((inf > 0) ? C_in : 0) * m;
    *C_iso_out = conc_to_delta(c, *C_out) * m;
    double tt0 = ((inf > 0) ? 1 : 0) * m;
    double mtt0 = ((inf > 0) ? C_in : 0) * m;
    for (int j = 0; j < A; j++) {
        double ttj = (j == 0) ? tt0 : 0.0, mttj = (j == 0) ? mtt0 : 0.0;
        double v;
        if (inf * ttj + sa[j] > 0)
            v = msa[j] * (sa[j] / (ttj * inf + sa[j])) + mttj * ((ttj * inf) / (inf * ttj + sa[j]));
        else
            v = msa[j];
        msa[j] = v * m;
    }
    sa[0] += inf * m;
}

/* storage concentration transport.py:538-562 */
static double storage_conc(const sas_ctx *c, const double *sa, const double *msa, double *tmp)
{
    const int A = c->A;
    for (int j = 0; j < A; j++) tmp[j] = msa[j] * sa[j];
    double s = np_sum(sa, A);
    return (s > 0) ? np_sum(tmp, A) / s : 0;
}

/* ageing transport.py:682-739 */
static void ageing(int A, double *sa, double *msa)
{
    double sam1_last = sa[A - 1], msam1_last = msa[A - 1];
    for (int j = A - 1; j >= 1; j--) sa[j] = sa[j - 1];
    for (int j = A - 1; j >= 1; j--) msa[j] = msa[j - 1];
    msa[0] = 0;
    double tot = sa[A - 1] + sam1_last;
    double v = (tot > 0) ? msam1_last * (sam1_last / tot) + msa[A - 1] * (sa[A - 1] / tot) : 0;
    msa[A - 1] = isnan(v) ? 0 : v;
    sa[0] = 0;
    sa[A - 1] += sam1_last;
    sa[A - 1] = (sa[A - 1] < 1e-8) ? 0 : sa[A - 1];
    msa[A - 1] = (sa[A - 1] <= 0) ? NAN : msa[A - 1];
}

/* numpy.interp for one scalar x (numpy/_core/src/multiarray/compiled_base.c: arr_interp + binary_search_with_guess,
 * initial guess 0), as called by calc_age_percentile (transport.py:9-56) */
static long np_binary_search_with_guess(double key, const double *arr, long len, long guess)
{
    long imin = 0, imax = len;
    if (key > arr[len - 1]) return len;
    else if (key < arr[0]) return -1;
    if (len <= 4) {
        long i;
        for (i = 1; i < len && key >= arr[i]; ++i);
        return i - 1;
    }
    if (guess > len - 3) guess = len - 3;
    if (guess < 1) guess = 1;
    if (key < arr[guess]) {
        if (key < arr[guess - 1]) {
            imax = guess - 1;
            if (guess > 8 && key >= arr[guess - 8]) imin = guess - 8;
        } else return guess - 1;
    } else {
        if (key < arr[guess + 1]) return guess;
        else if (key < arr[guess + 2]) return guess + 1;
        else {
            imin = guess + 2;
            if (guess < len - 8 - 1 && key < arr[guess + 8]) imax = guess + 8;
        }
    }
    while (imin < imax) {
        const long imid = imin + ((imax - imin) >> 1);
        if (key >= arr[imid]) imin = imid + 1;
        else imax = imid;
    }
    return imin - 1;
}
static double np_interp1(double x, const double *xp, const double *fp, long n)
{
    if (isnan(x)) return x;
    const long j = np_binary_search_with_guess(x, xp, n, 0);
    if (j == -1) return fp[0];
    if (j == n) return fp[n - 1];
    if (j == n - 1) return fp[j];
    if (xp[j] == x) return fp[j];
    const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double res = slope * (x - xp[j]) + fp[j];
    if (isnan(res)) {
        res = slope * (x - xp[j + 1]) + fp[j + 1];
        if (isnan(res) && fp[j] == fp[j + 1]) res = fp[j];
    }
    return res;
}
/* calc_age_percentile, transport.py:9-56: cdf = cumulative distribution at ages 1..A */
static double age_percentile(int A, const double *cdf, const double *ages, double p)
{
    double v = np_interp1(p, cdf, ages, A);
    return (np_max(cdf, A) <= 0) ? NAN : v;
}
/* transport.py:59-107 (transpiration), :110-158 (subsoil percolation) */
static void flux_age_statistics(int A, const double *tt, const double *ages, double *work, double *st5, double *avg)
{
    double *TT = work, *tmp = work + A;
    double acc = 0;
    for (int j = 0; j < A; j++) { acc += tt[j]; TT[j] = acc; } /* TT[1:] = cumsum(tt) */
    const double pcts[5] = {0.1, 0.25, 0.5, 0.75, 0.9};
    for (int k = 0; k < 5; k++) st5[k] = age_percentile(A, TT, ages, pcts[k]);
    for (int j = 0; j < A; j++) tmp[j] = ages[j] * tt[j];
    *avg = (np_sum(tt, A) > 0) ? np_sum(tmp, A) : NAN;
}
/* transport.py:161-312: residence-time statistics of a storage with StorAge sa (SA = cumsum(sa) * m) */
static void storage_age_statistics(int A, const double *sa, double m, const double *ages, double *work, const double *pcts, int np_, double *stp,
                                   double *avg)
{
    double *SA = work, *RT = SA + (A + 1), *rt = RT + (A + 1), *tmp = rt + A;
    calc_SA(A, sa, SA);
    for (int j = 0; j <= A; j++) SA[j] = SA[j] * m;
    const double mx = np_max(SA, A + 1);
    for (int j = 0; j <= A; j++) RT[j] = (mx > 0) ? SA[j] / mx : 0;
    for (int j = 0; j < A; j++) rt[j] = RT[j + 1] - RT[j];
    for (int k = 0; k < np_; k++) stp[k] = age_percentile(A, RT + 1, ages, pcts[k]);
    for (int j = 0; j < A; j++) tmp[j] = ages[j] * rt[j];
    *avg = (np_sum(rt, A) > 0) ? np_sum(tmp, A) : NAN;
}
static double nan0(double x) { return isnan(x) ? 0 : x; }
/* numpy.isclose(x, y, rtol, atol): (|x - y| <= atol + rtol * |y|) & isfinite(y) | (x == y) */
static int np_isclose(double x, double y, double rtol, double atol) { return ((fabs(x - y) <= atol + rtol * fabs(y)) && isfinite(y)) || (x == y); }
/* storage bounds and tolerances of numerics.sanity_check (set by the test / driver; default: unbounded, 1e-2) */
static const double *g_bounds[4] = {NULL, NULL, NULL, NULL};
static double g_atol = 1e-2, g_rtol = 1e-2;
void sas_oracle_set_sanity_bounds(const double *S_sat_rz, const double *S_sat_ss, const double *S_pwp_rz, const double *S_pwp_ss, double atol,
                                  double rtol)
{
    g_bounds[0] = S_sat_rz; g_bounds[1] = S_sat_ss; g_bounds[2] = S_pwp_rz; g_bounds[3] = S_pwp_ss;
    g_atol = atol; g_rtol = rtol;
}


/* ================================================================================================================
 * Conservative anions (bromide / chloride): msa holds the solute MASS of every age bin.
 * infiltration.py:2350-2422, :2504-2560; evapotranspiration.py:620-651, :905-982; subsurface_runoff.py (anion kernels);
 * capillary_rise.py (anion kernel); root_zone.py / subsoil.py / soil.py anion storage kernels; transport.py:582-596 calc_mtt,
 * :622-679 ageing.
 * ================================================================================================================ */
static const double *g_anion[4] = {NULL, NULL, NULL, NULL}; /* alpha_transp, alpha_q, lu_id, S_sat_rz */
/* settings.solute == 3, the virtual tracer: the soil evaporation carries the solute with partition coefficient 1
 * (evapotranspiration.py:723-791) and the solute balance counts it (numerics.py:517-549) */
static int g_tracer = 0;
void sas_oracle_set_anion_parameters(const double *alpha_transp, const double *alpha_q, const double *lu_id, const double *S_sat_rz)
{
    g_anion[0] = alpha_transp; g_anion[1] = alpha_q; g_anion[2] = lu_id; g_anion[3] = S_sat_rz;
}
static double np_nansum(const double *a, int n, double *tmp)
{
    for (int j = 0; j < n; j++) tmp[j] = isnan(a[j]) ? 0 : a[j];
    return np_sum(tmp, n);
}
/* transport.py:582-596 */
static void anion_mtt(int A, const double *sa, const double *tt, double flux, const double *msa, double alpha, double m, double *mtt)
{
    for (int j = 0; j < A; j++) {
        double v = ((sa[j] > 0) ? (msa[j] / sa[j]) : 0) * alpha * tt[j] * flux;
        v = (v <= 0) ? 0 : v;
        v = (v > msa[j]) ? msa[j] : v;
        mtt[j] = v * m;
    }
}
/* water part of a withdrawal: cumulative StorAge, travel-time distribution (kept in tt) */
static void anion_tt(const sas_ctx *c, const double *sa_src, double flux, const double *sp, double m, double *tt, double *work)
{
    const int A = c->A, N = A + 1;
    double *SA = work, *w2 = SA + N;
    calc_SA(A, sa_src, SA);
    for (int j = 0; j < N; j++) SA[j] = SA[j] * m;
    calc_tt(c, SA, sa_src, flux, sp, m, tt, w2);
    for (int j = 0; j < A; j++) tt[j] = tt[j] * m;
}
static void keep_tt(double *dst, const double *tt, int A)
{
    if (dst) memcpy(dst, tt, sizeof(double) * (size_t)A);
}
static void anion_ageing(int A, double *sa, double *msa)
{
    const double sam1_last = sa[A - 1], msam1_last = msa[A - 1];
    for (int j = A - 1; j >= 1; j--) sa[j] = sa[j - 1];
    sa[0] = 0;
    sa[A - 1] += sam1_last;
    sa[A - 1] = (sa[A - 1] < 1e-8) ? 0 : sa[A - 1];
    for (int j = A - 1; j >= 1; j--) msa[j] = msa[j - 1];
    msa[0] = 0;
    msa[A - 1] += msam1_last;
}

/* ---- nitrate: soil nitrogen cycle on the age-ranked arrays (roger/core/nitrate.py:68-696), called between the capillary rise
 * and the storage kernels of the anion day (transport.py:966-969) ------------------------------------------------------- */
static struct {
    const double *params;  /* RNP_COUNT planes */
    const double *forcing; /* RNF_COUNT planes of the day */
    double *nmin_rz, *nmin_ss; /* (cell, age) */
    double *out;           /* RNO_COUNT planes */
    double dx, dy, zmax, dt;
} g_nit = {NULL, NULL, NULL, NULL, NULL, 1, 1, 0.7, 24};
void sas_oracle_set_nitrate(const double *params, const double *forcing, double *nmin_rz, double *nmin_ss, double *out, double dx, double dy,
                            double zroot_to_zsoil_max, double dt_hours)
{
    g_nit.params = params; g_nit.forcing = forcing; g_nit.nmin_rz = nmin_rz; g_nit.nmin_ss = nmin_ss; g_nit.out = out;
    g_nit.dx = dx; g_nit.dy = dy; g_nit.zmax = zroot_to_zsoil_max; g_nit.dt = dt_hours;
}
/* soil temperature coefficient of the process kernels: T / (30 - lo) inside [lo, 30], 1 above 30, 0 below lo; "* maskCatch" twice */
static double temp_coeff(double T, double lo, double m)
{
    double c = (((T >= lo) && (T <= 30)) ? T / (30 - lo) : 0) * m;
    c = ((T > 30) ? 1 : c) * m;
    return c;
}
/* nitrate.py:176-257 calc_nit_soil / :349-427 calc_gaseous_loss / :430-467 calc_ammonium_uptake share the distribution step:
 * r[j] = (sum(w) > 0) ? (w[j] / sum(w)) * pot : 0, limited to Nmin[j] and to >= 0 (each statement "* maskCatch") */
static void distribute(int A, const double *w, double wsum, double pot, const double *nmin, double m, double *r)
{
    for (int j = 0; j < A; j++) {
        double v = ((wsum > 0) ? ((w[j] / wsum) * pot) : 0.0) * m;
        v = ((v > nmin[j]) ? nmin[j] : v) * m;
        v = ((v < 0) ? 0 : v) * m;
        r[j] = v;
    }
}
/* nitrate.py:68-173 calc_denit_soil: the oldest solute is removed first */
static void denit_soil(int A, const double *msa, double km, double Dmax, const double *sa, double S_sat, double T, double m, double *mra,
                       double *tmp)
{
    const double f = (g_nit.dt / (365 * 24)) * g_nit.dx * g_nit.dy * 100;
    const double S = np_sum(sa, A) * m;
    const double coeff = temp_coeff(T, 5, m);
    const double ms = np_sum(msa, A) * m;
    double mr_pot = (Dmax * (g_nit.dt / (365 * 24)) * g_nit.dx * g_nit.dy * 100 *
                     (ms / (km * (g_nit.dt / (365 * 24)) * g_nit.dx * g_nit.dy * 100 + ms))) * coeff * m;
    (void)f;
    mr_pot = ((S >= 0.7 * S_sat) ? mr_pot : 0) * m;
    /* ms_cuml = cumsum(msa[::-1])[::-1] * m */
    double acc = 0;
    for (int j = A - 1; j >= 0; j--) {
        acc = (j == A - 1) ? msa[j] : acc + msa[j];
        tmp[j] = ((acc * m < mr_pot) ? 0 : msa[j]) * m; /* msa with the bins younger-side of the potential emptied ... */
    }
    tmp[A - 1] += -((tmp[A - 1] >= mr_pot) ? mr_pot : 0) * m;
    for (int j = 0; j < A; j++) {
        double v = (msa[j] * m - tmp[j]) * m;
        v = ((v < 0) ? 0 : v) * m;
        mra[j] = v;
    }
}
static void nitrogen_cycle_cell(int A, int64_t C, int64_t i, double *s_rz, double *s_ss, double *ms_rz, double *ms_ss, double transp, double m,
                                double *work)
{
#define NP(k) g_nit.params[(size_t)(k) * (size_t)C + (size_t)i]
#define NF(k) g_nit.forcing[(size_t)(k) * (size_t)C + (size_t)i]
#define NO(k) g_nit.out[(size_t)(k) * (size_t)C + (size_t)i]
    double *n_rz = g_nit.nmin_rz + (size_t)i * A, *n_ss = g_nit.nmin_ss + (size_t)i * A;
    double *ma = work, *r = ma + (A + 2), *mr_rz = r + (A + 2), *mr_ss = mr_rz + (A + 2), *tmp = mr_ss + (A + 2);
    const double T = NF(RNF_temp_soil), z_root = NF(RNF_z_root), z_soil = NP(RNP_z_soil), lu = NF(RNF_lu_id);
    const double dt = g_nit.dt, dx = g_nit.dx, dy = g_nit.dy;
    /* calc_n_fixation :305-346, calc_min_soil :260-302 */
    double nfix = NP(RNP_kfix_rz) * (dt / (365 * 24)) * dx * dy * 100 * temp_coeff(T, 0, m) * (z_root / (g_nit.zmax * z_soil)) * m;
    {
        static const double fixers[10] = {541, 577, 578, 580, 581, 583, 584, 586, 587, 588};
        int is = 0;
        for (int k = 0; k < 10; k++) is = is || (lu == fixers[k]);
        nfix = is ? nfix : 0;
    }
    nfix = nfix * m;
    NO(RNO_nfix_s) = nfix;
    const double min_rz = (NP(RNP_kmin_rz) * (dt / (365 * 24)) * dx * dy * 100 * temp_coeff(T, 0, m) * m) * m;
    const double min_ss = (NP(RNP_kmin_ss) * (dt / (365 * 24)) * dx * dy * 100 * temp_coeff(T, 0, m) * m) * m;
    NO(RNO_min_s) = min_rz + min_ss;
    n_rz[0] += nfix;
    n_rz[0] += min_rz * m;
    n_ss[0] += min_ss * m;
    /* nitrification of the root zone :176-257 */
    {
        const double Ssum = np_sum(s_rz, A);
        const double S = Ssum * m;
        const double nsum = np_sum(n_rz, A);
        double pot = (NP(RNP_dmax_nit_rz) * (dt / (365 * 24)) * dx * dy * 100 *
                      (nsum / (NP(RNP_km_nit_rz) * (dt / (365 * 24)) * dx * dy * 100 + nsum))) * temp_coeff(T, 1, m) * m;
        pot = ((S < 0.9 * NP(RNP_S_sat_rz)) ? pot : 0) * m;
        distribute(A, s_rz, Ssum, pot, n_rz, m, ma);
        for (int j = 0; j < A; j++) ma[j] = ma[j] * m; /* vs.ma_rz = calc_nit_soil(...) * maskCatch */
        for (int j = 0; j < A; j++) n_rz[j] += -ma[j];
    }
    /* gaseous loss :349-427 */
    {
        const double S = np_sum(s_rz, A) * m;
        double pot = (NP(RNP_kngl_rz) * (dt / (365 * 24)) * dx * dy * 100) * temp_coeff(T, 5, m) * m;
        pot = ((S < 0.9 * NP(RNP_S_sat_rz)) ? pot : 0) * m;
        const double nsum = np_sum(n_rz, A);
        distribute(A, n_rz, nsum, pot, n_rz, m, r);
        NO(RNO_ngas_s) = np_sum(r, A);
        for (int j = 0; j < A; j++) n_rz[j] += -r[j] * m;
    }
    /* ammonium uptake :430-467 */
    {
        const double pot = ((transp > 0) ? NF(RNF_nup) * (z_root / (z_soil * g_nit.zmax)) : 0) * m;
        const double nsum = np_sum(n_rz, A);
        distribute(A, n_rz, nsum, pot, n_rz, m, r);
        for (int j = 0; j < A; j++) tmp[j] = r[j];
        NO(RNO_nh4_up) = np_sum(tmp, A) * m;
        for (int j = 0; j < A; j++) n_rz[j] += -r[j] * m;
    }
    for (int j = 0; j < A; j++) ms_rz[j] += ma[j];
    for (int j = 0; j < A; j++) n_ss[j] += -0.0; /* vs.ma_ss is never written by the reference: 0 */
    /* denitrification :68-173 */
    denit_soil(A, ms_rz, NP(RNP_km_denit_rz), NP(RNP_dmax_denit_rz), s_rz, NP(RNP_S_sat_rz), T, m, mr_rz, tmp);
    for (int j = 0; j < A; j++) mr_rz[j] = mr_rz[j] * m;
    for (int j = 0; j < A; j++) ms_rz[j] += -mr_rz[j];
    denit_soil(A, ms_ss, NP(RNP_km_denit_ss), NP(RNP_dmax_denit_ss), s_ss, NP(RNP_S_sat_ss), T, m, mr_ss, tmp);
    for (int j = 0; j < A; j++) mr_ss[j] = mr_ss[j] * m;
    for (int j = 0; j < A; j++) ms_ss[j] += -mr_ss[j];
    for (int j = 0; j < A; j++) tmp[j] = ma[j] + 0.0; /* ma_s = ma_rz + ma_ss */
    NO(RNO_nit_s) = np_sum(tmp, A);
    for (int j = 0; j < A; j++) tmp[j] = mr_rz[j] + mr_ss[j];
    NO(RNO_denit_s) = np_sum(tmp, A);
    NO(RNO_Nmin_s) = np_sum(n_rz, A) + np_sum(n_ss, A) * m;
#undef NP
#undef NF
#undef NO
}
/* transport.py:655-679 calc_ageing_msa on the mineral nitrogen (:809-824) */
static void nmin_ageing(int A, double *n)
{
    const double last = n[A - 1];
    for (int j = A - 1; j >= 1; j--) n[j] = n[j - 1];
    n[0] = 0;
    n[A - 1] += last;
}

static void flux_age_statistics(int A, const double *tt, const double *ages, double *work, double *st5, double *avg);
static void storage_age_statistics(int A, const double *sa, double m, const double *ages, double *work, const double *pcts, int np_, double *stp,
                                   double *avg);
static double nan0(double x);

/* crop variant of the offline model (svat_crop_transport_model_deterministic, transport.py:1042-1087): the day opens with the
 * redistribution after root growth (subsoil -> root zone) and root loss (root zone -> subsoil), crop.py:2461-2558, :2729-2828 */
static struct {
    const double *sp_rg, *sp_rl; /* (C, 8) SAS parameters */
    const double *re;            /* (2, C): today's re_rg, re_rl */
    double *out;                 /* (6, C): RCO_* */
} g_crop = {NULL, NULL, NULL, NULL};
void sas_oracle_set_crop(const double *sp_rg, const double *sp_rl, const double *re, double *out)
{
    g_crop.sp_rg = sp_rg; g_crop.sp_rl = sp_rl; g_crop.re = re; g_crop.out = out;
}
static void redistribution(const sas_ctx *c, double *s_src, double *ms_src, double *s_dst, double *ms_dst, double re_in, const double *sp, double m,
                           double *tt, double *mtt, double *w2, double *re_out, double *C_out, double *M_out)
{
    const int A = c->A;
    const double tot = np_sum(s_src, A);
    const double re = ((re_in > tot) ? tot : re_in) * m;
    anion_tt(c, s_src, re, sp, m, tt, w2);
    anion_mtt(A, s_src, tt, re, ms_src, 1.0, m, mtt);
    const double M = np_sum(mtt, A);
    *re_out = re;
    *C_out = ((re > 0) ? M / re : 0) * m;
    *M_out = M * m;
    update_sa(A, s_src, tt, re, m);
    for (int j = 0; j < A; j++) {
        s_dst[j] += tt[j] * re * m;
        ms_dst[j] += mtt[j] * m;
        ms_src[j] += -mtt[j] * m;
    }
}

static void anion_day_cell(const sas_ctx *c, int64_t C, int64_t i, double *s_rz, double *s_ss, double *ms_rz, double *ms_ss,
                           const double *forcing, double m, const double *sas_params, double *out, double *tt_out, double *work, double *tmp,
                           double *stats, double *balance_prev, double *ttk, const double *ages)
{
    const int A = c->A;
    double sum_sa_s_m1 = 0;
    if (stats) { /* vs.sa_s[taum1] as this step's soil kernel recomputes it */
        for (int j = 0; j < A; j++) tmp[j] = s_rz[j] + s_ss[j] * m;
        sum_sa_s_m1 = np_sum(tmp, A);
    }
#define F(k) forcing[(size_t)(k) * (size_t)C + (size_t)i]
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
#define SP(p) (sas_params + ((size_t)(p) * (size_t)C + (size_t)i) * 8)
#define TTO(p) (tt_out ? tt_out + ((size_t)(p) * (size_t)C + (size_t)i) * A : \
                (stats ? ((p) == RSP_evap_soil ? ttk : (p) == RSP_transp ? ttk + (A + 2) : (p) == RSP_q_ss ? ttk + 2 * (A + 2) : NULL) : NULL))
#define ST(k) stats[(size_t)(k) * (size_t)C + (size_t)i]
    double *tt = work, *mtt = tt + (A + 2), *w2 = mtt + (A + 2);
    const double C_in = F(RSF_C_in);
    const double alpha_q = g_anion[1] ? g_anion[1][i] : 1.0;
    if (g_crop.re) {
#define CO(k) g_crop.out[(size_t)(k) * (size_t)C + (size_t)i]
        redistribution(c, s_ss, ms_ss, s_rz, ms_rz, g_crop.re[i], g_crop.sp_rg + (size_t)i * 8, m, tt, mtt, w2, &CO(RCO_re_rg), &CO(RCO_C_re_rg),
                       &CO(RCO_M_re_rg));
        redistribution(c, s_rz, ms_rz, s_ss, ms_ss, g_crop.re[(size_t)C + (size_t)i], g_crop.sp_rl + (size_t)i * 8, m, tt, mtt, w2, &CO(RCO_re_rl),
                       &CO(RCO_C_re_rl), &CO(RCO_M_re_rl));
#undef CO
    }
    /* infiltration into the root zone */
    O(RSO_C_inf_mat_rz) = ((F(RSF_inf_mat_rz) > 0) ? C_in : 0) * m;
    O(RSO_C_inf_pf_rz) = ((F(RSF_inf_pf_rz) > 0) ? C_in : 0) * m;
    O(RSO_C_iso_inf_mat_rz) = O(RSO_C_inf_mat_rz) * F(RSF_inf_mat_rz) * m; /* M_inf_mat_rz */
    O(RSO_C_iso_inf_pf_rz) = O(RSO_C_inf_pf_rz) * F(RSF_inf_pf_rz) * m;    /* M_inf_pf_rz  */
    s_rz[0] += F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) * m;
    ms_rz[0] += O(RSO_C_iso_inf_mat_rz) + O(RSO_C_iso_inf_pf_rz) * m;
    /* soil evaporation: water only (evapotranspiration.py:620-651); the virtual tracer leaves with it (:723-791, alpha = 1 without mask) */
    anion_tt(c, s_rz, F(RSF_evap_soil), SP(RSP_evap_soil), m, tt, w2);
    if (g_tracer) {
        anion_mtt(A, s_rz, tt, F(RSF_evap_soil), ms_rz, 1.0, m, mtt);
        const double M = np_sum(mtt, A);
        O(RSO_C_evap_soil) = ((F(RSF_evap_soil) > 0) ? M / F(RSF_evap_soil) : 0) * m;
        O(RSO_C_iso_evap_soil) = M * m; /* M_evap_soil */
        update_sa(A, s_rz, tt, F(RSF_evap_soil), m);
        for (int j = 0; j < A; j++) ms_rz[j] += -mtt[j] * m;
    } else {
        update_sa(A, s_rz, tt, F(RSF_evap_soil), m);
        O(RSO_C_evap_soil) = 0;
        O(RSO_C_iso_evap_soil) = 0;
    }
    keep_tt(TTO(RSP_evap_soil), tt, A);
    /* transpiration */
    {
        anion_tt(c, s_rz, F(RSF_transp), SP(RSP_transp), m, tt, w2);
        const double lu = g_anion[2] ? g_anion[2][i] : 0.0, sat = g_anion[3] ? g_anion[3][i] : INFINITY;
        const int crop_wet = (lu > 500) && (lu < 599) && (np_sum(s_rz, A) >= 0.8 * sat);
        const double alpha = (crop_wet ? 0 : (g_anion[0] ? g_anion[0][i] : 1.0)) * m;
        anion_mtt(A, s_rz, tt, F(RSF_transp), ms_rz, alpha, m, mtt);
        const double M = np_sum(mtt, A);
        O(RSO_C_transp) = ((F(RSF_transp) > 0) ? M / F(RSF_transp) : 0) * m;
        O(RSO_C_iso_transp) = M * m;
        update_sa(A, s_rz, tt, F(RSF_transp), m);
        for (int j = 0; j < A; j++) ms_rz[j] += -mtt[j] * m;
        keep_tt(TTO(RSP_transp), tt, A);
    }
    /* percolation of the root zone into the subsoil */
    {
        anion_tt(c, s_rz, F(RSF_q_rz), SP(RSP_q_rz), m, tt, w2);
        anion_mtt(A, s_rz, tt, F(RSF_q_rz), ms_rz, alpha_q, m, mtt);
        const double M = np_sum(mtt, A);
        O(RSO_C_q_rz) = ((F(RSF_q_rz) > 0) ? M / F(RSF_q_rz) : 0) * m;
        O(RSO_C_iso_q_rz) = M * m;
        update_sa(A, s_rz, tt, F(RSF_q_rz), m);
        for (int j = 0; j < A; j++) {
            ms_rz[j] += -mtt[j] * m;
            ms_ss[j] += mtt[j] * m;
            s_ss[j] += tt[j] * F(RSF_q_rz) * m;
        }
        keep_tt(TTO(RSP_q_rz), tt, A);
    }
    /* infiltration into the subsoil */
    O(RSO_C_inf_pf_ss) = ((F(RSF_inf_pf_ss) > 0) ? C_in : 0) * m;
    O(RSO_C_iso_inf_pf_ss) = O(RSO_C_inf_pf_ss) * F(RSF_inf_pf_ss) * m; /* M_inf_pf_ss */
    s_ss[0] += F(RSF_inf_pf_ss) * m;
    ms_ss[0] += O(RSO_C_iso_inf_pf_ss) * m;
    /* percolation of the subsoil */
    {
        anion_tt(c, s_ss, F(RSF_q_ss), SP(RSP_q_ss), m, tt, w2);
        anion_mtt(A, s_ss, tt, F(RSF_q_ss), ms_ss, alpha_q, m, mtt);
        const double M = np_sum(mtt, A);
        O(RSO_C_q_ss) = ((F(RSF_q_ss) > 0) ? M / F(RSF_q_ss) : 0) * m;
        O(RSO_C_iso_q_ss) = M * m;
        update_sa(A, s_ss, tt, F(RSF_q_ss), m);
        for (int j = 0; j < A; j++) ms_ss[j] += -mtt[j] * m;
        keep_tt(TTO(RSP_q_ss), tt, A);
    }
    /* capillary rise into the root zone */
    {
        anion_tt(c, s_ss, F(RSF_cpr_rz), SP(RSP_cpr_rz), m, tt, w2);
        anion_mtt(A, s_ss, tt, F(RSF_cpr_rz), ms_ss, alpha_q, m, mtt);
        const double M = np_sum(mtt, A);
        O(RSO_C_cpr_rz) = ((F(RSF_cpr_rz) > 0) ? M / F(RSF_cpr_rz) : 0) * m;
        O(RSO_C_iso_cpr_rz) = M * m;
        update_sa(A, s_ss, tt, F(RSF_cpr_rz), m);
        for (int j = 0; j < A; j++) {
            ms_ss[j] += -mtt[j] * m;
            s_rz[j] += tt[j] * F(RSF_cpr_rz) * m;
            ms_rz[j] += mtt[j] * m;
        }
        keep_tt(TTO(RSP_cpr_rz), tt, A);
    }
    if (g_nit.params && g_nit.nmin_rz) /* settings.enable_nitrate: soil nitrogen cycle (transport.py:966-969) */
        nitrogen_cycle_cell(A, C, i, s_rz, s_ss, ms_rz, ms_ss, F(RSF_transp), m, work);
    /* storage kernels */
    for (int j = 0; j < A; j++) s_rz[j] = (s_rz[j] < 1e-8) ? 0 : s_rz[j];
    for (int j = 0; j < A; j++) ms_rz[j] = (s_rz[j] <= 0) ? 0 : ms_rz[j];
    {
        const double M = np_nansum(ms_rz, A, tmp) * m, S = np_sum(s_rz, A);
        O(RSO_C_iso_rz) = M;
        O(RSO_C_rz) = (S > 0) ? M / S : 0;
    }
    for (int j = 0; j < A; j++) s_ss[j] = (s_ss[j] < 1e-8) ? 0 : s_ss[j];
    for (int j = 0; j < A; j++) ms_ss[j] = (s_ss[j] <= 0) ? 0 : ms_ss[j];
    {
        const double M = np_nansum(ms_ss, A, tmp) * m, S = np_sum(s_ss, A);
        O(RSO_C_iso_ss) = M;
        O(RSO_C_ss) = (S > 0) ? M / S : 0;
    }
    {
        double *sa_s = tmp + (A + 2), *msa_s = sa_s + (A + 2);
        for (int j = 0; j < A; j++) {
            sa_s[j] = s_rz[j] + s_ss[j] * m;
            msa_s[j] = ms_rz[j] + ms_ss[j] * m;
        }
        const double M = np_nansum(msa_s, A, tmp) * m, S = np_sum(sa_s, A);
        O(RSO_C_iso_s) = M;
        O(RSO_C_s) = (S > 0) ? M / S : 0;
        if (stats) { /* age statistics (water only, transport.py:59-312) and balance errors (numerics.py:424-445, :640-673) */
            const double p5[5] = {0.1, 0.25, 0.5, 0.75, 0.9}, p3[3] = {0.25, 0.5, 0.75};
            double s5[5], s3[3], avg, w[3];
            const int procs[3] = {RSP_evap_soil, RSP_transp, RSP_q_ss};
            const double flx[3] = {F(RSF_evap_soil), F(RSF_transp), F(RSF_q_ss)};
            for (int k = 0; k < 3; k++) {
                const double *t = TTO(procs[k]);
                for (int j = 0; j < A; j++) tmp[j] = flx[k] * t[j];
                w[k] = np_sum(tmp, A);
            }
            flux_age_statistics(A, TTO(RSP_transp), ages, w2, s5, &avg);
            for (int k = 0; k < 5; k++) ST(RSS_tt10_transp + k) = s5[k];
            ST(RSS_ttavg_transp) = avg;
            flux_age_statistics(A, TTO(RSP_q_ss), ages, w2, s5, &avg);
            for (int k = 0; k < 5; k++) ST(RSS_tt10_q_ss + k) = s5[k];
            ST(RSS_ttavg_q_ss) = avg;
            storage_age_statistics(A, s_rz, m, ages, w2, p3, 3, s3, &avg);
            for (int k = 0; k < 3; k++) ST(RSS_rt25_rz + k) = s3[k];
            ST(RSS_rtavg_rz) = avg;
            storage_age_statistics(A, s_ss, m, ages, w2, p3, 3, s3, &avg);
            for (int k = 0; k < 3; k++) ST(RSS_rt25_ss + k) = s3[k];
            ST(RSS_rtavg_ss) = avg;
            storage_age_statistics(A, sa_s, m, ages, w2, p5, 5, s5, &avg);
            for (int k = 0; k < 5; k++) ST(RSS_rt10_s + k) = s5[k];
            ST(RSS_rtavg_s) = avg;
            double *bp_C = balance_prev + i;
            ST(RSS_dS_num_error) = fabs((S - sum_sa_s_m1) - (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - w[0] - w[1] - w[2]) * 1.0);
            if (g_nit.params && g_nit.nmin_rz) /* nitrate (numerics.py:673-705): the denitrified and the nitrified mass of the day as well */
                ST(RSS_dC_num_error) = fabs((S * O(RSO_C_s) - sum_sa_s_m1 * *bp_C) -
                                            (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                             F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - w[1] * nan0(O(RSO_C_transp)) - w[2] * nan0(O(RSO_C_q_ss)) -
                                             g_nit.out[(size_t)RNO_denit_s * (size_t)C + (size_t)i] + g_nit.out[(size_t)RNO_nit_s * (size_t)C + (size_t)i]) * 1.0);
            else if (g_tracer) /* numerics.py:517-549: the evaporated tracer as well */
                ST(RSS_dC_num_error) = fabs((S * O(RSO_C_s) - sum_sa_s_m1 * *bp_C) -
                                            (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                             F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - w[0] * nan0(O(RSO_C_evap_soil)) -
                                             w[1] * nan0(O(RSO_C_transp)) - w[2] * nan0(O(RSO_C_q_ss))) * 1.0);
            else
            ST(RSS_dC_num_error) = fabs((S * O(RSO_C_s) - sum_sa_s_m1 * *bp_C) -
                                        (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                         F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - w[1] * nan0(O(RSO_C_transp)) - w[2] * nan0(O(RSO_C_q_ss))) * 1.0);
            ST(RSS_sanity_ok) = NAN; /* the anion branch of numerics.sanity_check is not restated */
            *bp_C = O(RSO_C_s);
        }
    }
    anion_ageing(A, s_rz, ms_rz);
    anion_ageing(A, s_ss, ms_ss);
    if (g_nit.params && g_nit.nmin_rz) {
        nmin_ageing(A, g_nit.nmin_rz + (size_t)i * A);
        nmin_ageing(A, g_nit.nmin_ss + (size_t)i * A);
    }
#undef F
#undef O
#undef SP
#undef TTO
#undef ST
}


/* ---- explicit Euler solver (settings.sas_solver = "Euler"), isotope branch --------------------------------------------
 * svat_transport_model_euler (transport.py:2063-2413) with calc_TT_num (:859-907) and calc_TT_num_nonneg (:910-945);
 * driven `substeps` times per day by calculate_storage_selection (:3221-3258) with settings.h = 1 / substeps. */

/* backward travel-time distribution of one flux from the StorAge snapshot SA (nages points) */
static void euler_tt(const sas_ctx *c, const double *SA, const double *sas_params, double fh, double m, double *tt, double *work)
{
    const int A = c->A, N = A + 1;
    double *TTq = work, *Om = TTq + N, *ttq = Om + N, *TTn = ttq + N;
    double sp[8];
    memcpy(sp, sas_params, sizeof(sp));
    sas_sum_families(c, SA, sp, m, TTq, Om);
    for (int j = 0; j < N; j++) TTq[j] = (fh <= 0) ? 0 : TTq[j];
    for (int j = 0; j < A; j++) {
        const double sa_d = SA[j + 1] - SA[j];
        double q = (TTq[j + 1] - TTq[j]) * fh;
        q = (sa_d + q < 0) ? -sa_d : q;
        q = (q == 0) ? 0.0 : q; /* where(ttq == -0, 0, ttq): both zeros become +0 */
        ttq[j] = q;
    }
    const double ssum = np_sum(ttq, A);
    for (int j = 0; j < A; j++) ttq[j] = (ttq[j] > 0) ? ttq[j] / ssum : 0;
    TTn[0] = 0;
    {
        double acc = ttq[0];
        TTn[1] = acc;
        for (int j = 1; j < A; j++) { acc += ttq[j]; TTn[j + 1] = acc; }
    }
    for (int j = 0; j < A; j++) {
        const double d = TTn[j + 1] - TTn[j];
        tt[j] = (d >= 0) ? d : 0;
    }
}

/* isotope mixing of an inflow dsa1 carrying dmsa1 into one bin: the two-statement update of transport.py:2115-2130 /
 * :2278-2293; `need_pos` is the extra (msa > 0) condition the root-zone statements carry */
static double euler_mix(double msa, double sa, double dsa1, double dmsa1, int need_pos)
{
    const double tot = dsa1 + sa;
    const int c1 = need_pos ? ((tot > 0) && (msa > 0)) : (tot > 0);
    double v = (c1 ? msa * (sa / tot) : 0) + ((tot > 0) ? dmsa1 * (dsa1 / tot) : 0);
    v = ((dsa1 > 0) && (v <= 0)) ? dmsa1 : v;
    return v;
}

static double euler_flux_conc(int A, const double *mtt, const double *tt, double *tmp)
{ /* calc_conc_iso_flux transport.py:512-535 */
    for (int j = 0; j < A; j++) tmp[j] = mtt[j] * tt[j];
    const double st = np_sum(tt, A);
    double conc = (st > 0) ? np_sum(tmp, A) / st : NAN;
    return (conc != 0) ? conc : NAN;
}

static void euler_substep_cell(const sas_ctx *c, int64_t C, int64_t i, double *s_rz, double *s_ss, double *ms_rz, double *ms_ss,
                               const double *forcing, double m, const double *sas_params, double *out, double *tt_base, size_t tt_stride,
                               double *work)
{ /* tt_base (may be NULL): this cell's travel-time distributions, process p at tt_base + p * tt_stride */
#define F(k) forcing[(size_t)(k) * (size_t)C + (size_t)i]
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
#define SP(p) (sas_params + ((size_t)(p) * (size_t)C + (size_t)i) * 8)
    const int A = c->A, N = A + 1;
    const double h = 1.0 / c->substeps;
    double *SA_rz = work, *SA_ss = SA_rz + N, *tt_e = SA_ss + N, *tt_t = tt_e + N, *tt_q = tt_t + N, *tt_c = tt_q + N, *tt_s = tt_c + N;
    double *tmp = tt_s + N, *mtt = tmp + N, *w2 = mtt + N;
    const double inf_mat = F(RSF_inf_mat_rz), inf_pf = F(RSF_inf_pf_rz), inf_ss = F(RSF_inf_pf_ss), C_in = F(RSF_C_in);
    const double evap = F(RSF_evap_soil), transp = F(RSF_transp), q_rz = F(RSF_q_rz), q_ss = F(RSF_q_ss), cpr = F(RSF_cpr_rz);
    /* upper boundary condition :2071-2143: infiltration enters age bin 0 */
    for (int j = 0; j < A; j++) {
        const double tt_im = (j == 0) ? ((inf_mat > 0) ? 1 : 0) * m : 0.0, tt_ip = (j == 0) ? ((inf_pf > 0) ? 1 : 0) * m : 0.0;
        const double tt_is = (j == 0) ? ((inf_ss > 0) ? 1 : 0) * m : 0.0;
        const double mtt_im = (j == 0) ? ((inf_mat > 0) ? C_in : 0) * m : 0.0, mtt_ip = (j == 0) ? ((inf_pf > 0) ? C_in : 0) * m : 0.0;
        const double mtt_is = (j == 0) ? ((inf_ss > 0) ? C_in : 0) * m : 0.0;
        const double dsa_rz = (inf_mat * tt_im + inf_pf * tt_ip) * h;
        const double dsa_ss = (inf_ss * tt_is) * h;
        const double dmsa_rz1 = nan0(mtt_im) * ((dsa_rz > 0) ? ((inf_mat * tt_im * h) / dsa_rz) : 0) +
                                nan0(mtt_ip) * ((dsa_rz > 0) ? ((inf_pf * tt_ip * h) / dsa_rz) : 0);
        const double dmsa_ss1 = nan0(mtt_is) * ((dsa_ss > 0) ? ((inf_ss * tt_is * h) / dsa_ss) : 0);
        ms_rz[j] = euler_mix(ms_rz[j], s_rz[j], dsa_rz, dmsa_rz1, 1);
        ms_ss[j] = euler_mix(ms_ss[j], s_ss[j], dsa_ss, dmsa_ss1, 0);
        s_rz[j] += dsa_rz;
        s_ss[j] += dsa_ss;
        ms_rz[j] = (s_rz[j] <= 0) ? 0 : ms_rz[j];
        ms_ss[j] = (s_ss[j] <= 0) ? 0 : ms_ss[j];
    }
    calc_SA(A, s_rz, SA_rz);
    for (int j = 0; j < N; j++) SA_rz[j] = SA_rz[j] * m;
    calc_SA(A, s_ss, SA_ss);
    for (int j = 0; j < N; j++) SA_ss[j] = SA_ss[j] * m;
    /* travel-time distributions of the five fluxes from the same snapshot :2157-2250 */
    euler_tt(c, SA_rz, SP(RSP_evap_soil), evap * h, m, tt_e, w2);
    euler_tt(c, SA_rz, SP(RSP_transp), transp * h, m, tt_t, w2);
    euler_tt(c, SA_rz, SP(RSP_q_rz), q_rz * h, m, tt_q, w2);
    euler_tt(c, SA_ss, SP(RSP_cpr_rz), cpr * h, m, tt_c, w2);
    euler_tt(c, SA_ss, SP(RSP_q_ss), q_ss * h, m, tt_s, w2);
    /* flux concentrations :2322-2357 (calc_mtt isotope branch :575-580: mtt = where(tt > 0, msa, 0)) + delta_fluxes_svat :3659 */
    {
        const double *tts[5] = {tt_e, tt_t, tt_q, tt_c, tt_s};
        const double *src[5] = {ms_rz, ms_rz, ms_rz, ms_ss, ms_ss};
        const int oc[5] = {RSO_C_evap_soil, RSO_C_transp, RSO_C_q_rz, RSO_C_cpr_rz, RSO_C_q_ss};
        const int oi[5] = {RSO_C_iso_evap_soil, RSO_C_iso_transp, RSO_C_iso_q_rz, RSO_C_iso_cpr_rz, RSO_C_iso_q_ss};
        const int pr[5] = {RSP_evap_soil, RSP_transp, RSP_q_rz, RSP_cpr_rz, RSP_q_ss};
        for (int k = 0; k < 5; k++) {
            for (int j = 0; j < A; j++) mtt[j] = (tts[k][j] > 0) ? src[k][j] : 0;
            O(oc[k]) = euler_flux_conc(A, mtt, tts[k], tmp) * m;
            O(oi[k]) = conc_to_delta(c, O(oc[k])) * m;
            if (tt_base) memcpy(tt_base + (size_t)pr[k] * tt_stride, tts[k], sizeof(double) * A);
        }
        O(RSO_C_inf_mat_rz) = ((inf_mat > 0) ? C_in : NAN) * m;
        O(RSO_C_inf_pf_rz) = ((inf_pf > 0) ? C_in : NAN) * m;
        O(RSO_C_inf_pf_ss) = ((inf_ss > 0) ? C_in : NAN) * m;
        O(RSO_C_iso_inf_mat_rz) = conc_to_delta(c, O(RSO_C_inf_mat_rz)) * m;
        O(RSO_C_iso_inf_pf_rz) = conc_to_delta(c, O(RSO_C_inf_pf_rz)) * m;
        O(RSO_C_iso_inf_pf_ss) = conc_to_delta(c, O(RSO_C_inf_pf_ss)) * m;
    }
    /* update StorAge :2266-2319 */
    for (int j = 0; j < A; j++) {
        const double mtt_c = (tt_c[j] > 0) ? ms_ss[j] : 0, mtt_q = (tt_q[j] > 0) ? ms_rz[j] : 0;
        double dsa_rz = (cpr * tt_c[j] - evap * tt_e[j] - transp * tt_t[j] - q_rz * tt_q[j]) * h;
        dsa_rz = (s_rz[j] + dsa_rz < 0) ? -s_rz[j] : dsa_rz;
        double dsa_ss = (q_rz * tt_q[j] - cpr * tt_c[j] - q_ss * tt_s[j]) * h;
        dsa_ss = (s_ss[j] + dsa_ss < 0) ? -s_ss[j] : dsa_ss;
        const double dsa_rz1 = (cpr * tt_c[j]) * h;
        const double dmsa_rz1 = nan0(mtt_c) * ((dsa_rz1 > 0) ? ((cpr * tt_c[j] * h) / dsa_rz1) : 0);
        const double dsa_ss1 = (q_rz * tt_q[j]) * h;
        const double dmsa_ss1 = nan0(mtt_q) * ((dsa_ss1 > 0) ? ((q_rz * tt_q[j] * h) / dsa_ss1) : 0);
        ms_rz[j] = euler_mix(ms_rz[j], s_rz[j], dsa_rz1, dmsa_rz1, 1);
        ms_ss[j] = euler_mix(ms_ss[j], s_ss[j], dsa_ss1, dmsa_ss1, 0);
        s_rz[j] += dsa_rz;
        s_ss[j] += dsa_ss;
        ms_rz[j] = (s_rz[j] <= 0) ? 0 : ms_rz[j];
        ms_ss[j] = (s_ss[j] <= 0) ? 0 : ms_ss[j];
    }
#undef F
#undef O
#undef SP
}

/* ---- Runge-Kutta solver (settings.sas_solver = "RK4"), isotope branch: svat_transport_model_rk4 (transport.py:1138-2047).
 * Four stages of the Euler building blocks on stage copies of the StorAge; the stage updates are restated as written
 * (stage 1 advances by h, stages 2 and 3 by h / 2 with mass-like bookkeeping of the stage concentrations, stage 3 tests
 * sarkn - dsarkn < 0, dsarkn_ss1 of stage 1 lacks the flux), then tt = (k1 + 2 k2 + 2 k3 + k4) / 6 and one Euler-style update. */
static void rk4_substep_cell(const sas_ctx *c, int64_t C, int64_t i, double *s_rz, double *s_ss, double *ms_rz, double *ms_ss,
                             const double *forcing, double m, const double *sas_params, double *out, double *tt_base, size_t tt_stride,
                             double *work)
{
#define F(k) forcing[(size_t)(k) * (size_t)C + (size_t)i]
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
#define SP(p) (sas_params + ((size_t)(p) * (size_t)C + (size_t)i) * 8)
    const int A = c->A, N = A + 1;
    const double h = 1.0 / c->substeps;
    /* work: 2 SA + 4 stage arrays + 20 tt (5 fluxes x 4 stages) + 5 mtt + tmp + euler_tt scratch (4 N) */
    double *SAr = work, *SAs = SAr + N, *sr = SAs + N, *ss = sr + N, *mr = ss + N, *mss = mr + N;
    double *ttk = mss + N;       /* [stage][flux][N] */
    double *mttk = ttk + 20 * N; /* [flux][N] of the current stage */
    double *tmp = mttk + 5 * N, *w2 = tmp + N;
    const double inf_mat = F(RSF_inf_mat_rz), inf_pf = F(RSF_inf_pf_rz), inf_ss = F(RSF_inf_pf_ss), C_in = F(RSF_C_in);
    const double flux[5] = {F(RSF_evap_soil), F(RSF_transp), F(RSF_q_rz), F(RSF_cpr_rz), F(RSF_q_ss)};
    const int procs[5] = {RSP_evap_soil, RSP_transp, RSP_q_rz, RSP_cpr_rz, RSP_q_ss};
    const double evap = flux[0], transp = flux[1], q_rz = flux[2], cpr = flux[3], q_ss = flux[4];
    /* upper boundary condition :1146-1222 (the Euler solver's) */
    for (int j = 0; j < A; j++) {
        const double tt_im = (j == 0) ? ((inf_mat > 0) ? 1 : 0) * m : 0.0, tt_ip = (j == 0) ? ((inf_pf > 0) ? 1 : 0) * m : 0.0;
        const double tt_is = (j == 0) ? ((inf_ss > 0) ? 1 : 0) * m : 0.0;
        const double mtt_im = (j == 0) ? ((inf_mat > 0) ? C_in : 0) * m : 0.0, mtt_ip = (j == 0) ? ((inf_pf > 0) ? C_in : 0) * m : 0.0;
        const double mtt_is = (j == 0) ? ((inf_ss > 0) ? C_in : 0) * m : 0.0;
        const double dsa_rz = (inf_mat * tt_im + inf_pf * tt_ip) * h;
        const double dsa_ss = (inf_ss * tt_is) * h;
        const double dmsa_rz1 = nan0(mtt_im) * ((dsa_rz > 0) ? ((inf_mat * tt_im * h) / dsa_rz) : 0) +
                                nan0(mtt_ip) * ((dsa_rz > 0) ? ((inf_pf * tt_ip * h) / dsa_rz) : 0);
        const double dmsa_ss1 = nan0(mtt_is) * ((dsa_ss > 0) ? ((inf_ss * tt_is * h) / dsa_ss) : 0);
        ms_rz[j] = euler_mix(ms_rz[j], s_rz[j], dsa_rz, dmsa_rz1, 1);
        ms_ss[j] = euler_mix(ms_ss[j], s_ss[j], dsa_ss, dmsa_ss1, 0);
        s_rz[j] += dsa_rz;
        s_ss[j] += dsa_ss;
        ms_rz[j] = (s_rz[j] <= 0) ? 0 : ms_rz[j];
        ms_ss[j] = (s_ss[j] <= 0) ? 0 : ms_ss[j];
    }
    calc_SA(A, s_rz, SAr);
    for (int j = 0; j < N; j++) SAr[j] = SAr[j] * m;
    calc_SA(A, s_ss, SAs);
    for (int j = 0; j < N; j++) SAs[j] = SAs[j] * m;
    memcpy(sr, s_rz, sizeof(double) * A);
    memcpy(ss, s_ss, sizeof(double) * A);
    memcpy(mr, ms_rz, sizeof(double) * A);
    memcpy(mss, ms_ss, sizeof(double) * A);
#define TTK(st, f) (ttk + ((size_t)(st) * 5 + (size_t)(f)) * N)
#define MTT(f) (mttk + (size_t)(f) * N)
    for (int st = 0; st < 4; st++) {
        const int half = (st == 1 || st == 2);
        for (int f = 0; f < 5; f++) {
            const double fh = half ? flux[f] * h / 2 : flux[f] * h;
            euler_tt(c, (f < 3) ? SAr : SAs, SP(procs[f]), fh, m, TTK(st, f), w2);
            const double *msrc = (f < 3) ? mr : mss;
            for (int j = 0; j < A; j++) MTT(f)[j] = (TTK(st, f)[j] > 0) ? msrc[j] : 0;
        }
        const double *tt_e = TTK(st, 0), *tt_t = TTK(st, 1), *tt_q = TTK(st, 2), *tt_c = TTK(st, 3), *tt_s = TTK(st, 4);
        if (st == 0) { /* :1388-1431 */
            for (int j = 0; j < A; j++) {
                double dsa_rz = (cpr * tt_c[j] - evap * tt_e[j] - transp * tt_t[j] - q_rz * tt_q[j]) * h;
                dsa_rz = (sr[j] + dsa_rz < 0) ? -sr[j] : dsa_rz;
                double dsa_ss = (q_rz * tt_q[j] - cpr * tt_c[j] - q_ss * tt_s[j]) * h;
                dsa_ss = (ss[j] + dsa_ss < 0) ? -ss[j] : dsa_ss;
                const double dsa_rz1 = (cpr * tt_c[j]) * h;
                const double dmsa_rz1 = nan0(MTT(3)[j]) * ((dsa_rz1 > 0) ? ((cpr * tt_c[j] * h) / dsa_rz1) : 0);
                const double dsa_ss1 = (tt_q[j]) * h;
                const double dmsa_ss1 = nan0(MTT(2)[j]) * ((dsa_ss1 > 0) ? ((q_rz * tt_q[j] * h) / dsa_ss1) : 0);
                mr[j] = euler_mix(mr[j], sr[j], dsa_rz1, dmsa_rz1, 1);
                mss[j] = euler_mix(mss[j], ss[j], dsa_ss1, dmsa_ss1, 0);
                sr[j] += dsa_rz;
                ss[j] += dsa_ss;
                mr[j] = (sr[j] <= 0) ? 0 : mr[j];
                mss[j] = (ss[j] <= 0) ? 0 : mss[j];
            }
        } else if (st == 1 || st == 2) { /* :1551-1581, :1699-1729 */
            for (int j = 0; j < A; j++) {
                double dsa_rz = (cpr * tt_c[j] - evap * tt_e[j] - transp * tt_t[j] - q_rz * tt_q[j]) * h / 2;
                double dsa_ss = (q_rz * tt_q[j] - cpr * tt_c[j] - q_ss * tt_s[j]) * h / 2;
                if (st == 1) {
                    dsa_rz = (sr[j] + dsa_rz < 0) ? -sr[j] : dsa_rz;
                    dsa_ss = (ss[j] + dsa_ss < 0) ? -ss[j] : dsa_ss;
                } else {
                    dsa_rz = (sr[j] - dsa_rz < 0) ? -sr[j] : dsa_rz;
                    dsa_ss = (ss[j] - dsa_ss < 0) ? -ss[j] : dsa_ss;
                }
                sr[j] += dsa_rz;
                ss[j] += dsa_ss;
                const double r1 = nan0(MTT(3)[j]);
                const double r2 = nan0(MTT(0)[j]) + nan0(MTT(1)[j]) + nan0(MTT(2)[j]);
                double s1 = nan0(MTT(2)[j]);
                double s2 = nan0(MTT(3)[j]) + nan0(MTT(4)[j]);
                if (st == 1) { s1 = s1 * h / 2; s2 = s2 * h / 2; }
                double dr = ((r1 < 0) && (r2 >= 0)) ? r1 + r2 : r1 - r2;
                double ds = ((s1 < 0) && (s2 >= 0)) ? s1 + s2 : s1 - s2;
                dr = (dr > 0) ? ((mr[j] + dr > 0) ? -mr[j] : dr) : dr;
                ds = (ds > 0) ? ((mss[j] + ds > 0) ? -mss[j] : ds) : ds;
                mr[j] += dr;
                mss[j] += ds;
            }
        }
        if (st < 3) { /* SArkn[1, 1:] = cumsum(sarkn[1]) */
            double acc = sr[0];
            SAr[1] = acc;
            for (int j = 1; j < A; j++) { acc += sr[j]; SAr[j + 1] = acc; }
            acc = ss[0];
            SAs[1] = acc;
            for (int j = 1; j < A; j++) { acc += ss[j]; SAs[j + 1] = acc; }
        }
    }
    /* combination :1836-1856 into stage slot 0, solute travel-time distributions :1879-1897 */
    for (int f = 0; f < 5; f++) {
        double *t0 = TTK(0, f);
        const double *t1 = TTK(1, f), *t2 = TTK(2, f), *t3 = TTK(3, f);
        for (int j = 0; j < A; j++) t0[j] = (t0[j] + 2 * t1[j] + 2 * t2[j] + t3[j]) / 6.;
        const double *msrc = (f < 3) ? ms_rz : ms_ss;
        for (int j = 0; j < A; j++) MTT(f)[j] = (t0[j] > 0) ? msrc[j] : 0;
        if (tt_base) memcpy(tt_base + (size_t)procs[f] * tt_stride, t0, sizeof(double) * A);
    }
    {
        const int oc[5] = {RSO_C_evap_soil, RSO_C_transp, RSO_C_q_rz, RSO_C_cpr_rz, RSO_C_q_ss};
        const int oi[5] = {RSO_C_iso_evap_soil, RSO_C_iso_transp, RSO_C_iso_q_rz, RSO_C_iso_cpr_rz, RSO_C_iso_q_ss};
        for (int f = 0; f < 5; f++) {
            O(oc[f]) = euler_flux_conc(A, MTT(f), TTK(0, f), tmp) * m;
            O(oi[f]) = conc_to_delta(c, O(oc[f])) * m;
        }
        O(RSO_C_inf_mat_rz) = ((inf_mat > 0) ? C_in : NAN) * m;
        O(RSO_C_inf_pf_rz) = ((inf_pf > 0) ? C_in : NAN) * m;
        O(RSO_C_inf_pf_ss) = ((inf_ss > 0) ? C_in : NAN) * m;
        O(RSO_C_iso_inf_mat_rz) = conc_to_delta(c, O(RSO_C_inf_mat_rz)) * m;
        O(RSO_C_iso_inf_pf_rz) = conc_to_delta(c, O(RSO_C_inf_pf_rz)) * m;
        O(RSO_C_iso_inf_pf_ss) = conc_to_delta(c, O(RSO_C_inf_pf_ss)) * m;
    }
    { /* update :1899-1945 (the Euler solver's) */
        const double *tt_e = TTK(0, 0), *tt_t = TTK(0, 1), *tt_q = TTK(0, 2), *tt_c = TTK(0, 3), *tt_s = TTK(0, 4);
        for (int j = 0; j < A; j++) {
            const double mtt_c = MTT(3)[j], mtt_q = MTT(2)[j];
            double dsa_rz = (cpr * tt_c[j] - evap * tt_e[j] - transp * tt_t[j] - q_rz * tt_q[j]) * h;
            dsa_rz = (s_rz[j] + dsa_rz < 0) ? -s_rz[j] : dsa_rz;
            double dsa_ss = (q_rz * tt_q[j] - cpr * tt_c[j] - q_ss * tt_s[j]) * h;
            dsa_ss = (s_ss[j] + dsa_ss < 0) ? -s_ss[j] : dsa_ss;
            const double dsa_rz1 = (cpr * tt_c[j]) * h;
            const double dmsa_rz1 = nan0(mtt_c) * ((dsa_rz1 > 0) ? ((cpr * tt_c[j] * h) / dsa_rz1) : 0);
            const double dsa_ss1 = (q_rz * tt_q[j]) * h;
            const double dmsa_ss1 = nan0(mtt_q) * ((dsa_ss1 > 0) ? ((q_rz * tt_q[j] * h) / dsa_ss1) : 0);
            ms_rz[j] = euler_mix(ms_rz[j], s_rz[j], dsa_rz1, dmsa_rz1, 1);
            ms_ss[j] = euler_mix(ms_ss[j], s_ss[j], dsa_ss1, dmsa_ss1, 0);
            s_rz[j] += dsa_rz;
            s_ss[j] += dsa_ss;
            ms_rz[j] = (s_rz[j] <= 0) ? 0 : ms_rz[j];
            ms_ss[j] = (s_ss[j] <= 0) ? 0 : ms_ss[j];
        }
    }
#undef TTK
#undef MTT
#undef F
#undef O
#undef SP
}

/* one explicit Euler step, anion branch (bromide / chloride; transport.py:2144-2156, :2320-2343, :2378-2407): msa holds
 * solute mass per bin; calc_mtt (:582-596) without the deterministic kernels' crop switch */
static void euler_substep_cell_anion(const sas_ctx *c, int64_t C, int64_t i, double *s_rz, double *s_ss, double *ms_rz, double *ms_ss,
                                     const double *forcing, double m, const double *sas_params, double *out, double *tt_base,
                                     size_t tt_stride, double *work)
{
#define F(k) forcing[(size_t)(k) * (size_t)C + (size_t)i]
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
#define SP(p) (sas_params + ((size_t)(p) * (size_t)C + (size_t)i) * 8)
    const int A = c->A, N = A + 1;
    const double h = 1.0 / c->substeps;
    double *SA_rz = work, *SA_ss = SA_rz + N, *tt_e = SA_ss + N, *tt_t = tt_e + N, *tt_q = tt_t + N, *tt_c = tt_q + N, *tt_s = tt_c + N;
    double *mtt_t = tt_s + N, *mtt_q = mtt_t + N, *mtt_c = mtt_q + N, *mtt_s = mtt_c + N, *w2 = mtt_s + N;
    const double inf_mat = F(RSF_inf_mat_rz), inf_pf = F(RSF_inf_pf_rz), inf_ss = F(RSF_inf_pf_ss), C_in = F(RSF_C_in);
    const double evap = F(RSF_evap_soil), transp = F(RSF_transp), q_rz = F(RSF_q_rz), q_ss = F(RSF_q_ss), cpr = F(RSF_cpr_rz);
    const double alpha_t = g_anion[0] ? g_anion[0][i] : 1.0, alpha_q = g_anion[1] ? g_anion[1][i] : 1.0;
    /* upper boundary condition: water by the infiltration, solute as mtt_inf * h with mtt_inf[0] = C_in * mask (:2089-2101) */
    for (int j = 0; j < A; j++) {
        const double tt_im = (j == 0) ? ((inf_mat > 0) ? 1 : 0) * m : 0.0, tt_ip = (j == 0) ? ((inf_pf > 0) ? 1 : 0) * m : 0.0;
        const double tt_is = (j == 0) ? ((inf_ss > 0) ? 1 : 0) * m : 0.0;
        const double mtt_i = (j == 0) ? C_in * m : 0.0; /* the same value in all three infiltration arrays */
        s_rz[j] += (inf_mat * tt_im + inf_pf * tt_ip) * h;
        s_ss[j] += (inf_ss * tt_is) * h;
        ms_rz[j] += nan0(mtt_i) * h + nan0(mtt_i) * h;
        ms_ss[j] += nan0(mtt_i) * h;
    }
    calc_SA(A, s_rz, SA_rz);
    for (int j = 0; j < N; j++) SA_rz[j] = SA_rz[j] * m;
    calc_SA(A, s_ss, SA_ss);
    for (int j = 0; j < N; j++) SA_ss[j] = SA_ss[j] * m;
    euler_tt(c, SA_rz, SP(RSP_evap_soil), evap * h, m, tt_e, w2);
    euler_tt(c, SA_rz, SP(RSP_transp), transp * h, m, tt_t, w2);
    euler_tt(c, SA_rz, SP(RSP_q_rz), q_rz * h, m, tt_q, w2);
    euler_tt(c, SA_ss, SP(RSP_cpr_rz), cpr * h, m, tt_c, w2);
    euler_tt(c, SA_ss, SP(RSP_q_ss), q_ss * h, m, tt_s, w2);
    anion_mtt(A, s_rz, tt_t, transp * h, ms_rz, alpha_t, 1.0, mtt_t);
    anion_mtt(A, s_rz, tt_q, q_rz * h, ms_rz, alpha_q, 1.0, mtt_q);
    anion_mtt(A, s_ss, tt_c, cpr * h, ms_ss, alpha_q, 1.0, mtt_c);
    anion_mtt(A, s_ss, tt_s, q_ss * h, ms_ss, alpha_q, 1.0, mtt_s);
    if (tt_base) {
        const double *tts[5] = {tt_e, tt_t, tt_q, tt_c, tt_s};
        const int pr[5] = {RSP_evap_soil, RSP_transp, RSP_q_rz, RSP_cpr_rz, RSP_q_ss};
        for (int k = 0; k < 5; k++) memcpy(tt_base + (size_t)pr[k] * tt_stride, tts[k], sizeof(double) * A);
    }
    /* concentrations :2378-2407 from the solute travel-time distributions of this step */
    O(RSO_C_inf_mat_rz) = ((inf_mat * h > 0) ? C_in : 0) * m;
    O(RSO_C_inf_pf_rz) = ((inf_pf * h > 0) ? C_in : 0) * m;
    O(RSO_C_inf_pf_ss) = ((inf_ss * h > 0) ? C_in : 0) * m;
    O(RSO_C_transp) = ((transp > 0) ? np_sum(mtt_t, A) / (transp * h) : 0) * m;
    O(RSO_C_q_rz) = ((q_rz > 0) ? np_sum(mtt_q, A) / (q_rz * h) : 0) * m;
    O(RSO_C_cpr_rz) = ((cpr > 0) ? np_sum(mtt_c, A) / (cpr * h) : 0) * m;
    O(RSO_C_q_ss) = ((q_ss > 0) ? np_sum(mtt_s, A) / (q_ss * h) : 0) * m;
    /* update :2320-2343 */
    for (int j = 0; j < A; j++) {
        double dsa_rz = (cpr * tt_c[j] - evap * tt_e[j] - transp * tt_t[j] - q_rz * tt_q[j]) * h;
        dsa_rz = (s_rz[j] + dsa_rz < 0) ? -s_rz[j] : dsa_rz;
        double dsa_ss = (q_rz * tt_q[j] - cpr * tt_c[j] - q_ss * tt_s[j]) * h;
        dsa_ss = (s_ss[j] + dsa_ss < 0) ? -s_ss[j] : dsa_ss;
        s_rz[j] += dsa_rz;
        s_ss[j] += dsa_ss;
        double dmsa_rz = nan0(mtt_c[j]) - nan0(mtt_t[j]) - nan0(mtt_q[j]);
        double dmsa_ss = nan0(mtt_q[j]) - nan0(mtt_c[j]) - nan0(mtt_s[j]);
        dmsa_rz = (ms_rz[j] + dmsa_rz < 0) ? 0 : dmsa_rz;
        dmsa_ss = (ms_ss[j] + dmsa_ss < 0) ? 0 : dmsa_ss;
        ms_rz[j] += dmsa_rz;
        ms_ss[j] += dmsa_ss;
    }
#undef F
#undef O
#undef SP
}

/* anion storage kernels (root_zone.py:219-258, subsoil.py, soil.py anion kernels) */
static void storage_kernels_anion(const sas_ctx *c, int64_t C, int64_t i, double *s_rz, double *s_ss, double *ms_rz, double *ms_ss, double m,
                                  double *out, double *tmp)
{
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
    const int A = c->A;
    for (int j = 0; j < A; j++) s_rz[j] = (s_rz[j] < 1e-8) ? 0 : s_rz[j];
    for (int j = 0; j < A; j++) ms_rz[j] = (s_rz[j] <= 0) ? 0 : ms_rz[j];
    {
        const double M = np_nansum(ms_rz, A, tmp) * m, S = np_sum(s_rz, A);
        O(RSO_C_iso_rz) = M;
        O(RSO_C_rz) = (S > 0) ? M / S : 0;
    }
    for (int j = 0; j < A; j++) s_ss[j] = (s_ss[j] < 1e-8) ? 0 : s_ss[j];
    for (int j = 0; j < A; j++) ms_ss[j] = (s_ss[j] <= 0) ? 0 : ms_ss[j];
    {
        const double M = np_nansum(ms_ss, A, tmp) * m, S = np_sum(s_ss, A);
        O(RSO_C_iso_ss) = M;
        O(RSO_C_ss) = (S > 0) ? M / S : 0;
    }
    double *sa_s = tmp + (A + 2), *msa_s = sa_s + (A + 2);
    for (int j = 0; j < A; j++) {
        sa_s[j] = s_rz[j] + s_ss[j] * m;
        msa_s[j] = ms_rz[j] + ms_ss[j] * m;
    }
    const double M = np_nansum(msa_s, A, tmp) * m, S = np_sum(sa_s, A);
    O(RSO_C_iso_s) = M;
    O(RSO_C_s) = (S > 0) ? M / S : 0;
#undef O
}

/* storage kernels root_zone.py:189-215, subsoil.py:153-180, soil.py:1036-1092 (run after every Euler sub-step) */
static void storage_kernels_iso(const sas_ctx *c, int64_t C, int64_t i, double *s_rz, double *s_ss, const double *ms_rz, const double *ms_ss,
                                double m, double *out, double *tmp)
{
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
    const int A = c->A;
    for (int j = 0; j < A; j++) s_rz[j] = (s_rz[j] < 1e-8) ? 0 : s_rz[j];
    O(RSO_C_rz) = storage_conc(c, s_rz, ms_rz, tmp) * m;
    O(RSO_C_iso_rz) = conc_to_delta(c, O(RSO_C_rz)) * m;
    for (int j = 0; j < A; j++) s_ss[j] = (s_ss[j] < 1e-8) ? 0 : s_ss[j];
    O(RSO_C_ss) = storage_conc(c, s_ss, ms_ss, tmp) * m;
    O(RSO_C_iso_ss) = conc_to_delta(c, O(RSO_C_ss)) * m;
    double *sa_s = tmp + (A + 2), *msa_s = sa_s + (A + 2);
    for (int j = 0; j < A; j++) {
        sa_s[j] = s_rz[j] + s_ss[j] * m;
        const double tot = s_rz[j] + s_ss[j];
        const double v = (tot > 0) ? ms_rz[j] * (s_rz[j] / tot) + ms_ss[j] * (s_ss[j] / tot) : 0;
        msa_s[j] = isnan(v) ? 0 : v;
    }
    O(RSO_C_s) = storage_conc(c, sa_s, msa_s, tmp) * m;
    O(RSO_C_iso_s) = conc_to_delta(c, O(RSO_C_s)) * m;
#undef O
}

/* one day = `substeps` explicit Euler steps of every cell, ageing with the last one (transport.py:3221-3258) */
static void euler_day(int64_t C, const roger_sas_settings *st, double *sa_rz, double *sa_ss, double *msa_rz, double *msa_ss,
                      const double *forcing, const double *mask, const double *sas_params, double *out, double *tt_out,
                      double *stats, double *balance_prev)
{
    sas_ctx ctx = {st->ages, st->substeps, st->VSMOW_conc18O, st->d18O_min, st->d18O_max};
    const int A = ctx.A;
    g_tracer = (st->solute == 3);
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * (size_t)(16 * (A + 2)));
        double *tmp = (double *)malloc(sizeof(double) * (size_t)(4 * (A + 2)));
        double *ttk = (double *)malloc(sizeof(double) * (size_t)(RSP_COUNT * A)); /* tt of the last Euler step (statistics) */
        double *rkwork = (double *)malloc(sizeof(double) * (size_t)(40 * (A + 2)));
        double *ages = (double *)malloc(sizeof(double) * (size_t)A);
        for (int j = 0; j < A; j++) ages[j] = (double)(j + 1);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < C; i++) {
            const double m = mask[i];
            double *s_rz = sa_rz + (size_t)i * A, *s_ss = sa_ss + (size_t)i * A;
            double *ms_rz = msa_rz + (size_t)i * A, *ms_ss = msa_ss + (size_t)i * A;
#define F(k) forcing[(size_t)(k) * (size_t)C + (size_t)i]
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
#define ST(k) stats[(size_t)(k) * (size_t)C + (size_t)i]
            double sum_sa_s_m1 = 0, Cp = 0;
            for (int s = 0; s < ctx.substeps; s++) {
                const int last = (s == ctx.substeps - 1);
                if (stats && last) { /* [taum1] of the last Euler step: the state its predecessor left (after_substep_iso :3421) */
                    for (int j = 0; j < A; j++) tmp[j] = s_rz[j] + s_ss[j] * m;
                    sum_sa_s_m1 = np_sum(tmp, A);
                    Cp = (ctx.substeps > 1) ? O(RSO_C_s) : balance_prev[i];
                }
                if (st->solute == 1 || st->solute == 3) { /* the virtual tracer runs the same anion branch (transport.py:2308): identical */
                    euler_substep_cell_anion(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, forcing, m, sas_params, out,
                                             tt_out ? tt_out + (size_t)i * A : ((stats && last) ? ttk : NULL), tt_out ? (size_t)C * A : (size_t)A, work);
                    storage_kernels_anion(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, m, out, tmp);
                    continue;
                }
                if (st->solver == 2) {
                    rk4_substep_cell(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, forcing, m, sas_params, out,
                                     tt_out ? tt_out + (size_t)i * A : ((stats && last) ? ttk : NULL), tt_out ? (size_t)C * A : (size_t)A, rkwork);
                    storage_kernels_iso(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, m, out, tmp);
                    continue;
                }
                if (tt_out) euler_substep_cell(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, forcing, m, sas_params, out, tt_out + (size_t)i * A, (size_t)C * A, work);
                else euler_substep_cell(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, forcing, m, sas_params, out, (stats && last) ? ttk : NULL, (size_t)A, work);
                storage_kernels_iso(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, m, out, tmp);
            }
            if (st->solute == 1 || st->solute == 3) {
                if (stats) { /* age statistics (water only, transport.py:59-312) and the balance errors of the day's last Euler step:
                              * numerics.py:424-445 (water), :642-672 (bromide / chloride), both carrying settings.h */
                    const double h = 1.0 / ctx.substeps;
                    const double *TT[RSP_COUNT];
                    for (int p = 0; p < RSP_COUNT; p++) TT[p] = tt_out ? tt_out + ((size_t)p * (size_t)C + (size_t)i) * A : ttk + (size_t)p * A;
                    const double p5[5] = {0.1, 0.25, 0.5, 0.75, 0.9}, p3[3] = {0.25, 0.5, 0.75};
                    double s5[5], s3[3], avg, w[3];
                    double *sa_s = tmp + (A + 2);
                    for (int j = 0; j < A; j++) sa_s[j] = s_rz[j] + s_ss[j] * m;
                    const int procs[3] = {RSP_evap_soil, RSP_transp, RSP_q_ss};
                    const double flx[3] = {F(RSF_evap_soil), F(RSF_transp), F(RSF_q_ss)};
                    for (int k = 0; k < 3; k++) {
                        for (int j = 0; j < A; j++) tmp[j] = flx[k] * TT[procs[k]][j];
                        w[k] = np_sum(tmp, A);
                    }
                    flux_age_statistics(A, TT[RSP_transp], ages, work, s5, &avg);
                    for (int k = 0; k < 5; k++) ST(RSS_tt10_transp + k) = s5[k];
                    ST(RSS_ttavg_transp) = avg;
                    flux_age_statistics(A, TT[RSP_q_ss], ages, work, s5, &avg);
                    for (int k = 0; k < 5; k++) ST(RSS_tt10_q_ss + k) = s5[k];
                    ST(RSS_ttavg_q_ss) = avg;
                    storage_age_statistics(A, s_rz, m, ages, work, p3, 3, s3, &avg);
                    for (int k = 0; k < 3; k++) ST(RSS_rt25_rz + k) = s3[k];
                    ST(RSS_rtavg_rz) = avg;
                    storage_age_statistics(A, s_ss, m, ages, work, p3, 3, s3, &avg);
                    for (int k = 0; k < 3; k++) ST(RSS_rt25_ss + k) = s3[k];
                    ST(RSS_rtavg_ss) = avg;
                    storage_age_statistics(A, sa_s, m, ages, work, p5, 5, s5, &avg);
                    for (int k = 0; k < 5; k++) ST(RSS_rt10_s + k) = s5[k];
                    ST(RSS_rtavg_s) = avg;
                    const double S = np_sum(sa_s, A);
                    ST(RSS_dS_num_error) = fabs((S - sum_sa_s_m1) - (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - w[0] - w[1] - w[2]) * h);
                    ST(RSS_dC_num_error) = fabs((S * O(RSO_C_s) - sum_sa_s_m1 * Cp) -
                                                (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                                 F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - w[1] * nan0(O(RSO_C_transp)) - w[2] * nan0(O(RSO_C_q_ss))) * h);
                    ST(RSS_sanity_ok) = NAN; /* the anion branch of numerics.sanity_check is not restated */
                    balance_prev[i] = O(RSO_C_s);
                }
                anion_ageing(A, s_rz, ms_rz);
                anion_ageing(A, s_ss, ms_ss);
                continue;
            }
            if (stats) { /* age statistics (transport.py:59-312) and balance errors (numerics.py:424-445, :476-516, :1051-1122) of the
                          * last Euler step; the right-hand sides carry settings.h */
                const double h = 1.0 / ctx.substeps;
                const double *TT[RSP_COUNT];
                for (int p = 0; p < RSP_COUNT; p++) TT[p] = tt_out ? tt_out + ((size_t)p * (size_t)C + (size_t)i) * A : ttk + (size_t)p * A;
                const double p5[5] = {0.1, 0.25, 0.5, 0.75, 0.9}, p3[3] = {0.25, 0.5, 0.75};
                double s5[5], s3[3], avg;
                double *sa_s = tmp + (A + 2);
                for (int j = 0; j < A; j++) sa_s[j] = s_rz[j] + s_ss[j] * m;
                flux_age_statistics(A, TT[RSP_transp], ages, work, s5, &avg);
                for (int k = 0; k < 5; k++) ST(RSS_tt10_transp + k) = s5[k];
                ST(RSS_ttavg_transp) = avg;
                flux_age_statistics(A, TT[RSP_q_ss], ages, work, s5, &avg);
                for (int k = 0; k < 5; k++) ST(RSS_tt10_q_ss + k) = s5[k];
                ST(RSS_ttavg_q_ss) = avg;
                storage_age_statistics(A, s_rz, m, ages, work, p3, 3, s3, &avg);
                for (int k = 0; k < 3; k++) ST(RSS_rt25_rz + k) = s3[k];
                ST(RSS_rtavg_rz) = avg;
                storage_age_statistics(A, s_ss, m, ages, work, p3, 3, s3, &avg);
                for (int k = 0; k < 3; k++) ST(RSS_rt25_ss + k) = s3[k];
                ST(RSS_rtavg_ss) = avg;
                storage_age_statistics(A, sa_s, m, ages, work, p5, 5, s5, &avg);
                for (int k = 0; k < 5; k++) ST(RSS_rt10_s + k) = s5[k];
                ST(RSS_rtavg_s) = avg;
                const double sum_s = np_sum(sa_s, A);
                double w[3];
                const int procs[3] = {RSP_evap_soil, RSP_transp, RSP_q_ss};
                const double flx[3] = {F(RSF_evap_soil), F(RSF_transp), F(RSF_q_ss)};
                for (int k = 0; k < 3; k++) {
                    for (int j = 0; j < A; j++) tmp[j] = flx[k] * TT[procs[k]][j];
                    w[k] = np_sum(tmp, A);
                }
                const double lhs1 = sum_s - sum_sa_s_m1;
                const double rhs1 = (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - w[0] - w[1] - w[2]) * h;
                const double lhs2 = sum_s * O(RSO_C_s) - sum_sa_s_m1 * Cp;
                const double rhs2 = (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                     F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - w[0] * nan0(O(RSO_C_evap_soil)) -
                                     w[1] * nan0(O(RSO_C_transp)) - w[2] * nan0(O(RSO_C_q_ss))) * h;
                ST(RSS_dS_num_error) = fabs(lhs1 - rhs1);
                ST(RSS_dC_num_error) = fabs(lhs2 - rhs2);
                const int c1 = (m != 0) ? np_isclose(lhs1, rhs1, g_rtol, g_atol) : 1;
                const int c2 = (m != 0) ? np_isclose(lhs2, rhs2, g_rtol, g_atol) : 1;
                const double sat_rz = g_bounds[0] ? g_bounds[0][i] : INFINITY, sat_ss = g_bounds[1] ? g_bounds[1][i] : INFINITY;
                const double pwp_rz = g_bounds[2] ? g_bounds[2][i] : 0.0, pwp_ss = g_bounds[3] ? g_bounds[3][i] : 0.0;
                const double sum_rz = np_sum(s_rz, A), sum_ss = np_sum(s_ss, A);
                const int b3 = (sum_s <= (sat_rz + sat_ss) - (pwp_rz + pwp_ss)) && (sum_s >= 0);
                const int b4 = (sum_rz <= sat_rz - pwp_rz) && (sum_rz >= 0);
                const int b5 = (sum_ss <= sat_ss - pwp_ss) && (sum_ss >= 0);
                ST(RSS_sanity_ok) = (c1 && c2 && b3 && b4 && b5) ? 1.0 : 0.0;
                balance_prev[i] = O(RSO_C_s);
            }
#undef F
#undef O
#undef ST
            ageing(A, s_rz, ms_rz);
            ageing(A, s_ss, ms_ss);
        }
        free(work);
        free(tmp);
        free(ttk);
        free(rkwork);
        free(ages);
    }
}

void sas_oracle_day_stats(int64_t C, const roger_sas_settings *st, double *sa_rz, double *sa_ss, double *msa_rz, double *msa_ss,
                          const double *forcing, const double *mask, const double *sas_params, double *out, double *tt_out,
                          double *stats, double *balance_prev);

void sas_oracle_day(int64_t C, const roger_sas_settings *st, double *sa_rz, double *sa_ss, double *msa_rz, double *msa_ss,
                    const double *forcing, const double *mask, const double *sas_params, double *out, double *tt_out)
{
    sas_oracle_day_stats(C, st, sa_rz, sa_ss, msa_rz, msa_ss, forcing, mask, sas_params, out, tt_out, NULL, NULL);
}

/* stats: RSS_COUNT planes (or NULL); balance_prev: C_s of the previous step per cell (in / out) */
void sas_oracle_day_stats(int64_t C, const roger_sas_settings *st, double *sa_rz, double *sa_ss, double *msa_rz, double *msa_ss,
                          const double *forcing, const double *mask, const double *sas_params, double *out, double *tt_out,
                          double *stats, double *balance_prev)
{
    if (st->solver == 1 || st->solver == 2) { /* explicit Euler / RK4 (callers check the supported combinations) */
        euler_day(C, st, sa_rz, sa_ss, msa_rz, msa_ss, forcing, mask, sas_params, out, tt_out, stats, balance_prev);
        return;
    }
    sas_ctx ctx = {st->ages, st->substeps, st->VSMOW_conc18O, st->d18O_min, st->d18O_max};
    const int A = ctx.A;
    g_tracer = (st->solute == 3);
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * (size_t)(16 * (A + 2)));
        double *tmp = (double *)malloc(sizeof(double) * (size_t)(4 * (A + 2)));
        double *ttk = (double *)malloc(sizeof(double) * (size_t)(3 * (A + 2))); /* tt of evap, transp, q_ss (statistics) */
        double *ages = (double *)malloc(sizeof(double) * (size_t)A);
        for (int j = 0; j < A; j++) ages[j] = (double)(j + 1);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < C; i++) {
            const double m = mask[i];
            double *s_rz = sa_rz + (size_t)i * A, *s_ss = sa_ss + (size_t)i * A;
            double *ms_rz = msa_rz + (size_t)i * A, *ms_ss = msa_ss + (size_t)i * A;
            if (st->solute >= 1 && st->solute <= 3) {
                anion_day_cell(&ctx, C, i, s_rz, s_ss, ms_rz, ms_ss, forcing, m, sas_params, out, tt_out, work, tmp, stats, balance_prev, ttk, ages);
                continue;
            }
#define F(k) forcing[(size_t)(k) * (size_t)C + (size_t)i]
#define O(k) out[(size_t)(k) * (size_t)C + (size_t)i]
#define SP(p) (sas_params + ((size_t)(p) * (size_t)C + (size_t)i) * 8)
#define TT(p) (tt_out ? tt_out + ((size_t)(p) * (size_t)C + (size_t)i) * A : \
               (stats ? ((p) == RSP_evap_soil ? ttk : (p) == RSP_transp ? ttk + (A + 2) : (p) == RSP_q_ss ? ttk + 2 * (A + 2) : NULL) : NULL))
#define ST(k) stats[(size_t)(k) * (size_t)C + (size_t)i]
            const double C_in = F(RSF_C_in);
            /* vs.sa_s[taum1] as the soil kernel of THIS step recomputes it (soil.py:1044-1048 writes both time levels):
             * the aged StorAge of both layers as the step finds them */
            double sum_sa_s_m1 = 0;
            if (stats) {
                for (int j = 0; j < A; j++) tmp[j] = s_rz[j] + s_ss[j] * m;
                sum_sa_s_m1 = np_sum(tmp, A);
            }
            /* infiltration into the root zone: matrix, then preferential flow */
            inflow(&ctx, s_rz, ms_rz, F(RSF_inf_mat_rz), C_in, m, &O(RSO_C_inf_mat_rz), &O(RSO_C_iso_inf_mat_rz));
            inflow(&ctx, s_rz, ms_rz, F(RSF_inf_pf_rz), C_in, m, &O(RSO_C_inf_pf_rz), &O(RSO_C_iso_inf_pf_rz));
            withdrawal(&ctx, s_rz, ms_rz, NULL, NULL, F(RSF_evap_soil), SP(RSP_evap_soil), m, &O(RSO_C_evap_soil), &O(RSO_C_iso_evap_soil), TT(RSP_evap_soil), work);
            withdrawal(&ctx, s_rz, ms_rz, NULL, NULL, F(RSF_transp), SP(RSP_transp), m, &O(RSO_C_transp), &O(RSO_C_iso_transp), TT(RSP_transp), work);
            withdrawal(&ctx, s_rz, ms_rz, s_ss, ms_ss, F(RSF_q_rz), SP(RSP_q_rz), m, &O(RSO_C_q_rz), &O(RSO_C_iso_q_rz), TT(RSP_q_rz), work);
            inflow(&ctx, s_ss, ms_ss, F(RSF_inf_pf_ss), C_in, m, &O(RSO_C_inf_pf_ss), &O(RSO_C_iso_inf_pf_ss));
            withdrawal(&ctx, s_ss, ms_ss, NULL, NULL, F(RSF_q_ss), SP(RSP_q_ss), m, &O(RSO_C_q_ss), &O(RSO_C_iso_q_ss), TT(RSP_q_ss), work);
            withdrawal(&ctx, s_ss, ms_ss, s_rz, ms_rz, F(RSF_cpr_rz), SP(RSP_cpr_rz), m, &O(RSO_C_cpr_rz), &O(RSO_C_iso_cpr_rz), TT(RSP_cpr_rz), work);
            /* storage kernels root_zone.py:189-215, subsoil.py:153-180, soil.py:1036-1092 */
            for (int j = 0; j < A; j++) s_rz[j] = (s_rz[j] < 1e-8) ? 0 : s_rz[j];
            O(RSO_C_rz) = storage_conc(&ctx, s_rz, ms_rz, tmp) * m;
            O(RSO_C_iso_rz) = conc_to_delta(&ctx, O(RSO_C_rz)) * m;
            for (int j = 0; j < A; j++) s_ss[j] = (s_ss[j] < 1e-8) ? 0 : s_ss[j];
            O(RSO_C_ss) = storage_conc(&ctx, s_ss, ms_ss, tmp) * m;
            O(RSO_C_iso_ss) = conc_to_delta(&ctx, O(RSO_C_ss)) * m;
            {
                double *sa_s = tmp + (A + 2), *msa_s = sa_s + (A + 2);
                for (int j = 0; j < A; j++) {
                    sa_s[j] = s_rz[j] + s_ss[j] * m;
                    double tot = s_rz[j] + s_ss[j];
                    double v = (tot > 0) ? ms_rz[j] * (s_rz[j] / tot) + ms_ss[j] * (s_ss[j] / tot) : 0;
                    msa_s[j] = isnan(v) ? 0 : v;
                }
                O(RSO_C_s) = storage_conc(&ctx, sa_s, msa_s, tmp) * m;
                O(RSO_C_iso_s) = conc_to_delta(&ctx, O(RSO_C_s)) * m;
                if (stats) {
                    const double p5[5] = {0.1, 0.25, 0.5, 0.75, 0.9}, p3[3] = {0.25, 0.5, 0.75};
                    /* the five percentiles of a group are RSS-consecutive but plane-strided: write them one by one */
                    double s5[5], s3[3], avg;
                    flux_age_statistics(A, TT(RSP_transp), ages, work, s5, &avg);
                    for (int k = 0; k < 5; k++) ST(RSS_tt10_transp + k) = s5[k];
                    ST(RSS_ttavg_transp) = avg;
                    flux_age_statistics(A, TT(RSP_q_ss), ages, work, s5, &avg);
                    for (int k = 0; k < 5; k++) ST(RSS_tt10_q_ss + k) = s5[k];
                    ST(RSS_ttavg_q_ss) = avg;
                    storage_age_statistics(A, s_rz, m, ages, work, p3, 3, s3, &avg);
                    for (int k = 0; k < 3; k++) ST(RSS_rt25_rz + k) = s3[k];
                    ST(RSS_rtavg_rz) = avg;
                    storage_age_statistics(A, s_ss, m, ages, work, p3, 3, s3, &avg);
                    for (int k = 0; k < 3; k++) ST(RSS_rt25_ss + k) = s3[k];
                    ST(RSS_rtavg_ss) = avg;
                    storage_age_statistics(A, sa_s, m, ages, work, p5, 5, s5, &avg);
                    for (int k = 0; k < 5; k++) ST(RSS_rt10_s + k) = s5[k];
                    ST(RSS_rtavg_s) = avg;
                    /* numerics.py:424-445 and :476-516 (settings.h = 1) */
                    double *bp_C = balance_prev + i; /* vs.C_s[taum1] */
                    const double sum_s = np_sum(sa_s, A);
                    double w[3];
                    const int procs[3] = {RSP_evap_soil, RSP_transp, RSP_q_ss};
                    const double flx[3] = {F(RSF_evap_soil), F(RSF_transp), F(RSF_q_ss)};
                    for (int k = 0; k < 3; k++) {
                        const double *t = TT(procs[k]);
                        for (int j = 0; j < A; j++) tmp[j] = flx[k] * t[j];
                        w[k] = np_sum(tmp, A);
                    }
                    ST(RSS_dS_num_error) = fabs((sum_s - sum_sa_s_m1) - (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - w[0] - w[1] - w[2]) * 1.0);
                    ST(RSS_dC_num_error) = fabs((sum_s * O(RSO_C_s) - sum_sa_s_m1 * *bp_C) -
                                                (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                                 F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - w[0] * nan0(O(RSO_C_evap_soil)) -
                                                 w[1] * nan0(O(RSO_C_transp)) - w[2] * nan0(O(RSO_C_q_ss))) * 1.0);
                    { /* numerics.py:1051-1122 */
                        const double lhs1 = sum_s - sum_sa_s_m1;
                        const double rhs1 = (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - w[0] - w[1] - w[2]) * 1.0;
                        const double lhs2 = sum_s * O(RSO_C_s) - sum_sa_s_m1 * *bp_C;
                        const double rhs2 = (F(RSF_inf_mat_rz) * nan0(O(RSO_C_inf_mat_rz)) + F(RSF_inf_pf_rz) * nan0(O(RSO_C_inf_pf_rz)) +
                                             F(RSF_inf_pf_ss) * nan0(O(RSO_C_inf_pf_ss)) - w[0] * nan0(O(RSO_C_evap_soil)) -
                                             w[1] * nan0(O(RSO_C_transp)) - w[2] * nan0(O(RSO_C_q_ss))) * 1.0;
                        const int c1 = (m != 0) ? np_isclose(lhs1, rhs1, g_rtol, g_atol) : 1;
                        const int c2 = (m != 0) ? np_isclose(lhs2, rhs2, g_rtol, g_atol) : 1;
                        const double sat_rz = g_bounds[0] ? g_bounds[0][i] : INFINITY, sat_ss = g_bounds[1] ? g_bounds[1][i] : INFINITY;
                        const double pwp_rz = g_bounds[2] ? g_bounds[2][i] : 0.0, pwp_ss = g_bounds[3] ? g_bounds[3][i] : 0.0;
                        const double sum_rz = np_sum(s_rz, A), sum_ss = np_sum(s_ss, A);
                        const int b3 = (sum_s <= (sat_rz + sat_ss) - (pwp_rz + pwp_ss)) && (sum_s >= 0);
                        const int b4 = (sum_rz <= sat_rz - pwp_rz) && (sum_rz >= 0);
                        const int b5 = (sum_ss <= sat_ss - pwp_ss) && (sum_ss >= 0);
                        ST(RSS_sanity_ok) = (c1 && c2 && b3 && b4 && b5) ? 1.0 : 0.0;
                    }
                    *bp_C = O(RSO_C_s);
                }
            }
            ageing(A, s_rz, ms_rz);
            ageing(A, s_ss, ms_ss);
#undef F
#undef O
#undef SP
#undef TT
#undef ST
        }
        free(work);
        free(tmp);
        free(ttk);
        free(ages);
    }
}
