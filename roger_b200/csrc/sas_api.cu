// sas_api.cu -- kernels and C ABI of the offline oxygen-18 StorAge-selection transport step
// (include/roger_b200_sas.h).  sm_100a only, no CPU fallback.
//
// HBM layout: every age-ranked array (sa_rz, sa_ss, msa_rz, msa_ss, and the two scratch arrays of the sub-step
// solver) is AGE-MAJOR: A planes of C doubles, cell index fastest; age bin j of the state arrays lives in plane
// (base + j) mod A so that the daily ageing shift is a decrement of `base` (sas_device.cuh).  The C ABI speaks
// the reference's (cell, age) order; upload / download transpose through a bounded staging buffer.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "sas_device.cuh"

using namespace rsas;

#define ROGER_EINVAL 22
#define ROGER_ENODEV 19
#define ROGER_ECUDA 5
#define ROGER_ENOTSUP 95

extern "C" const char* roger_last_error(void);
namespace roger { int set_last_error(int code, const std::string& msg); }
static int fail(int code, const std::string& msg) { return roger::set_last_error(code, msg); }
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ROGER_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    } while (0)

__device__ const roger_pow_log_entry d_pow_log_tab[ROGER_POW_N] = {ROGER_POW_LOG_TABLE};
__device__ const roger_pow_exp_entry d_pow_exp_tab[ROGER_POW_N] = {ROGER_POW_EXP_TABLE};

#ifndef ROGER_SAS_BLOCK
#define ROGER_SAS_BLOCK 256
#endif
#ifndef ROGER_SAS_MIN_BLOCKS
#define ROGER_SAS_MIN_BLOCKS 2
#endif

// One daily transport step of every cell (transport.py:948-991), one thread per cell.  STATS additionally evaluates the
// age statistics (transport.py:59-312) and the water / solute balance errors (numerics.py:424-445, :476-516).
template <bool STATS>
__global__ void __launch_bounds__(ROGER_SAS_BLOCK, ROGER_SAS_MIN_BLOCKS)
k_sas_day(SasParams P, Planes pl, const PwPlan* __restrict__ plan, const double* __restrict__ forcing,
          const double* __restrict__ mask, const double* __restrict__ sas_params, double* __restrict__ out,
          double* __restrict__ stats, double* __restrict__ C_s_prev, const double* __restrict__ bounds, double atol, double rtol) {
    __shared__ roger_pow_log_entry s_log[ROGER_POW_N];
    __shared__ roger_pow_exp_entry s_exp[ROGER_POW_N];
    for (int k = threadIdx.x; k < ROGER_POW_N; k += blockDim.x) {
        s_log[k] = d_pow_log_tab[k];
        s_exp[k] = d_pow_exp_tab[k];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t C = P.C;
    if (i >= C) return;
    CellCtx cx{P, pl, plan, roger_pow_tables{s_log, s_exp}, i};
    cx.assume_spaces();
    const double m = mask[i];
#define F(k) forcing[(int64_t)(k) * C + i]
#define O(k) out[(int64_t)(k) * C + i]
#define ST(k) stats[(int64_t)(k) * C + i]
#define SP(p) (sas_params + (int64_t)(p) * 8 * C + i)
    const double C_in = F(RSF_C_in);
    double sum_sa_s_m1 = 0;
    if (STATS) {  // vs.sa_s[taum1] as this step's soil kernel recomputes it (soil.py:1044-1048): the aged StorAge as found
        double sm[1];
        stream_sums<1, 2>(cx.plan,
            [&](int j, double* L) { const int64_t o = cx.ring(j); L[0] = pl.sa[0][o]; L[1] = pl.sa[1][o]; },
            [&](int, const double* L, double* v) { v[0] = L[0] + L[1] * m; }, sm);
        sum_sa_s_m1 = sm[0];
    }
    double a, b;
    FluxStats fs_evap, fs_transp, fs_qss;
    inflow(cx, 0, F(RSF_inf_mat_rz), C_in, m, &a, &b);
    O(RSO_C_inf_mat_rz) = a; O(RSO_C_iso_inf_mat_rz) = b;
    const double C_inf_mat_rz = a;
    inflow(cx, 0, F(RSF_inf_pf_rz), C_in, m, &a, &b);
    O(RSO_C_inf_pf_rz) = a; O(RSO_C_iso_inf_pf_rz) = b;
    const double C_inf_pf_rz = a;
    withdrawal<STATS>(cx, 0, -1, F(RSF_evap_soil), SP(RSP_evap_soil), C, m, &a, &b, RSP_evap_soil, &fs_evap);
    O(RSO_C_evap_soil) = a; O(RSO_C_iso_evap_soil) = b;
    const double C_evap = a;
    withdrawal<STATS>(cx, 0, -1, F(RSF_transp), SP(RSP_transp), C, m, &a, &b, RSP_transp, &fs_transp);
    O(RSO_C_transp) = a; O(RSO_C_iso_transp) = b;
    const double C_transp = a;
    withdrawal<STATS>(cx, 0, 1, F(RSF_q_rz), SP(RSP_q_rz), C, m, &a, &b, RSP_q_rz, nullptr);
    O(RSO_C_q_rz) = a; O(RSO_C_iso_q_rz) = b;
    inflow(cx, 1, F(RSF_inf_pf_ss), C_in, m, &a, &b);
    O(RSO_C_inf_pf_ss) = a; O(RSO_C_iso_inf_pf_ss) = b;
    const double C_inf_pf_ss = a;
    withdrawal<STATS>(cx, 1, -1, F(RSF_q_ss), SP(RSP_q_ss), C, m, &a, &b, RSP_q_ss, &fs_qss);
    O(RSO_C_q_ss) = a; O(RSO_C_iso_q_ss) = b;
    const double C_q_ss = a;
    withdrawal<STATS>(cx, 1, 0, F(RSF_cpr_rz), SP(RSP_cpr_rz), C, m, &a, &b, RSP_cpr_rz, nullptr);
    O(RSO_C_cpr_rz) = a; O(RSO_C_iso_cpr_rz) = b;
    double sum_sa_s, C_s, sum_sa_rz, sum_sa_ss;
    storage_concentrations(cx, m, out, C, &sum_sa_s, &C_s, &sum_sa_rz, &sum_sa_ss);
    if (STATS) {
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
        ST(RSS_dS_num_error) = fabs((sum_sa_s - sum_sa_s_m1) -
                                    (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - fs_evap.wsum - fs_transp.wsum - fs_qss.wsum) * 1.0);
        ST(RSS_dC_num_error) = fabs((sum_sa_s * C_s - sum_sa_s_m1 * Cp) -
                                    (F(RSF_inf_mat_rz) * nan0(C_inf_mat_rz) + F(RSF_inf_pf_rz) * nan0(C_inf_pf_rz) +
                                     F(RSF_inf_pf_ss) * nan0(C_inf_pf_ss) - fs_evap.wsum * nan0(C_evap) - fs_transp.wsum * nan0(C_transp) -
                                     fs_qss.wsum * nan0(C_q_ss)) * 1.0);
        {   // numerics.sanity_check, oxygen-18 branch (numerics.py:1051-1122), per cell
            auto isclose = [&](double x, double y) { return ((fabs(x - y) <= atol + rtol * fabs(y)) && isfinite(y)) || (x == y); };
            const double lhs1 = sum_sa_s - sum_sa_s_m1;
            const double rhs1 = (F(RSF_inf_mat_rz) + F(RSF_inf_pf_rz) + F(RSF_inf_pf_ss) - fs_evap.wsum - fs_transp.wsum - fs_qss.wsum) * 1.0;
            const double lhs2 = sum_sa_s * C_s - sum_sa_s_m1 * Cp;
            const double rhs2 = (F(RSF_inf_mat_rz) * nan0(C_inf_mat_rz) + F(RSF_inf_pf_rz) * nan0(C_inf_pf_rz) +
                                 F(RSF_inf_pf_ss) * nan0(C_inf_pf_ss) - fs_evap.wsum * nan0(C_evap) - fs_transp.wsum * nan0(C_transp) -
                                 fs_qss.wsum * nan0(C_q_ss)) * 1.0;
            const bool c1 = (m != 0) ? isclose(lhs1, rhs1) : true;
            const bool c2 = (m != 0) ? isclose(lhs2, rhs2) : true;
            const double inf = __longlong_as_double(0x7ff0000000000000LL);
            const double sat_rz = bounds ? bounds[i] : inf, sat_ss = bounds ? bounds[C + i] : inf;
            const double pwp_rz = bounds ? bounds[2 * C + i] : 0.0, pwp_ss = bounds ? bounds[3 * C + i] : 0.0;
            const bool b3 = (sum_sa_s <= (sat_rz + sat_ss) - (pwp_rz + pwp_ss)) && (sum_sa_s >= 0);
            const bool b4 = (sum_sa_rz <= sat_rz - pwp_rz) && (sum_sa_rz >= 0);
            const bool b5 = (sum_sa_ss <= sat_ss - pwp_ss) && (sum_sa_ss >= 0);
            ST(RSS_sanity_ok) = (c1 && c2 && b3 && b4 && b5) ? 1.0 : 0.0;
        }
        C_s_prev[i] = C_s;
    }
    ageing(cx);
#undef F
#undef O
#undef ST
#undef SP
}

// One daily step of the conservative-anion transport (bromide / chloride), one thread per cell.
template <bool STATS>
__global__ void __launch_bounds__(ROGER_SAS_BLOCK, ROGER_SAS_MIN_BLOCKS)
k_sas_day_anion(SasParams P, Planes pl, const PwPlan* __restrict__ plan, const double* __restrict__ forcing,
                const double* __restrict__ mask, const double* __restrict__ sas_params, double* __restrict__ out, AnionParams ap,
                double* __restrict__ stats, double* __restrict__ C_s_prev, CropParams crop) {
    __shared__ roger_pow_log_entry s_log[ROGER_POW_N];
    __shared__ roger_pow_exp_entry s_exp[ROGER_POW_N];
    for (int k = threadIdx.x; k < ROGER_POW_N; k += blockDim.x) {
        s_log[k] = d_pow_log_tab[k];
        s_exp[k] = d_pow_exp_tab[k];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.C) return;
    CellCtx cx{P, pl, plan, roger_pow_tables{s_log, s_exp}, i};
    cx.assume_spaces();
    anion_day<STATS>(cx, forcing, mask[i], sas_params, ap, out, stats, C_s_prev, nullptr, crop);
}

// One day of the nitrate transport: the anion day with the soil nitrogen cycle between the transport and the storage kernels
template <bool STATS>
__global__ void __launch_bounds__(ROGER_SAS_BLOCK, ROGER_SAS_MIN_BLOCKS)
k_sas_day_nitrate(SasParams P, Planes pl, const PwPlan* __restrict__ plan, const double* __restrict__ forcing,
                  const double* __restrict__ mask, const double* __restrict__ sas_params, double* __restrict__ out, AnionParams ap,
                  NitrateParams np_, CropParams crop, double* __restrict__ stats, double* __restrict__ C_s_prev) {
    __shared__ roger_pow_log_entry s_log[ROGER_POW_N];
    __shared__ roger_pow_exp_entry s_exp[ROGER_POW_N];
    for (int k = threadIdx.x; k < ROGER_POW_N; k += blockDim.x) {
        s_log[k] = d_pow_log_tab[k];
        s_exp[k] = d_pow_exp_tab[k];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.C) return;
    CellCtx cx{P, pl, plan, roger_pow_tables{s_log, s_exp}, i};
    cx.assume_spaces();
    anion_day<STATS, true>(cx, forcing, mask[i], sas_params, ap, out, stats, C_s_prev, &np_, crop);
}

// One day of the explicit Euler solver (`substeps` Euler steps, ageing with the last one), one thread per cell.
#ifndef ROGER_EULER_BLOCK
#define ROGER_EULER_BLOCK 256
#endif
#ifndef ROGER_EULER_MIN_BLOCKS
#define ROGER_EULER_MIN_BLOCKS 1
#endif
template <bool STATS, bool RK4>
__global__ void __launch_bounds__(ROGER_EULER_BLOCK, ROGER_EULER_MIN_BLOCKS)
k_sas_day_euler(SasParams P, Planes pl, const PwPlan* __restrict__ plan, const double* __restrict__ forcing,
                const double* __restrict__ mask, const double* __restrict__ sas_params, double* __restrict__ out,
                double* __restrict__ stats, double* __restrict__ C_s_prev, const double* __restrict__ bounds, double atol, double rtol) {
    __shared__ roger_pow_log_entry s_log[ROGER_POW_N];
    __shared__ roger_pow_exp_entry s_exp[ROGER_POW_N];
    for (int k = threadIdx.x; k < ROGER_POW_N; k += blockDim.x) {
        s_log[k] = d_pow_log_tab[k];
        s_exp[k] = d_pow_exp_tab[k];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.C) return;
    CellCtx cx{P, pl, plan, roger_pow_tables{s_log, s_exp}, i};
    cx.assume_spaces();
    euler_day<STATS, RK4>(cx, forcing, mask[i], sas_params, out, stats, C_s_prev, bounds, atol, rtol);
}

template <bool STATS>
__global__ void __launch_bounds__(ROGER_EULER_BLOCK, ROGER_EULER_MIN_BLOCKS)
k_sas_day_euler_anion(SasParams P, Planes pl, const PwPlan* __restrict__ plan, const double* __restrict__ forcing,
                      const double* __restrict__ mask, const double* __restrict__ sas_params, double* __restrict__ out, AnionParams ap,
                      double* __restrict__ stats, double* __restrict__ C_s_prev) {
    __shared__ roger_pow_log_entry s_log[ROGER_POW_N];
    __shared__ roger_pow_exp_entry s_exp[ROGER_POW_N];
    for (int k = threadIdx.x; k < ROGER_POW_N; k += blockDim.x) {
        s_log[k] = d_pow_log_tab[k];
        s_exp[k] = d_pow_exp_tab[k];
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.C) return;
    CellCtx cx{P, pl, plan, roger_pow_tables{s_log, s_exp}, i};
    cx.assume_spaces();
    euler_day_anion<STATS>(cx, forcing, mask[i], sas_params, ap, out, stats, C_s_prev);
}

// (cells, A) row-major staging block <-> age-major planes, 32 x 32 tiles through shared memory
__global__ void k_stage_to_planes(const double* __restrict__ stage, double* __restrict__ planes, int64_t C, int A, int base,
                                  int64_t cell0, int ncell) {
    __shared__ double tile[32][33];
    const int jt = blockIdx.x * 32, ct = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = ct + r, j = jt + threadIdx.x;
        if (c < ncell && j < A) tile[r][threadIdx.x] = stage[(int64_t)c * A + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = jt + r, c = ct + threadIdx.x;
        if (c < ncell && j < A) {
            int p = base + j;
            if (p >= A) p -= A;
            planes[(int64_t)p * C + cell0 + c] = tile[threadIdx.x][r];
        }
    }
}
__global__ void k_planes_to_stage(double* __restrict__ stage, const double* __restrict__ planes, int64_t C, int A, int base,
                                  int64_t cell0, int ncell) {
    __shared__ double tile[32][33];
    const int jt = blockIdx.x * 32, ct = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = jt + r, c = ct + threadIdx.x;
        if (c < ncell && j < A) {
            int p = base + j;
            if (p >= A) p -= A;
            tile[r][threadIdx.x] = planes[(int64_t)p * C + cell0 + c];
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = ct + r, j = jt + threadIdx.x;
        if (c < ncell && j < A) stage[(int64_t)c * A + j] = tile[threadIdx.x][r];
    }
}

// benchmarks/SVATOXYGEN18_benchmark.py:172-218: sa[1:] = diff(linspace(0, S, ages)), sa[0] = 0; msa = c0, msa[0] = 0
__global__ void k_init_linear(Planes pl, SasParams P, const double* __restrict__ S_rz, const double* __restrict__ S_ss, double c0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.C) return;
    const int A = P.A;
    for (int L = 0; L < 2; L++) {
        const double S = (L == 0) ? S_rz[i] : S_ss[i];
        const double step = (S - 0.0) / (double)(A - 1);  // numpy.linspace: step = delta / div
        Cursor c(P, i);
        double prev = 0.0;
        for (int j = 0; j < A; j++, c.next()) {
            double y = (double)j * step + 0.0;
            if (j == A - 1) y = S;  // endpoint is set exactly
            pl.sa[L][c.off] = (j == 0) ? 0.0 : y - prev;
            pl.msa[L][c.off] = (j == 0) ? 0.0 : c0;
            prev = y;
        }
    }
}

__global__ void k_pow_array(const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out, int64_t n, int use_exp) {
    __shared__ roger_pow_log_entry s_log[ROGER_POW_N];
    __shared__ roger_pow_exp_entry s_exp[ROGER_POW_N];
    for (int k = threadIdx.x; k < ROGER_POW_N; k += blockDim.x) {
        s_log[k] = d_pow_log_tab[k];
        s_exp[k] = d_pow_exp_tab[k];
    }
    __syncthreads();
    const roger_pow_tables tabs{s_log, s_exp};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = use_exp ? roger_exp(x[i], tabs) : roger_pow(x[i], y[i], tabs);
}

__global__ void k_gammainc_array(const double* __restrict__ a, const double* __restrict__ x, const double* __restrict__ G, double* __restrict__ out,
                                 int64_t n) {
    __shared__ roger_pow_log_entry s_log[ROGER_POW_N];
    __shared__ roger_pow_exp_entry s_exp[ROGER_POW_N];
    for (int k = threadIdx.x; k < ROGER_POW_N; k += blockDim.x) {
        s_log[k] = d_pow_log_tab[k];
        s_exp[k] = d_pow_exp_tab[k];
    }
    __syncthreads();
    const roger_pow_tables tabs{s_log, s_exp};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = roger_gammainc(a[i], x[i], G[i], tabs);
}

// div_by() against the division instruction on pseudo-random operand pairs (splitmix64 bits -> doubles with
// exponents inside the guarded range, quotients on both sides of one)
__global__ void k_div_selftest(uint64_t seed, int64_t n, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(t + 1);
        uint64_t w[2];
        for (int k = 0; k < 2; k++) {
            z += 0x9E3779B97F4A7C15ULL;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
            w[k] = x ^ (x >> 31);
        }
        // mantissas random; exponents: b in 2^[-40, 40], a = b * 2^[-60, 4] (a / b mostly below one, as SA / S)
        const int eb = (int)((w[0] >> 52) % 81) - 40;
        const int ea = eb + (int)((w[1] >> 52) % 65) - 60;
        const double b = __longlong_as_double((long long)(((uint64_t)(1023 + eb) << 52) | (w[0] & 0xFFFFFFFFFFFFFULL)));
        double a = __longlong_as_double((long long)(((uint64_t)(1023 + ea) << 52) | (w[1] & 0xFFFFFFFFFFFFFULL)));
        if ((t & 7) == 0) a = -a;
        if ((t & 1023) == 5) a = b;  // quotient exactly one
        const double q = rsas::div_by(a, b, 1.0 / b);
        if (q != a / b) bad++;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// ================================================================================================
struct roger_sas {
    int64_t C = 0;
    int device = 0;
    roger_sas_settings st{};
    int base = 0;
    double* d_state[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // sa_rz, sa_ss, msa_rz, msa_ss (nitrate: Nmin_rz, Nmin_ss)
    double* d_nit_params = nullptr;   // RNP_COUNT planes
    double* d_nit_forcing = nullptr;  // RNF_COUNT planes of the day
    double* d_nit_out = nullptr;      // RNO_COUNT planes
    double nit_dx = 1, nit_dy = 1, nit_zmax = 0.7, nit_dt = 24;
    bool have_nit_params = false, have_nit_forcing = false;
    double* d_crop_params = nullptr;  // crop variant: SAS parameter rows of re_rg and re_rl, [2][8][C]
    double* d_crop_re = nullptr;      // today's re_rg, re_rl, [2][C]
    double* d_crop_out = nullptr;     // RCO_COUNT planes
    bool have_crop_re = false;
    CropParams crop() const {
        if (!d_crop_params) return CropParams{nullptr, nullptr, nullptr, nullptr};
        return CropParams{d_crop_params, d_crop_params + 8 * C, d_crop_re, d_crop_out};
    }
    double* d_san = nullptr;
    double* d_TTn = nullptr;
    double* d_om = nullptr;   // ROGER_SAS_SPLIT: SAS function values between the two walks of a sub-step pass; RK4: a stage array
    double* d_rev[2] = {nullptr, nullptr};  // SAS family 52: cumulative StorAge read backwards (allocated by roger_sas_set_parameters)
    double* d_rk[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // RK4: stage array + accumulated travel-time distributions
    double* d_tt = nullptr;
    double* d_forcing = nullptr;
    double* d_out = nullptr;
    double* d_stats = nullptr;     // RSS_COUNT planes (settings.age_statistics)
    double* d_C_s_prev = nullptr;  // vs.C_s[taum1]
    double* d_bounds = nullptr;    // S_sat_rz, S_sat_ss, S_pwp_rz, S_pwp_ss (sanity check), optional
    double* d_anion = nullptr;     // alpha_transp, alpha_q, lu_id, S_sat_rz (anion transport), optional
    double atol = 1e-2, rtol = 1e-2;
    double* d_mask = nullptr;
    double* d_params = nullptr;
    double* d_stage = nullptr;
    int64_t stage_cells = 0;
    PwPlan* d_plan = nullptr;
    bool have_params = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_kernel_ms = 0;
    int64_t last_launches = 0;

    SasParams params() const {
        SasParams P;
        P.A = st.ages; P.substeps = st.substeps; P.store_tt = st.store_tt; P.base = base;
        P.vsmow = st.VSMOW_conc18O; P.dmin = st.d18O_min; P.dmax = st.d18O_max; P.C = C;
        return P;
    }
    Planes planes() const {
        Planes pl;
        pl.sa[0] = d_state[0]; pl.sa[1] = d_state[1]; pl.msa[0] = d_state[2]; pl.msa[1] = d_state[3];
        pl.san = d_san; pl.TTn = d_TTn; pl.tt = d_tt;
        pl.nmin[0] = d_state[4]; pl.nmin[1] = d_state[5];
        pl.om = d_om;
        for (int k = 0; k < 6; k++) pl.rk[k] = d_rk[k];
        pl.rev[0] = d_rev[0]; pl.rev[1] = d_rev[1];
        return pl;
    }
};

static void plan_rec(int n, std::vector<int>& len, std::vector<int>& merges) {
    if (n <= 128) {  // numpy PW_BLOCKSIZE
        len.push_back(n);
        merges.push_back(0);
        return;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    plan_rec(n2, len, merges);
    plan_rec(n - n2, len, merges);
    merges.back() += 1;
}

static int transfer_state(roger_sas* h, double* planes, int base, double* host, bool upload) {
    const int A = h->st.ages;
    dim3 tb(32, 8);
    for (int64_t c0 = 0; c0 < h->C; c0 += h->stage_cells) {
        const int n = (int)std::min<int64_t>(h->stage_cells, h->C - c0);
        dim3 grid((A + 31) / 32, (n + 31) / 32);
        const size_t bytes = sizeof(double) * (size_t)n * (size_t)A;
        if (upload) {
            CU(cudaMemcpyAsync(h->d_stage, host + (size_t)c0 * A, bytes, cudaMemcpyHostToDevice, h->stream));
            k_stage_to_planes<<<grid, tb, 0, h->stream>>>(h->d_stage, planes, h->C, A, base, c0, n);
        } else {
            k_planes_to_stage<<<grid, tb, 0, h->stream>>>(h->d_stage, planes, h->C, A, base, c0, n);
            CU(cudaMemcpyAsync(host + (size_t)c0 * A, h->d_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

extern "C" {

int roger_sas_create(roger_sas** out, int64_t n_cells, const roger_sas_settings* settings, int device) {
    if (!out || !settings || n_cells <= 0) return fail(ROGER_EINVAL, "roger_sas_create: bad argument");
    if (settings->ages < 2 || settings->substeps < 1) return fail(ROGER_EINVAL, "roger_sas_create: ages >= 2 and substeps >= 1 required");
    if (settings->solute < 0 || settings->solute > 3)
        return fail(ROGER_EINVAL, "roger_sas_create: solute must be 0 (isotope), 1 (anion), 2 (nitrate) or 3 (virtual tracer)");
    if (settings->solute == 2 && settings->solver != 0)
        return fail(ROGER_ENOTSUP, "roger_sas_create: nitrate runs with the deterministic solver");
    if (settings->solver < 0 || settings->solver > 2)
        return fail(ROGER_EINVAL, "roger_sas_create: solver must be 0 (deterministic), 1 (Euler) or 2 (RK4)");
    if (settings->solver == 2 && settings->solute != 0)
        return fail(ROGER_ENOTSUP, "roger_sas_create: the RK4 solver runs the isotope branch only");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ROGER_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this backend has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ROGER_EINVAL, "roger_sas_create: bad device index");
    CU(cudaSetDevice(device));
    std::vector<int> len, merges;
    plan_rec(settings->ages, len, merges);
    if ((int)len.size() > kMaxLeaves) return fail(ROGER_EINVAL, "roger_sas_create: too many age bins for the summation plan");
    roger_sas* h = new roger_sas();
    h->C = n_cells;
    h->device = device;
    h->st = *settings;
    const size_t arr = sizeof(double) * (size_t)n_cells * (size_t)settings->ages;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int k = 0; k < (settings->solute == 2 ? 6 : 4); k++) {
        CU(cudaMalloc(&h->d_state[k], arr));
        CU(cudaMemsetAsync(h->d_state[k], 0, arr, h->stream));
    }
    if (settings->solute == 2) {
        CU(cudaMalloc(&h->d_nit_params, sizeof(double) * RNP_COUNT * (size_t)n_cells));
        CU(cudaMalloc(&h->d_nit_forcing, sizeof(double) * RNF_COUNT * (size_t)n_cells));
        CU(cudaMalloc(&h->d_nit_out, sizeof(double) * RNO_COUNT * (size_t)n_cells));
        CU(cudaMemsetAsync(h->d_nit_out, 0, sizeof(double) * RNO_COUNT * (size_t)n_cells, h->stream));
    }
    // the scratch arrays carry a few planes of padding: the sub-step pass prefetches ROGER_SAS_UNROLL bins ahead unguarded
    static_assert(2 * ROGER_SAS_UNROLL <= 8, "the scratch padding below covers a look-ahead of 2 * ROGER_SAS_UNROLL planes");
    const size_t pad = sizeof(double) * (size_t)n_cells * 8;
    CU(cudaMalloc(&h->d_san, arr + pad));
    CU(cudaMalloc(&h->d_TTn, arr + pad));
    CU(cudaMemsetAsync(h->d_san, 0, arr + pad, h->stream));
    CU(cudaMemsetAsync(h->d_TTn, 0, arr + pad, h->stream));
#if ROGER_SAS_SPLIT
    static_assert(2 * ROGER_SAS_SPLIT_UNROLL <= 8, "scratch padding vs. the look-ahead of the split pass");
    CU(cudaMalloc(&h->d_om, arr + pad));
    CU(cudaMemsetAsync(h->d_om, 0, arr + pad, h->stream));
#endif
    if (settings->solver == 2) {  // stage copies and the accumulated stage distributions of the Runge-Kutta solver
        if (!h->d_om) {
            CU(cudaMalloc(&h->d_om, arr + pad));
            CU(cudaMemsetAsync(h->d_om, 0, arr + pad, h->stream));
        }
        for (int k = 0; k < 6; k++) {
            CU(cudaMalloc(&h->d_rk[k], arr));
            CU(cudaMemsetAsync(h->d_rk[k], 0, arr, h->stream));
        }
    }
    if (settings->store_tt) {
        CU(cudaMalloc(&h->d_tt, arr * RSP_COUNT));
        CU(cudaMemsetAsync(h->d_tt, 0, arr * RSP_COUNT, h->stream));
    }
    CU(cudaMalloc(&h->d_forcing, sizeof(double) * RSF_COUNT * (size_t)n_cells));
    CU(cudaMalloc(&h->d_out, sizeof(double) * RSO_COUNT * (size_t)n_cells));
    CU(cudaMemsetAsync(h->d_out, 0, sizeof(double) * RSO_COUNT * (size_t)n_cells, h->stream));
    if (settings->age_statistics) {
        CU(cudaMalloc(&h->d_stats, sizeof(double) * RSS_COUNT * (size_t)n_cells));
        CU(cudaMemsetAsync(h->d_stats, 0, sizeof(double) * RSS_COUNT * (size_t)n_cells, h->stream));
        CU(cudaMalloc(&h->d_C_s_prev, sizeof(double) * (size_t)n_cells));
        CU(cudaMemsetAsync(h->d_C_s_prev, 0, sizeof(double) * (size_t)n_cells, h->stream));
    }
    CU(cudaMalloc(&h->d_mask, sizeof(double) * (size_t)n_cells));
    CU(cudaMalloc(&h->d_params, sizeof(double) * RSP_COUNT * 8 * (size_t)n_cells));
    const int64_t budget = (int64_t)128 << 20;  // staging buffer bound
    h->stage_cells = std::max<int64_t>(32, std::min<int64_t>(n_cells, budget / (8 * (int64_t)settings->ages)));
    CU(cudaMalloc(&h->d_stage, sizeof(double) * (size_t)h->stage_cells * (size_t)settings->ages));
    PwPlan plan{};
    plan.n_leaves = (int)len.size();
    for (size_t l = 0; l < len.size(); l++) { plan.len[l] = (short)len[l]; plan.merges[l] = (unsigned char)merges[l]; }
    CU(cudaMalloc(&h->d_plan, sizeof(PwPlan)));
    CU(cudaMemcpyAsync(h->d_plan, &plan, sizeof(PwPlan), cudaMemcpyHostToDevice, h->stream));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaStreamSynchronize(h->stream));
    *out = h;
    return 0;
}

void roger_sas_destroy(roger_sas* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto p : h->d_state) cudaFree(p);
    cudaFree(h->d_stats); cudaFree(h->d_C_s_prev); cudaFree(h->d_bounds); cudaFree(h->d_anion);
    cudaFree(h->d_nit_params); cudaFree(h->d_nit_forcing); cudaFree(h->d_nit_out);
    for (int k = 0; k < 6; k++) cudaFree(h->d_rk[k]);
    cudaFree(h->d_rev[0]); cudaFree(h->d_rev[1]);
    cudaFree(h->d_crop_params); cudaFree(h->d_crop_re); cudaFree(h->d_crop_out);
    cudaFree(h->d_san); cudaFree(h->d_TTn); cudaFree(h->d_om); cudaFree(h->d_tt); cudaFree(h->d_forcing); cudaFree(h->d_out);
    cudaFree(h->d_mask); cudaFree(h->d_params); cudaFree(h->d_stage); cudaFree(h->d_plan);
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
}

int roger_sas_upload_state(roger_sas* h, int which, const double* host) {
    if (!h || !host || which < 0 || which > 5 || !h->d_state[which]) return fail(ROGER_EINVAL, "roger_sas_upload_state: bad argument");
    CU(cudaSetDevice(h->device));
    return transfer_state(h, h->d_state[which], h->base, const_cast<double*>(host), true);
}

int roger_sas_download_state(roger_sas* h, int which, double* host) {
    if (!h || !host || which < 0 || which > 5 || !h->d_state[which]) return fail(ROGER_EINVAL, "roger_sas_download_state: bad argument");
    CU(cudaSetDevice(h->device));
    return transfer_state(h, h->d_state[which], h->base, host, false);
}

uint64_t roger_sas_state_ptr(roger_sas* h, int which) {
    if (!h || which < 0 || which > 5) return 0;
    return (uint64_t)(uintptr_t)h->d_state[which];
}

int roger_sas_age_base(roger_sas* h) { return h ? h->base : -1; }

int roger_sas_init_linear(roger_sas* h, const double* S_rz, const double* S_ss, double conc0) {
    if (!h || !S_rz || !S_ss) return fail(ROGER_EINVAL, "roger_sas_init_linear: bad argument");
    CU(cudaSetDevice(h->device));
    // the two storage vectors travel through the forcing buffer (2 planes of C doubles)
    const size_t plane = sizeof(double) * (size_t)h->C;
    CU(cudaMemcpyAsync(h->d_forcing, S_rz, plane, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_forcing + h->C, S_ss, plane, cudaMemcpyHostToDevice, h->stream));
    h->base = 0;
    k_init_linear<<<(int)((h->C + 127) / 128), 128, 0, h->stream>>>(h->planes(), h->params(), h->d_forcing, h->d_forcing + h->C, conc0);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// Validates `nproc` blocks of (C, 8) SAS parameter rows and packs them plane-major ([proc][8][C]) for the device.
static int pack_sas_rows(roger_sas* h, const double* rows, int nproc, std::vector<double>& t, const char* who) {
    const int64_t C = h->C;
    t.resize((size_t)nproc * 8 * (size_t)C);
    bool any52 = false;
    for (int p = 0; p < nproc; p++)
        for (int64_t i = 0; i < C; i++) {
            const double* row = rows + ((size_t)p * (size_t)C + (size_t)i) * 8;
            const double id = row[0];
            const bool ok = id == 0 || id == 1 || id == 2 || id == 3 || (id >= 31 && id <= 37 && id == std::floor(id)) || id == 4 ||
                            id == 51 || id == 52 || id == 6 || id == 61 || id == 62;
            if (!ok) {
                char buf[320];
                snprintf(buf, sizeof buf, "%s: SAS function id %g (process %d, cell %lld) is not one of the reference's "
                         "(roger/core/sas.py: 1, 2, 3, 31-37, 4, 51, 52, 6, 61, 62)", who, id, p, (long long)i);
                return fail(ROGER_ENOTSUP, buf);
            }
            for (int k = 0; k < 8; k++) t[((size_t)p * 8 + (size_t)k) * (size_t)C + (size_t)i] = row[k];
            if (id == 4) {  // gamma family (sas.py:137-160): Gamma(a) = exp(gammaln(a)) is a per-flux constant, computed here once
                if (!(row[1] > 0 && row[1] <= ROGER_GAMMA_A_MAX && row[2] >= 0 && row[2] <= ROGER_GAMMA_B_MAX)) {
                    char buf[320];
                    snprintf(buf, sizeof buf, "%s: gamma SAS function (process %d, cell %lld) with a = %g, b = %g: this backend "
                             "evaluates 0 < a <= %g, 0 <= b <= %g", who, p, (long long)i, row[1], row[2], ROGER_GAMMA_A_MAX, ROGER_GAMMA_B_MAX);
                    return fail(ROGER_ENOTSUP, buf);
                }
                t[((size_t)p * 8 + 3) * (size_t)C + (size_t)i] = std::exp(std::lgamma(row[1]));
            }
            if (id == 52) any52 = true;
        }
    if (any52 && !h->d_rev[0]) {  // the backwards-read cumulative StorAge of family 52 (two scratch arrays of cells x ages)
        const size_t arr = sizeof(double) * (size_t)C * (size_t)h->st.ages;
        for (int k = 0; k < 2; k++) {
            CU(cudaMalloc(&h->d_rev[k], arr));
            CU(cudaMemsetAsync(h->d_rev[k], 0, arr, h->stream));
        }
    }
    return 0;
}

int roger_sas_set_parameters(roger_sas* h, const double* mask_catch, const double* sas_params) {
    if (!h || !mask_catch || !sas_params) return fail(ROGER_EINVAL, "roger_sas_set_parameters: bad argument");
    CU(cudaSetDevice(h->device));
    const int64_t C = h->C;
    std::vector<double> t;
    if (int rc = pack_sas_rows(h, sas_params, RSP_COUNT, t, "roger_sas_set_parameters")) return rc;
    CU(cudaMemcpyAsync(h->d_params, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_mask, mask_catch, sizeof(double) * (size_t)C, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_params = true;
    return 0;
}

int roger_sas_set_crop_parameters(roger_sas* h, const double* sas_params_re_rg, const double* sas_params_re_rl) {
    if (!h || !sas_params_re_rg || !sas_params_re_rl) return fail(ROGER_EINVAL, "roger_sas_set_crop_parameters: bad argument");
    if (h->st.solute == 0 || h->st.solver != 0)
        return fail(ROGER_ENOTSUP, "roger_sas_set_crop_parameters: the crop variant runs anion / nitrate solutes with the deterministic solver");
    CU(cudaSetDevice(h->device));
    const int64_t C = h->C;
    std::vector<double> rows((size_t)2 * 8 * (size_t)C), t;
    std::memcpy(rows.data(), sas_params_re_rg, sizeof(double) * 8 * (size_t)C);
    std::memcpy(rows.data() + 8 * (size_t)C, sas_params_re_rl, sizeof(double) * 8 * (size_t)C);
    if (int rc = pack_sas_rows(h, rows.data(), 2, t, "roger_sas_set_crop_parameters")) return rc;
    if (!h->d_crop_params) {
        CU(cudaMalloc(&h->d_crop_params, sizeof(double) * 16 * (size_t)C));
        CU(cudaMalloc(&h->d_crop_re, sizeof(double) * 2 * (size_t)C));
        CU(cudaMalloc(&h->d_crop_out, sizeof(double) * RCO_COUNT * (size_t)C));
        CU(cudaMemsetAsync(h->d_crop_out, 0, sizeof(double) * RCO_COUNT * (size_t)C, h->stream));
    }
    CU(cudaMemcpyAsync(h->d_crop_params, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_sas_set_redistribution(roger_sas* h, const double* re_rg, const double* re_rl) {
    if (!h || !re_rg || !re_rl) return fail(ROGER_EINVAL, "roger_sas_set_redistribution: bad argument");
    if (!h->d_crop_params) return fail(ROGER_EINVAL, "roger_sas_set_redistribution: roger_sas_set_crop_parameters has not been called");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->d_crop_re, re_rg, sizeof(double) * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_crop_re + h->C, re_rl, sizeof(double) * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));  // the host buffers are the caller's
    h->have_crop_re = true;
    return 0;
}

int roger_sas_download_redistribution(roger_sas* h, double* host) {
    if (!h || !host) return fail(ROGER_EINVAL, "roger_sas_download_redistribution: bad argument");
    if (!h->d_crop_out) return fail(ROGER_EINVAL, "roger_sas_download_redistribution: roger_sas_set_crop_parameters has not been called");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(host, h->d_crop_out, sizeof(double) * RCO_COUNT * (size_t)h->C, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

static int step_common(roger_sas* h, const double* d_forcing) {
    if (!h->have_params) return fail(ROGER_EINVAL, "roger_sas_step: roger_sas_set_parameters has not been called");
    if (h->d_crop_params && !h->have_crop_re)
        return fail(ROGER_EINVAL, "roger_sas_step: the crop variant needs the day's roger_sas_set_redistribution");
    CU(cudaEventRecord(h->ev0, h->stream));
    const int threads = ROGER_SAS_BLOCK;
    const int blocks = (int)((h->C + threads - 1) / threads);
    if (h->st.solver != 0) {
        const int eb = (int)((h->C + ROGER_EULER_BLOCK - 1) / ROGER_EULER_BLOCK);
        // the virtual tracer (solute 3) runs the anion branch of the Euler scheme unchanged: no solute leaves with the evaporation there
        // (transport.py:2308-2331), so its balance error (numerics.py:517-549) subtracts w * C_evap_soil = w * 0 and keeps the anion bits
        if (h->st.solute == 1 || h->st.solute == 3) {
            AnionParams ap{nullptr, nullptr, nullptr, nullptr};
            if (h->d_anion) ap = AnionParams{h->d_anion, h->d_anion + h->C, h->d_anion + 2 * h->C, h->d_anion + 3 * h->C};
            if (h->st.age_statistics)
                k_sas_day_euler_anion<true><<<eb, ROGER_EULER_BLOCK, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask,
                                                                                     h->d_params, h->d_out, ap, h->d_stats, h->d_C_s_prev);
            else
                k_sas_day_euler_anion<false><<<eb, ROGER_EULER_BLOCK, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask,
                                                                                      h->d_params, h->d_out, ap, nullptr, nullptr);
        } else {
            auto launch = [&](auto kernel) {
                if (h->st.age_statistics)
                    kernel<<<eb, ROGER_EULER_BLOCK, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params, h->d_out,
                                                                    h->d_stats, h->d_C_s_prev, h->d_bounds, h->atol, h->rtol);
                else
                    kernel<<<eb, ROGER_EULER_BLOCK, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params, h->d_out,
                                                                    nullptr, nullptr, nullptr, 0.0, 0.0);
            };
            if (h->st.solver == 2) {
                if (h->st.age_statistics) launch(k_sas_day_euler<true, true>);
                else launch(k_sas_day_euler<false, true>);
            } else {
                if (h->st.age_statistics) launch(k_sas_day_euler<true, false>);
                else launch(k_sas_day_euler<false, false>);
            }
        }
    } else if (h->st.solute == 2) {
        if (!h->have_nit_params || !h->have_nit_forcing)
            return fail(ROGER_EINVAL, "roger_sas_step: nitrate needs roger_sas_set_nitrate_parameters and the day's roger_sas_set_nitrate_forcing");
        AnionParams ap{nullptr, nullptr, nullptr, nullptr};
        if (h->d_anion) ap = AnionParams{h->d_anion, h->d_anion + h->C, h->d_anion + 2 * h->C, h->d_anion + 3 * h->C};
        NitrateParams np_{h->d_nit_params, h->d_nit_forcing, h->d_nit_out, h->nit_dx, h->nit_dy, h->nit_zmax, h->nit_dt};
        if (h->st.age_statistics)
            k_sas_day_nitrate<true><<<blocks, threads, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params, h->d_out,
                                                                       ap, np_, h->crop(), h->d_stats, h->d_C_s_prev);
        else
            k_sas_day_nitrate<false><<<blocks, threads, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params, h->d_out,
                                                                        ap, np_, h->crop(), nullptr, nullptr);
        h->have_nit_forcing = false;
        h->have_crop_re = false;
    } else if (h->st.solute == 1 || h->st.solute == 3) {
        AnionParams ap{nullptr, nullptr, nullptr, nullptr};
        if (h->d_anion) ap = AnionParams{h->d_anion, h->d_anion + h->C, h->d_anion + 2 * h->C, h->d_anion + 3 * h->C};
        ap.tracer = (h->st.solute == 3);
        if (h->st.age_statistics)
            k_sas_day_anion<true><<<blocks, threads, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params,
                                                                     h->d_out, ap, h->d_stats, h->d_C_s_prev, h->crop());
        else
            k_sas_day_anion<false><<<blocks, threads, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params,
                                                                      h->d_out, ap, nullptr, nullptr, h->crop());
        h->have_crop_re = false;
    } else if (h->st.age_statistics)
        k_sas_day<true><<<blocks, threads, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params, h->d_out,
                                                           h->d_stats, h->d_C_s_prev, h->d_bounds, h->atol, h->rtol);
    else
        k_sas_day<false><<<blocks, threads, 0, h->stream>>>(h->params(), h->planes(), h->d_plan, d_forcing, h->d_mask, h->d_params, h->d_out,
                                                            nullptr, nullptr, nullptr, 0.0, 0.0);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->base = (h->base == 0) ? h->st.ages - 1 : h->base - 1;  // ageing: the plane ring rotates by one bin
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_kernel_ms = ms;
    h->last_launches = 1;
    return 0;
}

int roger_sas_step(roger_sas* h, const double* forcing) {
    if (!h || !forcing) return fail(ROGER_EINVAL, "roger_sas_step: bad argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->d_forcing, forcing, sizeof(double) * RSF_COUNT * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
    return step_common(h, h->d_forcing);
}

int roger_sas_step_device(roger_sas* h, uint64_t forcing) {
    if (!h || !forcing) return fail(ROGER_EINVAL, "roger_sas_step_device: bad argument");
    CU(cudaSetDevice(h->device));
    return step_common(h, (const double*)(uintptr_t)forcing);
}

int roger_sas_set_nitrate_parameters(roger_sas* h, const double* params, double dx, double dy, double zroot_to_zsoil_max, double dt_hours) {
    if (!h || !params) return fail(ROGER_EINVAL, "roger_sas_set_nitrate_parameters: bad argument");
    if (h->st.solute != 2) return fail(ROGER_EINVAL, "roger_sas_set_nitrate_parameters: handle was not created with settings.solute = 2");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->d_nit_params, params, sizeof(double) * RNP_COUNT * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->nit_dx = dx; h->nit_dy = dy; h->nit_zmax = zroot_to_zsoil_max; h->nit_dt = dt_hours;
    h->have_nit_params = true;
    return 0;
}

int roger_sas_set_nitrate_forcing(roger_sas* h, const double* planes) {
    if (!h || !planes) return fail(ROGER_EINVAL, "roger_sas_set_nitrate_forcing: bad argument");
    if (h->st.solute != 2) return fail(ROGER_EINVAL, "roger_sas_set_nitrate_forcing: handle was not created with settings.solute = 2");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->d_nit_forcing, planes, sizeof(double) * RNF_COUNT * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_nit_forcing = true;
    return 0;
}

int roger_sas_download_nitrate_outputs(roger_sas* h, double* host) {
    if (!h || !host) return fail(ROGER_EINVAL, "roger_sas_download_nitrate_outputs: bad argument");
    if (h->st.solute != 2) return fail(ROGER_EINVAL, "roger_sas_download_nitrate_outputs: handle was not created with settings.solute = 2");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(host, h->d_nit_out, sizeof(double) * RNO_COUNT * (size_t)h->C, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_sas_download_outputs(roger_sas* h, double* host) {
    if (!h || !host) return fail(ROGER_EINVAL, "roger_sas_download_outputs: bad argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(host, h->d_out, sizeof(double) * RSO_COUNT * (size_t)h->C, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_sas_download_statistics(roger_sas* h, double* host) {
    if (!h || !host) return fail(ROGER_EINVAL, "roger_sas_download_statistics: bad argument");
    if (!h->d_stats) return fail(ROGER_EINVAL, "roger_sas_download_statistics: handle was created without settings.age_statistics");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(host, h->d_stats, sizeof(double) * RSS_COUNT * (size_t)h->C, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_sas_set_balance_reference(roger_sas* h, const double* C_s_prev) {
    if (!h || !C_s_prev) return fail(ROGER_EINVAL, "roger_sas_set_balance_reference: bad argument");
    if (!h->d_C_s_prev) return fail(ROGER_EINVAL, "roger_sas_set_balance_reference: handle was created without settings.age_statistics");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->d_C_s_prev, C_s_prev, sizeof(double) * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_sas_set_sanity_bounds(roger_sas* h, const double* S_sat_rz, const double* S_sat_ss, const double* S_pwp_rz, const double* S_pwp_ss,
                                double atol, double rtol) {
    if (!h || !S_sat_rz || !S_sat_ss || !S_pwp_rz || !S_pwp_ss) return fail(ROGER_EINVAL, "roger_sas_set_sanity_bounds: bad argument");
    CU(cudaSetDevice(h->device));
    const size_t plane = sizeof(double) * (size_t)h->C;
    if (!h->d_bounds) CU(cudaMalloc(&h->d_bounds, 4 * plane));
    const double* src[4] = {S_sat_rz, S_sat_ss, S_pwp_rz, S_pwp_ss};
    for (int k = 0; k < 4; k++) CU(cudaMemcpyAsync(h->d_bounds + (size_t)k * h->C, src[k], plane, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->atol = atol;
    h->rtol = rtol;
    return 0;
}

int roger_sas_set_anion_parameters(roger_sas* h, const double* alpha_transp, const double* alpha_q, const double* lu_id, const double* S_sat_rz) {
    if (!h || !alpha_transp || !alpha_q || !lu_id || !S_sat_rz) return fail(ROGER_EINVAL, "roger_sas_set_anion_parameters: bad argument");
    CU(cudaSetDevice(h->device));
    const size_t plane = sizeof(double) * (size_t)h->C;
    if (!h->d_anion) CU(cudaMalloc(&h->d_anion, 4 * plane));
    const double* src[4] = {alpha_transp, alpha_q, lu_id, S_sat_rz};
    for (int k = 0; k < 4; k++) CU(cudaMemcpyAsync(h->d_anion + (size_t)k * h->C, src[k], plane, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_sas_download_tt(roger_sas* h, int process, double* host) {
    if (!h || !host || process < 0 || process >= RSP_COUNT) return fail(ROGER_EINVAL, "roger_sas_download_tt: bad argument");
    if (!h->d_tt) return fail(ROGER_EINVAL, "roger_sas_download_tt: handle was created without settings.store_tt");
    CU(cudaSetDevice(h->device));
    return transfer_state(h, h->d_tt + (size_t)process * (size_t)h->st.ages * (size_t)h->C, 0, host, false);
}

// out[i] = roger_pow(x[i], y[i]) (or roger_exp(x[i]) when y is NULL) evaluated on the device: lets the tests prove
// that the device and host builds of roger_pow.h return identical bits.  Host pointers.
int roger_pow_device(const double* x, const double* y, double* out, int64_t n, int device) {
    if (!x || !out || n <= 0) return fail(ROGER_EINVAL, "roger_pow_device: bad argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(ROGER_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
    CU(cudaSetDevice(device));
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    const size_t bytes = sizeof(double) * (size_t)n;
    CU(cudaMalloc(&dx, bytes));
    CU(cudaMalloc(&dout, bytes));
    CU(cudaMemcpy(dx, x, bytes, cudaMemcpyHostToDevice));
    if (y) {
        CU(cudaMalloc(&dy, bytes));
        CU(cudaMemcpy(dy, y, bytes, cudaMemcpyHostToDevice));
    }
    k_pow_array<<<148 * 8, 256>>>(dx, dy, dout, n, y ? 0 : 1);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return 0;
}

// out[i] = roger_gammainc(a[i], x[i]) evaluated by the device build of roger_gamma.h, Gamma(a) = exp(lgamma(a)) computed here on the host
// as roger_sas_set_parameters does.  Host pointers.
int roger_gammainc_device(const double* a, const double* x, double* out, int64_t n, int device) {
    if (!a || !x || !out || n <= 0) return fail(ROGER_EINVAL, "roger_gammainc_device: bad argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(ROGER_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
    CU(cudaSetDevice(device));
    std::vector<double> G((size_t)n);
    for (int64_t i = 0; i < n; i++) G[(size_t)i] = std::exp(std::lgamma(a[i]));
    double *da = nullptr, *dx = nullptr, *dG = nullptr, *dout = nullptr;
    const size_t bytes = sizeof(double) * (size_t)n;
    CU(cudaMalloc(&da, bytes));
    CU(cudaMalloc(&dx, bytes));
    CU(cudaMalloc(&dG, bytes));
    CU(cudaMalloc(&dout, bytes));
    CU(cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dx, x, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dG, G.data(), bytes, cudaMemcpyHostToDevice));
    k_gammainc_array<<<148 * 8, 256>>>(da, dx, dG, dout, n);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(dx); cudaFree(dG); cudaFree(dout);
    return 0;
}

// number of operand pairs (out of n) for which the hoisted-reciprocal division of the SAS kernel differs from the
// division instruction; the parity tests require 0
int roger_selftest_division(int64_t n, uint64_t seed, int device, int64_t* mismatches) {
    if (n <= 0 || !mismatches) return fail(ROGER_EINVAL, "roger_selftest_division: bad argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(ROGER_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
    CU(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    CU(cudaMalloc(&d, sizeof(unsigned long long)));
    CU(cudaMemset(d, 0, sizeof(unsigned long long)));
    k_div_selftest<<<148 * 8, 256>>>(seed, n, d);
    CU(cudaGetLastError());
    unsigned long long hbad = 0;
    CU(cudaMemcpy(&hbad, d, sizeof hbad, cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = (int64_t)hbad;
    return 0;
}

int roger_sas_synchronize(roger_sas* h) {
    if (!h) return fail(ROGER_EINVAL, "roger_sas_synchronize: null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_sas_last_timing(roger_sas* h, double* kernel_ms, int64_t* launches) {
    if (!h) return fail(ROGER_EINVAL, "roger_sas_last_timing: null handle");
    if (kernel_ms) *kernel_ms = h->last_kernel_ms;
    if (launches) *launches = h->last_launches;
    return 0;
}

}  // extern "C"
