// svat_device.cuh -- fused per-cell SVAT / oneD time step for sm_100a (device code).
//
// One thread owns one soil column for the whole step: every parameter and state value of the
// cell is loaded once into registers, the complete process chain of RogerSetup.step
// (roger/roger.py:389-447: interception -> evapotranspiration -> snow -> infiltration ->
// percolation / lateral flow -> capillary rise -> storages -> balance check -> after_timestep)
// runs on registers, and the state is stored once.  The reference executes the same chain as
// ~60 whole-array NumPy kernels with several hundred full-array copies per step.
//
// Template switches:
//   LAT    oneD model (settings.enable_lateral_flow, roger/core/subsurface_runoff.py:1455-1466)
//   MASK1  the cell's maskCatch is exactly 1.0: every `* maskCatch` of the reference is the
//          identity (x * 1.0 == x for all x incl. NaN/Inf/-0) and is elided at compile time.
//
// The arithmetic keeps the reference's operation order, its where()-chains (including the
// order in which masks are captured), its operator-precedence slips and its NaN behaviour
// (comparisons with NaN are false; no fmin/fmax shortcuts) -- see SURVEY.md section 7
// "Reference quirks".  Built with -fmad=false so that a*b+c rounds twice as NumPy does.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/roger_b200.h"
#define ROGER_POW_CLEAN_INPUTS 1  // see roger_pow_from_log: 1 % on the step kernel (4.78e9 vs 4.73e9 cell-timesteps/s)
#include "roger_pow.h"

namespace roger {

// Everything a step needs besides the cell's own planes (kernel parameter, lives in the constant bank).
struct StepPlan {
    double dt;            // vs.dt (h) chosen by the adaptive time stepping for this step
    double ta_fm, sf, retention, r_mp, r_mp2, r_mp3_6, l_sc, pi, transp_water_stress, dx;
    double atol, rtol;
    double sqrt6x2;       // 6**0.5 * 2 (infiltration.py:628)
    double u_prec, u_ta_a, u_pet, u_ta_b;  // uniform forcing: aggregates computed by the scheduler
    int class_a;          // prec/ta[tau] first block: 0 keep, 1 daily, 2 hourly, 3 ten-minute
    int class_b;          // pet/ta second block (cond6..cond11): 0 keep, 1 ten-minute, 2 hourly, 3 daily
    int itt_day;          // vs.itt_day at entry of adaptive_time_stepping
    int substeps;         // int32(round(dt / (1/5)))  (infiltration.py:513)
    int cond1, cond2s, cond3s, cond4, cond5;  // event gates (infiltration.py:2155-2167)
    int irrigation_month; // month in 4..9 (root_zone.py:155)
    int uniform;          // forcing shared by all cells
    int lbc;              // settings.enable_macropore_lower_boundary_condition
    int clamp;            // after_timestep clamp of tiny negative storages (SVAT setups)
    double kf_max;        // settings.kf_max (groundwater boundary: calc_q_re)
    unsigned long long flux_mask[2];  // FLUX kernels: which flux planes are stored (bit f - first flux id); all ones in parity mode,
                                      // only the accumulated ones when a diagnostics accumulator is what asked for the flux kernel
};

#define RD __device__ __forceinline__

// ---- arithmetic helpers ---------------------------------------------------------------------
// ROGER_FAST_MATH = 0 (default): every power of the reference (`**` / npx.power on float64) is roger_pow()
// (roger_pow.h), the fixed IEEE operation sequence the CPU oracle can be switched to as well, so that kernel and oracle
// agree bit for bit -- the step takes hard decisions on cancellation residues (e.g. z_sat = +-1e-18 switches the
// lateral-flow branch, subsurface_runoff.py:52-246), which turn a last-bit difference of pow into a visibly different
// trajectory of single cells.
// ROGER_FAST_MATH = 1: powers as exp(y*log(x)) with the logarithm shared between powers of the same base, cbrt() and
// y*y*y for the macropore front: a few ulp from libm's pow, ~1e-10 parity on most cells but not bit-reproducible.
#ifndef ROGER_FAST_MATH
#define ROGER_FAST_MATH 0
#endif

static __device__ const roger_pow_log_entry g_svat_pow_log_tab[ROGER_POW_N] = {ROGER_POW_LOG_TABLE};
static __device__ const roger_pow_exp_entry g_svat_pow_exp_tab[ROGER_POW_N] = {ROGER_POW_EXP_TABLE};
RD double xpow(double x, double y) { return roger_pow(x, y, roger_pow_tables{g_svat_pow_log_tab, g_svat_pow_exp_tab}); }

// IEEE-identical a / b.  nvcc's inline double division branches to a ~60-instruction slow path whenever an
// operand's exponent field is 0 (i.e. also for an exact zero); on dry steps most numerators of the
// reference's formulas are exactly 0, so those cases are answered directly.
RD double dv(double a, double b) {
    if (a == 0.0 || b == 0.0) {
        if (a != a || b != b) return a + b;
        const unsigned long long sign = (__double_as_longlong(a) ^ __double_as_longlong(b)) & 0x8000000000000000ULL;
        if (a == 0.0) {
            if (b == 0.0) return __longlong_as_double(0x7ff8000000000000ULL);
            return __longlong_as_double(sign);
        }
        return __longlong_as_double(sign | 0x7ff0000000000000ULL);
    }
    return a / b;
}

constexpr double kDblMax = 1.7976931348623157e308;
RD bool pos_finite(double x) { return (x > 0) && (x <= kDblMax); }
RD bool nonneg_finite(double x) { return (x >= 0) && (x <= kDblMax); }

RD double powr(double x, double y) {  // x ** y (numpy power on float64)
#if ROGER_FAST_MATH
    return exp(y * log(x));
#else
    return xpow(x, y);
#endif
}
RD double cube_root(double x) {  // x ** (1 / 3)
#if ROGER_FAST_MATH
    return cbrt(x);
#else
    return xpow(x, 1.0 / 3);
#endif
}
RD double cube(double x) {  // x ** 3
#if ROGER_FAST_MATH
    return x * x * x;
#else
    return xpow(x, 3.0);
#endif
}

struct Cell {
#define ROGER_DECL(kind, name) double name;
    ROGER_SVAT_FIELDS(ROGER_DECL)
#undef ROGER_DECL
    double m;  // maskCatch
    // [taum1] companions (the reference copies [tau] -> [taum1] after every step)
    double prec_m1, z_wf_m1, z_wf_t0_m1, z_wf_t1_m1, y_mp_m1, y_sc_m1, S_m1, S_rz_m1, S_ss_m1, S_s_m1;
};

template <bool MASK1>
RD double mk(const Cell& c, double x) { return MASK1 ? x : x * c.m; }
#define M(x) mk<MASK1>(c, (x))

// Plane access of one thread.  Loads are issued where a field is first needed and stores as soon as a
// field is final, so that only the values a process actually works on are alive in registers
// (load everything up front costs 250+ registers and two resident blocks per SM).
#ifndef ROGER_L2_PREFETCH
#define ROGER_L2_PREFETCH 0
#endif
constexpr int kFirstFlux = RF_pet_res;  // the enum is ordered P.., S.., F..
constexpr int kNumPSPlanes = RF_pet_res;
constexpr int kNumFPlanes = RF_COUNT - RF_pet_res;
// Tile-major layout: the grid is cut into tiles of kTile consecutive cells and a tile holds ALL its planes back to back,
// [tile][plane][kTile cells].  A thread keeps one pointer to its own cell inside its tile; plane `id` of that cell sits at
// the compile-time byte offset id * kTile * 8, so every access is `LDG/STG [base + immediate]` -- no address arithmetic
// at all (the plane-major layout of round 1 spent 14 % of the step's instructions on it) -- and a block's whole working
// set is ONE contiguous ~220 KB region instead of ~110 streams a plane (134 MB at 4096^2 cells) apart, which is what the
// bulk L2 prefetch below and the DRAM page / TLB locality rest on.
constexpr int kTileShift = 8;
constexpr int kTile = 1 << kTileShift;  // cells per tile == threads of a k_svat_step block
constexpr int kTileMask = kTile - 1;
constexpr size_t kTileBytesPS = (size_t)kNumPSPlanes * kTile * sizeof(double);
constexpr size_t kTileBytesF = (size_t)kNumFPlanes * kTile * sizeof(double);

// Physical order of the parameter / state planes inside a tile (the public field ids of include/roger_b200_fields.h are
// untouched).  Planes the state-only step READS come first, in the order the processes first need them, so that ONE bulk
// L2 prefetch of the head of the tile fetches exactly what the block is going to load -- and in that order; then the
// oneD-only parameters (read by the lateral-flow kernels only), then the planes a step reads only under rare conditions,
// and last the planes the step only writes or never touches (derived parameters that are recomputed in registers).
#define ROGER_SVAT_PS_ORDER(X) \
    X(maskCatch) X(prec) X(ta) X(pet) \
    X(theta_ac) X(theta_ufc) X(theta_pwp) X(z_root) X(z_soil) \
    X(throughfall_coeff_top) X(throughfall_coeff_ground) X(S_int_top_tot) X(S_int_ground_tot) X(lu_id) \
    X(swe_top) X(S_int_top) X(S_int_ground) X(S_snow) X(z0) X(prec_event_csum) X(swe_top_tot) X(swe_ground) \
    X(S_dep) X(de) X(rew) X(tew) X(basal_evap_coeff) X(S_fp_rz) X(swe) X(theta_rz) X(basal_transp_coeff) X(S_lp_rz) X(z_evap) \
    X(sealing) X(no_wf) X(z_wf) X(z_wf_t0) X(z_wf_t1) X(z_wf_fc) X(inf_mat_event_csum) X(inf_mat_pot_event_csum) \
    X(inf_mp_event_csum) X(y_mp) X(inf_sc_event_csum) X(y_sc) X(theta_d) X(theta_d_rel) X(theta_d_t0) X(theta_d_rel_t0) \
    X(theta_d_fp) X(t_event_csum) X(z_sc) \
    X(ks) X(wfs) X(t_sat) X(lmpv) X(dmpv) X(mp_drain_area) X(inf_mat_rz) X(S_lp_ss) X(S_fp_ss) X(maskRiverLake) \
    X(z_sat) X(k_rz) X(h_rz) X(h_ss) X(n_salv) X(ha) X(z_gw) X(kf) X(ks_ss) \
    X(lambda_bc) X(m_bc) X(S) \
    /* ---- end of the planes every state-only SVAT step reads (kHotSvat) ---- */ \
    X(dmph) X(slope) X(v_mp_layer_1) X(v_mp_layer_2) X(v_mp_layer_3) X(v_mp_layer_4) \
    X(v_mp_layer_5) X(v_mp_layer_6) X(v_mp_layer_7) X(v_mp_layer_8) \
    /* ---- end of the planes every state-only oneD step reads (kHotOned) ---- */ \
    X(theta_4) X(theta_27) X(z_sc_max) X(k_stress_transp) X(k_ss) X(S_rz) X(S_ss) X(S_s) X(theta) \
    X(k_stress_evap) X(theta_ss) X(k) X(h) X(S_sur) \
    X(theta_sat) X(theta_fc) X(S_ac_rz) X(S_ufc_rz) X(S_pwp_rz) X(S_ac_ss) X(S_ufc_ss) X(S_pwp_ss) X(S_pwp_s) \
    /* groundwater boundary only (k_svat_step<.., GWB = true>) */ \
    X(n0) X(S_vad) X(S_vad_tot) X(z_gw_m1)
#define ROGER_SLOT_ENUM(name) SLOT_##name,
enum roger_svat_slot { ROGER_SVAT_PS_ORDER(ROGER_SLOT_ENUM) SLOT_COUNT };
#undef ROGER_SLOT_ENUM
static_assert(SLOT_COUNT == kNumPSPlanes, "ROGER_SVAT_PS_ORDER must list every P and S field exactly once");
constexpr int kHotSvat = SLOT_S + 1;            // planes [0, kHotSvat) are read by every SVAT step
constexpr int kHotOned = SLOT_v_mp_layer_8 + 1;  // planes [0, kHotOned) by every oneD step
// field id -> slot; folds to a constant wherever the id is one (every access of the step kernels)
__host__ __device__ constexpr int ps_slot(int id) {
    switch (id) {
#define ROGER_SLOT_CASE(name) case RF_##name: return SLOT_##name;
        ROGER_SVAT_PS_ORDER(ROGER_SLOT_CASE)
#undef ROGER_SLOT_CASE
        default: return -1;
    }
}
constexpr bool ps_order_is_permutation() {
    for (int id = 0; id < kNumPSPlanes; id++) {
        if (ps_slot(id) < 0) return false;
        for (int other = 0; other < id; other++) if (ps_slot(other) == ps_slot(id)) return false;
    }
    return true;
}
static_assert(ps_order_is_permutation(), "ROGER_SVAT_PS_ORDER: a field is missing or listed twice");
struct Io {
    char* __restrict__ ps;  // this thread's cell in plane 0 of its tile (parameter + state planes)
    char* __restrict__ fl;  // likewise in the flux planes (FLUX kernels only)
    unsigned long long fm0, fm1;  // which flux planes are stored
    RD Io(double* ps_, double* fl_, int64_t i, unsigned long long fm0_ = ~0ULL, unsigned long long fm1_ = ~0ULL)
        : ps(reinterpret_cast<char*>(ps_) + (size_t)(i >> kTileShift) * kTileBytesPS + (size_t)(i & kTileMask) * sizeof(double)),
          fl(fl_ ? reinterpret_cast<char*>(fl_) + (size_t)(i >> kTileShift) * kTileBytesF + (size_t)(i & kTileMask) * sizeof(double) : nullptr),
          fm0(fm0_), fm1(fm1_) {}
    RD double* at(int id) const { return reinterpret_cast<double*>(ps + (size_t)ps_slot(id) * (kTile * sizeof(double))); }
    RD double ldp(int id) const { return __ldg(at(id)); }
    RD double lds(int id) const { return *at(id); }
    RD void sts(int id, double v) const { *at(id) = v; }
    template <bool FLUX>
    RD void stf(int id, double v) const {
        if (FLUX) {
            const int b = id - kFirstFlux;
            if (((b < 64 ? fm0 : fm1) >> (b & 63)) & 1ULL) *reinterpret_cast<double*>(fl + (size_t)b * (kTile * sizeof(double))) = v;
        }
    }
    RD void pf(int id) const { asm volatile("prefetch.global.L2 [%0];" ::"l"(at(id))); }
};
#define LDP(name) c.name = io.ldp(RF_##name)
#define LDS(name) c.name = io.lds(RF_##name)
#define STS(name) io.sts(RF_##name, c.name)
#define STF(name) io.template stf<FLUX>(RF_##name, c.name)

// ---- interception.py:6-343 ------------------------------------------------------------------
template <bool FLUX, bool MASK1>
RD void interception(Cell& c, const StepPlan& p, const Io& io) {
    const double ta_fm = p.ta_fm;
    const bool warm = c.ta > ta_fm;
    const double tf_top = 1. - c.throughfall_coeff_top;
    const double tf_gr = 1. - c.throughfall_coeff_ground;
    // rain, upper storage (:6-71)
    c.rain_top = M(warm ? c.prec : 0.0);
    {
        double wtmx = p.retention * c.swe_top;
        double tot = M((c.S_int_top_tot < wtmx) ? wtmx : c.S_int_top_tot);
        double free_ = M((c.S_int_top < tot) ? tot - c.S_int_top : 0.0);
        double want = c.prec * tf_top;
        bool m1 = (free_ >= want) && warm && (free_ > 0);
        bool m2 = (free_ < want) && warm && (free_ > 0);
        double v = 0.0 + M((c.prec * tf_top) * (m1 ? 1.0 : 0.0));
        c.int_rain_top = M(m2 ? free_ : v);
        c.S_int_top += M(c.int_rain_top);
    }
    // rain, lower storage (:74-151)
    {
        double rain = M((c.prec - c.int_rain_top) * (warm ? 1.0 : 0.0));
        double free_ = M(((c.S_int_ground < c.S_int_ground_tot) && (c.S_snow <= 0)) ? c.S_int_ground_tot - c.S_int_ground : 0.0);
        double want = rain * tf_gr;
        bool m1 = (free_ >= want) && warm && (free_ > 0);
        bool m2 = (free_ < want) && warm && (free_ > 0);
        double v = 0.0 + M((rain * tf_gr) * (m1 ? 1.0 : 0.0));
        v = M(m2 ? free_ : v);
        c.int_rain_ground = M((c.lu_id == 599) ? 0.0 : v);
        c.S_int_ground += M(c.int_rain_ground);
        c.rain_ground = M(c.rain_top - c.int_rain_top - c.int_rain_ground);
        double add = M((c.S_snow > 0) ? 0.0 : c.rain_ground);
        c.z0 += add;
        c.prec_event_csum += add;
    }
    // snow, upper storage (:154-245)
    const bool cold = c.ta <= ta_fm;
    c.snow_top = M(cold ? c.prec : 0.0);
    {
        const double ta = c.ta, lu = c.lu_id;
        double s = c.swe_top_tot;
        s = M(((ta > -1) && (lu == 10)) ? 9.0 : s);
        s = M(((ta > -1) && (lu == 11)) ? 15.0 : s);
        s = M(((ta > -1) && (lu == 12)) ? 25.0 : s);
        const bool mid = (ta >= -3) && (ta <= -1);
        s = M((mid && (lu == 10)) ? 2.5 + 0.5 * ta * 9 : s);
        s = M((mid && (lu == 11)) ? 2.5 + 0.5 * ta * 15 : s);
        s = M((mid && (lu == 12)) ? 2.5 + 0.5 * ta * 25 : s);
        s = M(((ta < -3) && (lu == 10)) ? 18.0 : s);
        s = M(((ta < -3) && (lu == 11)) ? 30.0 : s);
        s = M(((ta < -3) && (lu == 12)) ? 50.0 : s);
        c.swe_top_tot = s;
        double free_ = M((c.swe_top >= s) ? 0.0 : s - c.swe_top);
        double want = c.prec * tf_top;
        bool m1 = (free_ >= want) && cold && (free_ > 0);
        bool m2 = (free_ < want) && cold && (free_ > 0);
        double v = 0.0 + M((c.prec * tf_top) * (m1 ? 1.0 : 0.0));
        c.int_snow_top = M(m2 ? free_ : v);
        c.S_int_top += M(c.int_snow_top);
        c.swe_top += M(c.int_snow_top);
    }
    // snow, lower storage (:248-318)
    {
        double snow = M((c.prec - c.int_snow_top) * (cold ? 1.0 : 0.0));
        double free_ = M((c.S_int_ground >= c.S_int_ground_tot) ? 0.0 : c.S_int_ground_tot - c.S_int_ground);
        double want = snow * tf_gr;
        bool m1 = (free_ >= want) && cold && (free_ > 0);
        bool m2 = (free_ < want) && cold && (free_ > 0);
        double v = 0.0 + M((snow * tf_gr) * (m1 ? 1.0 : 0.0));
        v = M(m2 ? free_ : v);
        c.int_snow_ground = M((c.lu_id == 599) ? 0.0 : v);
        c.S_int_ground += M(c.int_snow_ground);
        c.swe_ground += M(c.int_snow_ground);
        c.snow_ground = M(c.snow_top - c.int_snow_top - c.int_snow_ground);
        c.prec_event_csum += M(c.snow_ground);
    }
    // :321-343
    c.int_top = M(c.int_rain_top + c.int_snow_top);
    c.int_ground = M(c.int_rain_ground + c.int_snow_ground);
    c.int_prec = M(c.int_rain_top + c.int_rain_ground + c.int_snow_top + c.int_snow_ground);
    STS(swe_top_tot);
    STF(rain_top); STF(int_rain_top); STF(rain_ground); STF(int_rain_ground); STF(snow_top); STF(int_snow_top);
    STF(snow_ground); STF(int_snow_ground); STF(int_top); STF(int_ground); STF(int_prec);
}

// ---- evapotranspiration.py:9-599 -----------------------------------------------------------
RD bool lu_is_crop(double lu) { return (lu >= 500) && (lu < 600) && (lu == floor(lu)); }
RD bool lu_is_tree(double lu) { return lu == 10 || lu == 11 || lu == 12 || lu == 15 || lu == 16 || lu == 17; }

template <bool FLUX, bool MASK1>
RD void evapotranspiration(Cell& c, const StepPlan& p, const Io& io) {
    c.k_stress_evap = 0.0;  // fully rewritten below (de <= rew | rew < de <= tew | de > tew) unless de is NaN
    const bool dry = c.prec <= 0;
    // upper interception storage (:9-66)
    {
        bool ok = (c.S_int_top <= c.S_int_top_tot) && (c.S_int_top_tot > 0) && (c.S_int_top > 0);
        bool m1 = ok && (c.pet_res <= c.S_int_top);
        bool m2 = ok && (c.pet_res > c.S_int_top);
        double e = 0.0 + M(c.pet_res * (m1 ? 1.0 : 0.0));
        c.pet_res = M(m1 ? 0.0 : c.pet_res);
        e += M(c.S_int_top * (m2 ? 1.0 : 0.0));
        c.pet_res += M(-c.S_int_top * (m2 ? 1.0 : 0.0));
        c.evap_int_top = e;
        c.S_int_top += M(-e);
    }
    // lower interception storage (:69-134)
    {
        bool ok = (c.S_int_ground <= c.S_int_ground_tot) && (c.S_int_ground_tot > 0) && (c.S_int_ground > 0);
        bool m1 = ok && (c.pet_res <= c.S_int_ground);
        bool m2 = ok && (c.pet_res > c.S_int_ground);
        double e = 0.0 + M(c.pet_res * (m1 ? 1.0 : 0.0));
        c.pet_res = M(m1 ? 0.0 : c.pet_res);
        e += M(c.S_int_ground * (m2 ? 1.0 : 0.0));
        c.pet_res += M(-c.S_int_ground * (m2 ? 1.0 : 0.0));
        c.evap_int_ground = e;
        c.S_int_ground += M(-e);
        c.evap_int = c.evap_int_ground + M(c.evap_int_top);  // :129, mask binds to the last term
    }
    // depression storage (:137-194)
    {
        bool ok = (c.S_dep > 0) && (c.pet_res > 0) && dry;
        bool m1 = ok && (c.S_dep <= c.pet_res);
        bool m2 = ok && (c.S_dep > c.pet_res);
        double e = 0.0 + M(c.S_dep * (m1 ? 1.0 : 0.0));
        c.pet_res += M(-c.S_dep * (m1 ? 1.0 : 0.0));
        e += M(c.pet_res * (m2 ? 1.0 : 0.0));
        c.pet_res = M(m2 ? 0.0 : c.pet_res);
        c.evap_dep = e;
        bool m3 = (c.S_dep > 0) && (e > 0);
        c.S_dep += M(-e * (m3 ? 1.0 : 0.0));
    }
    c.evap_sur = c.evap_int_top + c.evap_int_ground + M(c.evap_dep);  // :197-212
    // soil evaporation (:215-345)
    {
        bool m3 = c.de <= c.rew;
        bool m4 = (c.de > c.rew) && (c.de <= c.tew);
        bool m5 = c.de > c.tew;
        double k = M(m3 ? 1.0 : c.k_stress_evap);
        k = M(m4 ? (c.tew - c.de) / (c.tew - c.rew) : k);
        k = M(m5 ? 0.0 : k);
        c.k_stress_evap = k;
        c.evap_coeff = M(c.basal_evap_coeff * k);
        double pe = M(c.pet_res * c.evap_coeff);
        c.pevap_soil = pe;
        bool ok = (c.S_fp_rz > 0) && (pe > 0) && (c.swe <= 0) && dry;
        bool m1 = ok && (pe <= c.S_fp_rz);
        bool m2 = ok && (pe > c.S_fp_rz);
        double e = 0.0 + M(pe * (m1 ? 1.0 : 0.0));
        c.pet_res += M(-pe * (m1 ? 1.0 : 0.0));
        c.pet_res = M((c.pet_res < 0) ? 0.0 : c.pet_res);
        e += M(c.S_fp_rz * (m2 ? 1.0 : 0.0));
        c.pet_res += M(-c.S_fp_rz * (m2 ? 1.0 : 0.0));
        c.pet_res = M((c.pet_res < 0) ? 0.0 : c.pet_res);
        c.evap_soil = M(e);
        c.S_fp_rz += M(-c.evap_soil);
    }
    // transpiration (:348-543)
    {
        double tws = p.transp_water_stress * c.theta_ufc + M(c.theta_pwp);  // :360
        double ks;
        if (lu_is_crop(c.lu_id)) { LDS(k_stress_transp); ks = c.k_stress_transp; }
        else ks = ((c.theta_rz - c.theta_pwp) / (tws - c.theta_pwp));
        ks = M(ks);
        ks = (ks > 1) ? 1.0 : ks;
        c.k_stress_transp = ks;
        double tc = M(c.basal_transp_coeff * ks);
        bool anoxia = (c.lu_id > 500) && (c.lu_id < 599) && (c.theta_rz >= 0.8 * c.theta_sat);
        if (anoxia) {
            double r = c.S_lp_rz / c.S_ac_rz;
            tc = ((r >= 0) && (r <= 1)) ? 1 - powr(r, 1.5) : 1.0;
        }
        tc = M(tc);
        c.transp_coeff = tc;
        double pt_ = M((c.pevap_soil < c.pet) ? c.pet - c.pevap_soil : 0.0);
        double ptr_ = M((c.evap_soil < c.pet) ? c.pet - c.evap_soil : 0.0);
        c.pt = M(pt_ * c.basal_transp_coeff);
        double ptransp = M(ptr_ * tc);
        ptransp = M(lu_is_tree(c.lu_id) ? c.pet * tc : ptransp);
        c.ptransp = ptransp;
        double res = M(ptransp);
        double t_lp = 0.0, t_fp = 0.0;
        const bool act = (ptransp > 0) && dry;
        bool m1 = (c.S_lp_rz > 0) && (res <= c.S_lp_rz) && act;
        t_lp += M(m1 ? res : 0.0);
        res = M(m1 ? 0.0 : res);
        bool m2 = (c.S_lp_rz > 0) && (res > c.S_lp_rz) && act;
        t_lp += M(m2 ? c.S_lp_rz : 0.0);
        res += M(m2 ? -c.S_lp_rz : 0.0);
        bool m3 = (c.S_fp_rz > 0) && (res <= c.S_fp_rz) && (c.S_lp_rz <= 0) && act;
        t_fp += M(m3 ? res : 0.0);
        res = M(m3 ? 0.0 : res);
        bool m4 = (c.S_fp_rz > 0) && (res > c.S_fp_rz) && (c.S_lp_rz <= 0) && act;
        t_fp += M(m4 ? c.S_fp_rz : 0.0);
        res += M(m4 ? -c.S_fp_rz : 0.0);
        res = M((res < 0) ? 0.0 : res);
        c.ptransp_res = res;
        c.S_lp_rz += M(-t_lp);
        c.S_fp_rz += M(-t_fp);
        c.transp = M(t_fp + t_lp);
    }
    // :546-599
    c.de += c.evap_soil + M(c.transp * (c.z_evap / c.z_root));  // :557
    c.aet_soil = M(c.evap_soil + c.transp);
    c.aet = M(c.evap_int_top + c.evap_int_ground + c.evap_dep + c.evap_soil + c.transp);
    STS(S_dep); STS(k_stress_evap); STS(k_stress_transp); STS(S_int_ground);
    STF(evap_int_top); STF(evap_int_ground); STF(evap_int); STF(evap_dep); STF(evap_sur); STF(evap_coeff);
    STF(pevap_soil); STF(evap_soil); STF(transp_coeff); STF(pt); STF(ptransp); STF(ptransp_res); STF(transp);
    STF(aet_soil); STF(aet);
}

// ---- snow.py:6-290 ---------------------------------------------------------------------------
template <bool FLUX, bool MASK1>
RD void snow(Cell& c, const StepPlan& p, const Io& io) {
    const double ta_fm = p.ta_fm;
    const bool cold = c.ta <= ta_fm;
    double add = M(c.snow_ground * (cold ? 1.0 : 0.0));  // :6-26
    c.S_snow += add;
    c.swe += add;
    c.S_snow += M(c.rain_ground * (((c.swe > 0) && (c.ta > ta_fm)) ? 1.0 : 0.0));  // :29-44
    const double pot = M(p.sf * (c.ta - ta_fm) * p.dt);
    // upper interception storage (:47-133)
    {
        bool has = (pot > 0) && (c.swe_top > 0);
        double melt = 0.0;
        melt = M((has && (pot <= c.swe_top)) ? pot : melt);
        melt = M((has && (pot > c.swe_top)) ? c.swe_top : melt);
        c.snow_melt_top = melt;
        bool m4 = (melt > 0) && (melt <= c.swe_top);
        bool m5 = (melt > 0) && (melt > c.swe_top);
        c.pet_res += M(-melt * (m4 ? 1.0 : 0.0));
        c.swe_top += M(-melt * (m4 ? 1.0 : 0.0));
        c.pet_res += M(-c.swe_top * (m5 ? 1.0 : 0.0));
        c.swe_top += M(m5 ? 0.0 : -c.swe_top);  // :98-101
        c.pet_res = M((c.pet_res < 0) ? 0.0 : c.pet_res);
        double wtmx = p.retention * c.swe_top;
        double q_ret = M((c.S_int_top > c.S_int_top_tot) ? c.S_int_top - c.swe_top : 0.0);
        bool over = c.S_int_top_tot < c.S_int_top;
        double drip = (q_ret > wtmx) ? q_ret - wtmx : (((wtmx <= 0) && over) ? c.S_int_top - c.S_int_top_tot : 0.0);
        drip = M(drip);
        c.snow_melt_drip = drip;
        c.S_snow += M(over ? drip : 0.0);
        c.S_int_top += M(over ? -drip : 0.0);
    }
    // lower interception storage (:136-193)
    {
        bool has = (pot > 0) && (c.swe_ground > 0);
        double melt = 0.0;
        melt = M((has && (pot <= c.swe_ground)) ? pot : melt);
        melt = M((has && (pot > c.swe_ground)) ? c.swe_ground : melt);
        c.snow_melt_ground = melt;
        bool m4 = (melt > 0) && (melt <= c.swe_ground);
        bool m5 = (melt > 0) && (melt > c.swe_ground);
        c.pet_res += M(-melt * (m4 ? 1.0 : 0.0));
        c.swe_ground += M(-melt * (m4 ? 1.0 : 0.0));
        c.pet_res += M(-c.swe_ground * (m5 ? 1.0 : 0.0));
        c.swe_ground += M(m5 ? 0.0 : -c.swe_ground);
    }
    // snow cover (:196-290)
    {
        bool has = (pot > 0) && (c.swe > 0);
        double melt = 0.0;
        melt = M((has && (pot <= c.swe)) ? pot : melt);
        melt = M((has && (pot > c.swe)) ? c.swe : melt);
        c.snow_melt = melt;
        bool m5 = (melt > 0) && (melt <= c.swe);
        bool m6 = (melt > 0) && (melt > c.swe);
        c.pet_res += M(-melt * (m5 ? 1.0 : 0.0));
        c.swe += M(-melt * (m5 ? 1.0 : 0.0));
        c.pet_res += M(-c.swe * (m6 ? 1.0 : 0.0));
        c.swe = M(m6 ? 0.0 : c.swe);
        c.pet_res = M((c.pet_res < 0) ? 0.0 : c.pet_res);
        double wtmx = p.retention * c.swe;
        double q_ret = M((c.S_snow > 0) ? c.S_snow - c.swe : 0.0);
        double q = (q_ret > wtmx) ? q_ret - wtmx : ((wtmx <= 0) ? c.S_snow : 0.0);
        q = M(q);
        c.q_snow = q;
        c.S_snow += M(-q);
        c.z0 += M(q);
        c.prec_event_csum += M(q);
    }
    STS(S_int_top); STS(swe_top); STS(swe_ground); STS(swe); STS(S_snow);
    STF(pet_res); STF(snow_melt_top); STF(snow_melt_drip); STF(snow_melt_ground); STF(snow_melt); STF(q_snow);
    // S_sur = (S_int_top + S_int_ground + S_dep + S_snow + z0) * maskCatch (surface.py:8-28): the first four are final
    c.S_sur = c.S_int_top + c.S_int_ground + c.S_dep + c.S_snow;
}

// ---- infiltration.py ---------------------------------------------------------------------------
template <bool MASK1>
RD double theta_deficit(const Cell& c) {  // :1563-1594
    double t = 0.0;
    t = M((c.z_root > 0) ? (c.theta_sat - c.theta_rz) * (1 - c.sealing / 1) : t);
    t = M((c.z_soil <= 0) ? 0.01 : t);
    t = M((t <= 0) ? 0.01 : t);
    return t;
}
template <bool MASK1>
RD double theta_deficit_rel(const Cell& c) {  // :1597-1632
    double t = 0.0;
    t = M((c.z_root > 0) ? ((c.theta_sat - c.theta_rz) / (c.theta_sat - c.theta_pwp)) * (1 - c.sealing / 1) : t);
    t = M((c.z_soil <= 0) ? 0.01 : t);
    t = M((t <= 0) ? 0.01 : t);
    return t;
}
template <bool MASK1>
RD double theta_deficit_fp(const Cell& c) {  // :1635-1666
    double t = 0.0;
    t = M((c.z_soil > 0) ? (c.theta_fc - c.theta_rz) * (1 - c.sealing / 1) : t);
    t = M((c.z_soil <= 0) ? 0.01 : t);
    t = M((t <= 0) ? 0.01 : t);
    return t;
}

template <bool FLUX, bool MASK1>
RD void event_bookkeeping_start(Cell& c, const StepPlan& p, const Io& io) {
    if (!p.cond1) { c.z_wf_m1 = c.z_wf; c.z_wf_t1_m1 = c.z_wf_t1; c.y_mp_m1 = c.y_mp; c.y_sc_m1 = c.y_sc; }
    if (p.cond1) {
        // calc_depth_shrinkage_cracks :1767-1826
        const double th = c.theta_rz;
        double z = (th < c.theta_4) ? c.z_sc_max
                                    : (((th >= c.theta_4) && (th < c.theta_27)) ? (th - c.theta_4) / (c.theta_27 - c.theta_4) : 0.0) * c.z_sc_max;
        z = M(z);
        z = M((th < c.theta_4) ? c.z_sc_max : z);
        z = M((th > c.theta_27) ? 0.0 : z);
        z = M((1 - c.sealing / 1) * z);
        z = M((z > c.z_root) ? c.z_root : z);
        z = M((c.lu_id == 13) ? 0.0 : z);
        c.z_sc = z;
        // set_event_vars :1829-1976
        c.no_wf = 1;
        c.z_wf = 0; c.z_wf_m1 = 0;
        c.z_wf_t0 = 0; c.z_wf_t0_m1 = 0;
        c.z_wf_t1 = 0; c.z_wf_t1_m1 = 0;
        c.z_wf_fc = 0;
        c.inf_mat_event_csum = 0;
        c.inf_mat_pot_event_csum = 0;
        c.inf_mp_event_csum = 0;
        c.y_mp = 0; c.y_mp_m1 = 0;
        c.inf_sc_event_csum = 0;
        c.y_sc = 0; c.y_sc_m1 = 0;
        double td = theta_deficit<MASK1>(c);
        double tdr = theta_deficit_rel<MASK1>(c);
        c.theta_d = M(td);
        c.theta_d_rel = M(tdr);
        c.theta_d_t0 = M(td);
        c.theta_d_rel_t0 = M(tdr);
        c.theta_d_fp = M(theta_deficit_fp<MASK1>(c));
        c.prec_event_csum = 0;
        c.t_event_csum = 0;
        c.de = 0;
    }
    // cond2 / cond3 are `scalar & grid.any()` gates whose kernels re-apply the per-cell mask
    // (infiltration.py:1987, :2005): a cell acts iff the scalar part holds and its own mask is true.
    if (p.cond2s) {  // set_event_vars_start_rainfall_pause :1979-1995 (+ calc_z_wf_fc :1535-1560)
        double z = M((c.theta_d_fp > 0) ? dv(c.inf_mat_event_csum, c.theta_d_fp) : c.z_wf);
        z = M((z > c.z_soil) ? c.z_soil : z);
        bool mk_ = (c.prec == 0) && (c.prec_m1 != 0);
        c.z_wf_fc = M(mk_ ? z : c.z_wf_fc);
    }
    if (p.cond3s) {  // set_event_vars_end_rainfall_pause :1998-2053
        bool mk_ = (c.prec != 0) && (c.prec_m1 == 0);
        c.no_wf = mk_ ? 2.0 : c.no_wf;
        double td = theta_deficit<MASK1>(c);
        c.theta_d = M(mk_ ? td : c.theta_d);
        double tdr = theta_deficit_rel<MASK1>(c);
        c.theta_d_rel = M(mk_ ? tdr : c.theta_d_rel);
        if (mk_) { c.z_wf_t1 = 0; c.z_wf_t1_m1 = 0; }
        c.prec_event_csum = mk_ ? 0.0 : c.prec_event_csum;
        c.t_event_csum = mk_ ? 0.0 : c.t_event_csum;
    }
    if (p.cond5) c.t_event_csum += p.dt;  // :2175-2180
    STS(theta_d_fp); STS(theta_d_rel_t0); STS(de); STS(z_wf_fc); STS(prec_event_csum); STS(t_event_csum);
}

template <bool FLUX, bool MASK1>
RD void green_ampt_and_matrix(Cell& c, const StepPlan& p, const Io& io) {
    c.inf_mat = 0.0;  // mask7 | mask8 rewrite it unless z0 / inf_mat_pot is NaN
    const double dt = p.dt;
    // calc_green_ampt_params :7-48 with calc_pi_gr/pi_m/sat_time/Fs :1669-1764
    const double ktw = c.ks * c.theta_d * c.wfs;
    c.pi_gr = M(c.ks * (((c.theta_d * c.wfs) / (c.prec_event_csum + 1)) + 1));
    const double pi_m = M(ktw);
    c.pi_m = M(pi_m);
    {
        bool le = c.pi_m <= c.prec_event_csum, z = (c.t_sat == 0);
        bool m1 = le && (c.pi_m > c.pi_gr) && z;
        bool m2 = ((c.prec * (1 / dt) - c.ks) * c.prec_event_csum > ktw) && le && (c.pi_m <= c.pi_gr) && z;
        double t = m1 ? c.t_event_csum - dt : c.t_sat;
        t = m2 ? c.t_event_csum + (ktw / (c.pi_m * (c.pi_m * -c.ks))) - (dt / c.pi_m) * c.prec_event_csum : t;
        c.t_sat = M(t);
    }
    {
        double Fs = M(ktw / (pi_m - c.ks));
        Fs = M((pi_m <= c.ks) ? pi_m : Fs);
        c.Fs = M(Fs);
    }
    // calc_inf_mat :51-427   (vs.Fs_t0 and vs.theta_d_t1 are never written by the reference: 0)
    const bool mask1 = (c.pi_m <= c.prec_event_csum) && (c.t_event_csum > c.t_sat) && (c.t_sat > 0);
    const bool mask2 = (c.pi_m > c.prec_event_csum) && (c.t_event_csum > c.t_sat) && (c.t_sat > 0);
    const bool mask3 = (c.t_sat > c.t_event_csum - dt) && (c.t_sat < c.t_event_csum);
    const bool mask4 = (c.pi_m > c.prec_event_csum) && (c.t_sat <= 0);
    const double ksdt = c.ks * dt;
    const double seal = ((1 - c.sealing) / 1);
    double pot = ksdt;
    double rec = 0.0;
    if (mask1 || mask3) {
        double a = M(c.ks * (c.t_event_csum - c.t_sat));
        double b = c.Fs + M(2 * c.theta_d * c.wfs);  // :77
        rec = (ksdt / 2) * (1 + (1 + 2 * b / a) / sqrt(1 + (4 * b / a) + (4 * (0.0 * 0.0) / (a * a))));
    }
    pot = M(mask1 ? rec * seal : pot);
    if (mask2) {
        double num = ksdt * c.wfs * c.theta_d;
        double l1 = M((c.z0 > ksdt) ? num / (c.z0 - ksdt) : num / ksdt);
        pot = ksdt * (1 + ((c.wfs * c.theta_d) / l1)) * seal;
    }
    pot = M(pot);
    {
        double rec3 = M(mask3 ? rec : 0.0);
        double sat3 = M(mask3 ? c.z0 * (c.t_sat - (c.t_event_csum - dt)) : 0.0);
        pot = M(mask3 ? sat3 + rec3 * seal : pot);
    }
    pot = M(mask4 ? c.pi_gr * seal : pot);
    c.inf_mat_pot = pot;
    double inf = c.inf_mat;  // previous value survives only if z0 / pot is NaN
    inf = M((c.z0 < pot) ? c.z0 : inf);
    inf = M((c.z0 >= pot) ? pot : inf);
    {
        double cap = (c.S_ac_rz + c.S_ufc_rz) - (c.S_lp_rz + c.S_fp_rz);
        inf = M((inf > cap) ? cap : inf);
        inf = M((inf < 0) ? 0.0 : inf);
    }
    c.inf_mat = inf;
    c.inf_mat_event_csum += M(inf);
    c.inf_mat_pot_event_csum += M(pot);
    // wetting fronts
    const bool wf1 = (c.no_wf == 1), wf2 = (c.no_wf == 2);
    double dz = 0.0;
    dz = M(wf1 ? dv(inf, c.theta_d_t0) : dz);
    dz = M(wf2 ? dv(inf, c.theta_d) : dz);
    dz = M(isfinite(dz) ? dz : 0.0);
    c.z_wf_t0 += dz;
    c.z_wf_t1 += dz;
    c.z_wf_t0 = M((c.z_wf_t0 > c.z_soil) ? c.z_soil : c.z_wf_t0);
    c.z_wf_t1 = M((c.z_wf_t1 > c.z_soil) ? c.z_soil : c.z_wf_t1);
    c.z0 += M(-inf);
    c.z0 = M((c.z0 < 0) ? 0.0 : c.z0);
    const bool pause = (c.z_wf_fc > 0) && (c.rain_ground <= 0);
    {
        double d0 = M((pause && wf1) ? dv(pot, c.theta_d_t0) : 0.0);
        c.z_wf_t0 += M(isfinite(d0) ? d0 : 0.0);
        c.z_wf_t0 = M(((c.z_wf_t0 > c.z_wf_fc) && (c.z_wf_fc > 0)) ? c.z_wf_fc : c.z_wf_t0);
        c.z_wf_t0 = M((c.z_wf_t0 > c.z_soil) ? c.z_soil : c.z_wf_t0);
        double d1 = M((pause && wf2) ? dv(pot, c.theta_d) : 0.0);
        c.z_wf_t1 += M(isfinite(d1) ? d1 : 0.0);
        c.z_wf_t1 = M(((c.z_wf_t1 > c.z_wf_fc) && (c.z_wf_fc > 0)) ? c.z_wf_fc : c.z_wf_t1);
        c.z_wf_t1 = M((c.z_wf_t1 > c.z_soil) ? c.z_soil : c.z_wf_t1);
    }
    const bool m14 = (c.z_wf_t0 >= c.z_wf_t1) && (c.z_wf_t1 <= 0);
    const bool m15 = (c.z_wf_t0 > c.z_wf_t1) && (c.z_wf_t1 > 0);
    const bool m20 = (c.z_wf_t0 <= c.z_wf_t1) && (c.z_wf_t1 > 0);
    c.z_wf = M(m14 ? c.z_wf_t0 : c.z_wf);
    c.theta_d = M(m14 ? c.theta_d_t0 : c.theta_d);
    c.theta_d_rel = M(m14 ? c.theta_d_rel_t0 : c.theta_d_rel);
    c.z_wf_m1 = M(m15 ? 0.0 : c.z_wf_m1);
    c.z_wf = M(m15 ? c.z_wf_t1 : c.z_wf);
    c.no_wf = m20 ? 1.0 : c.no_wf;
    c.z_wf = M(m20 ? c.z_wf_t0 : c.z_wf);
    c.theta_d = M(m20 ? c.theta_d_t0 : c.theta_d);
    c.theta_d_rel = M(m20 ? c.theta_d_rel_t0 : c.theta_d_rel);
    c.z_wf = M((c.z_wf > c.z_soil) ? c.z_soil : c.z_wf);
    c.theta_d = M(c.theta_d_t0);  // :411-416 (theta_d_t1 <= 0 always holds)
    STF(pi_gr); STF(pi_m); STF(Fs); STF(inf_mat_pot); STF(inf_mat);
    STS(inf_mat_event_csum); STS(inf_mat_pot_event_csum); STS(theta_d_rel); STS(no_wf);
    // reset_event_vars (cond4, :2056-2144) zeroes these after the infiltration kernels; nothing reads them in between
    if (p.cond4) {
        io.sts(RF_t_sat, 0.0); io.sts(RF_z_wf_t0, 0.0);
        if (FLUX) { c.pi_gr = 0; c.pi_m = 0; c.Fs = 0; STF(pi_gr); STF(pi_m); STF(Fs); }
    } else {
        STS(t_sat); STS(z_wf_t0);
    }
}

// Vertical macropores: radial wetting front y(t) advanced in `substeps` explicit sub-steps
// (infiltration.py:430-1077, loop body :582-777).
//
// Exact early exit.  One pass of the loop body is a pure function G of the carried triple
// (ym1, t, inf_mp_event_csum) plus loop-invariant inputs; the accumulator inf_mp only receives
// `+= inf_mp_di`.  If a pass returns the triple bit-identical to its input and added
// inf_mp_di == 0, every later pass repeats it verbatim, so the remaining passes are skipped
// without changing a single bit of the result.  This is what happens on dry steps (z0 == 0):
// the front stops after the second pass and the reference's other 118 of 120 daily sub-steps
// recompute the same numbers.  If inf_mp turns NaN it stays NaN (NaN + x) and so does
// inf_mp_event_csum and y; the reference then stores y_mp = inf_mp = 0 (:808-824), so the loop
// can stop at the first NaN as well.
template <bool FLUX, bool MASK1>
RD void macropores(Cell& c, const StepPlan& p, const Io& io) {
    const double dt = p.dt, r_mp = p.r_mp, pi = p.pi;
    const int substeps = p.substeps;
    // the reference's second `update(z_wf, ...)` discards the first (:443-462)
    const double z_wf = M((c.no_wf == 2) ? 0.0 : c.z_wf_t1);
    const double z_wf_m1 = M((c.no_wf == 2) ? 0.0 : c.z_wf_t1_m1);
    double nonsat = c.lmpv - M(z_wf);  // :468
    nonsat = M((nonsat < 0) ? 0.0 : nonsat);
    double dz = z_wf - M(z_wf_m1);  // :480
    dz = M((z_wf >= c.lmpv) ? nonsat : dz);
    dz = M((nonsat <= 0) ? 0.0 : dz);
    dz = M((dz <= 0) ? 0.0 : dz);
    nonsat = c.lmpv - M(c.z_wf);  // :503
    nonsat = M((nonsat < 0) ? 0.0 : nonsat);
    nonsat = M((substeps == 1) ? nonsat + dv(dz, 1.39) : nonsat);
    c.lmpv_non_sat = nonsat;

    const double a = M(c.theta_d * p.r_mp2);
    double ym1 = M(c.y_mp_m1);
    double y = ym1;
    double csum = M(c.inf_mp_event_csum);
    double t = 0.0, inf_mp = 0.0;
    const double z0_di = M(c.z0 * (c.mp_drain_area / substeps));
    const double dts = M(dt / substeps);
    const double ksw = c.ks * c.wfs;
    const double rtd2 = r_mp * (c.theta_d * c.theta_d);
    const double dmtd = c.dmpv * c.theta_d;
    const double tfac = c.theta_d / (ksw * r_mp);
    // Exact shortcut for a cell without ponded water (z0_di == 0), the state of every cell on a dry step.  With
    // theta_d, a, dmpv * theta_d positive and finite and ym1, nonsat, csum non-negative and finite, every quantity of a
    // pass is non-NaN up to pot_di >= 0 (or NaN only through inf * 0 with nonsat == 0, which the nonsat select turns into
    // 0), so inf_mp_di = min(pot_di, z0_di) = 0 in EVERY pass: csum never changes, inf_mp stays 0 and the front the loop
    // leaves behind is r_mp + sqrt(csum / (dmpv theta_d) / pi) whatever the number of passes -- provided that value is
    // finite (an infinite ym1 would turn the next pass's y*y - ym1*ym1 into NaN).  The general loop below needs two passes
    // with two powers each to find that out; anything outside these conditions takes it.
    bool quiet = false;
    if (MASK1 && (z0_di == 0.0) && pos_finite(c.theta_d) && pos_finite(a) && pos_finite(dmtd) && nonneg_finite(ym1) &&
        nonneg_finite(nonsat) && nonneg_finite(csum)) {
        double yq = r_mp + sqrt(dv(dv(csum, dmtd), pi));
        yq = (yq < r_mp) ? r_mp : yq;
        if (yq <= kDblMax) { y = yq; quiet = true; }
    }
    for (int i = 0; i < substeps && !quiet; i++) {
        const double ym1_in = ym1, t_in = t, csum_in = csum;
        t += dts;
        double cc = M(ksw * t);
        cc = M(isnan(cc) ? 0.0 : cc);
        double b1 = M(p.sqrt6x2 * sqrt(cc * (6 * cc - a)));
        b1 = M(isnan(b1) ? 0.0 : b1);
        double b2 = M(rtd2 * (12 * cc - a + b1));
        b2 = M(isnan(b2) ? 0.0 : b2);
        b2 = M((b2 <= 0) ? 0.0 : b2);
        const double cb = cube_root(b2);
        double y1 = M(dv(cb, c.theta_d) * 0.5);
        double y2 = M(dv(a, cb) * 0.5);
        y = M(y1 + y2 + ym1);
        y = M((y < r_mp) ? r_mp : y);
        y = M((y < ym1) ? ym1 : y);
        double pot_di = M(pi * (y * y - ym1 * ym1) * nonsat * c.theta_d * c.dmpv * 1e-06);
        double di = M((pot_di > z0_di) ? z0_di : pot_di);
        di = M((nonsat == 0) ? 0.0 : di);
        inf_mp += M(di);
        csum += M(di);
        y = r_mp + M(sqrt(dv(dv(csum, dmtd), pi)));  // :728-730
        y = M((y < r_mp) ? r_mp : y);
        t = M(tfac * (cube(y) / 3.0 - (y * y) * r_mp / 2.0 + p.r_mp3_6));
        inf_mp = M((inf_mp < 0) ? 0.0 : inf_mp);
        ym1 = M(y);
        if (isnan(inf_mp)) break;
        if (di == 0.0 && ym1 == ym1_in && t == t_in && csum == csum_in) break;
    }
    y = M(y);
    c.y_mp = M(isnan(y) ? 0.0 : y);
    inf_mp = M(inf_mp);
    inf_mp = M(isnan(inf_mp) ? 0.0 : inf_mp);

    double share = M((nonsat > 0) ? 1.0 - dv(c.lmpv - c.z_root, nonsat) : 0.0);
    share = M((c.lmpv <= c.z_root) ? 1.0 : share);
    share = M((z_wf >= c.z_root) ? 0.0 : share);
    share = M((share < 0) ? 0.0 : share);
    share = M((share > 1) ? 1.0 : share);
    double mp_rz = M(inf_mp * share);
    {   // :880-913; vs.inf_mat_rz still holds the PREVIOUS step's value here
        double free_rz = (c.S_ac_rz + c.S_ufc_rz) - (c.inf_mat_rz + c.S_lp_rz + c.S_fp_rz);
        mp_rz = M(((mp_rz > free_rz) && (free_rz >= 0)) ? free_rz : mp_rz);
        mp_rz = M((free_rz < 0) ? 0.0 : mp_rz);
    }
    c.inf_mp_rz = mp_rz;
    double mp_ss = M(inf_mp * (1 - share));
    {
        double free_ss = (c.S_ac_ss + c.S_ufc_ss) - (c.S_lp_ss + c.S_fp_ss);
        mp_ss = M(((mp_ss > free_ss) && (free_ss > 0)) ? free_ss : mp_ss);
    }
    c.inf_ss = M(mp_ss);
    c.S_fp_ss += M(c.inf_ss);
    bool ex = c.S_fp_ss > c.S_ufc_ss;
    c.S_lp_ss += M(ex ? (c.S_fp_ss - c.S_ufc_ss) : 0.0);
    c.S_fp_ss = M(ex ? c.S_ufc_ss : c.S_fp_ss);
    ex = c.S_lp_ss > c.S_ac_ss;
    mp_ss += M(ex ? -(c.S_lp_ss - c.S_ac_ss) : 0.0);
    mp_ss = M((mp_ss < 0) ? 0.0 : mp_ss);
    c.inf_mp_ss = mp_ss;
    c.S_lp_ss = M(ex ? c.S_ac_ss : c.S_lp_ss);
    c.inf_mp = c.inf_mp_rz + M(c.inf_mp_ss);  // :990
    c.inf_mp_event_csum += M(c.inf_mp);
    c.z0 += M(-c.inf_mp);
    c.z0 = M((c.z0 < 0) ? 0.0 : c.z0);
    if (p.lbc) {  // :1013-1065
        bool lbc = c.z_wf > c.lmpv;
        double lim = c.ks * dt * c.mp_drain_area;
        double add = lbc ? ((c.z0 > lim) ? lim : c.z0) : 0.0;
        c.inf_mp += M(add);
        double room = c.S_ac_ss - c.S_lp_ss;
        c.inf_mp = M(((c.z_wf > c.z_root) && (c.inf_mp > room) && (room > 0)) ? room : c.inf_mp);
        double sub = lbc ? ((c.z0 > lim) ? lim : c.z0) : 0.0;
        c.z0 += M(-sub);
        c.z0 = M((c.z0 < 0) ? 0.0 : c.z0);
    }
    STS(inf_mp_event_csum); STF(lmpv_non_sat); STF(inf_ss);
    if (p.cond4) io.sts(RF_y_mp, 0.0); else STS(y_mp);
}

// Shrinkage cracks (infiltration.py:1080-1318).  In the reference the loop's `inf_sc` carry is
// never updated (always 0) and `inf_sc_di` is doubled (:1248-1252); only y_sc and z_sc_non_sat
// evolve.  Same exact early exit as for the macropores: the pass is a pure function of the
// carried triple (ym1, t, inf_sc_event_csum).
template <bool FLUX, bool MASK1>
RD void shrinkage_cracks(Cell& c, const StepPlan& p, const Io& io) {
    const double dt = p.dt, l_sc = p.l_sc;
    const int substeps = p.substeps;
    const double z_wf = M((c.no_wf == 2) ? 0.0 : c.z_wf_t1);
    const double z_wf_m1 = M((c.no_wf == 2) ? 0.0 : c.z_wf_t1_m1);
    double nonsat = c.z_sc - M(z_wf);
    nonsat = M((nonsat < 0) ? 0.0 : nonsat);
    double dz = z_wf - M(z_wf_m1);
    dz = M((z_wf >= c.z_sc) ? nonsat : dz);
    dz = M((nonsat <= 0) ? 0.0 : dz);
    dz = M((dz <= 0) ? 0.0 : dz);
    nonsat = c.z_sc - M(c.z_wf);
    nonsat = M((nonsat < 0) ? 0.0 : nonsat);
    nonsat = M((substeps == 1) ? nonsat + dv(dz, 1.39) : nonsat);
    c.z_sc_non_sat = nonsat;

    double ym1 = M(c.y_sc_m1), y = ym1;
    double csum = M(c.inf_sc_event_csum);
    double t = 0.0;
    const double z0_di = M(dv(c.z0, (double)substeps));
    const double dts = M(dt / substeps);
    const double ksw = c.ks * c.wfs;
    const double geo = nonsat * c.theta_d * l_sc;
    // Exact shortcut without ponded water (z0_di == 0), as for the macropores: while every quantity of a pass stays finite
    // the pass computes a finite pot_di, clips it to >= 0 and takes min(pot_di, 0) = 0, so csum never changes and the front
    // left behind is csum / l_sc / 2 whatever the number of passes.  Finiteness is guaranteed by keeping the operands of
    // the pass inside [1e-30, 1e30] in magnitude (y = sqrt(2 ks wfs t / theta_d), t = ym1^2 theta_d / (2 ks wfs) with
    // ym1 = y_sc[taum1] or csum / l_sc / 2, and t <= 24 h: nothing leaves the double range); physical values are far inside.
    bool quiet = false;
    {
        auto mid = [](double v) { return (v >= 1e-30) && (v <= 1e30); };
        auto small = [](double v) { return fabs(v) <= 1e30; };
        if (MASK1 && (z0_di == 0.0) && mid(c.theta_d) && mid(ksw) && mid(fabs(l_sc)) && small(ym1) && small(nonsat) && small(csum) &&
            small(geo)) {
            y = dv(csum, l_sc) / 2;
            quiet = true;
        }
    }
    for (int i = 0; i < substeps && !quiet; i++) {
        const double ym1_in = ym1, t_in = t, csum_in = csum;
        t += dts;
        y = M(sqrt((ksw * t * 2) / c.theta_d));
        double pot_di = M(geo * (y - ym1) * 1e-06);
        pot_di = M((pot_di <= 0) ? 0.0 : pot_di);
        double di = M((pot_di > z0_di) ? z0_di : pot_di);
        di = M((nonsat <= 0) ? 0.0 : di);
        di += M(di);
        csum += M(di);
        y = M(dv(csum, l_sc) / 2);
        t = M(dv((ym1 * ym1) * c.theta_d, ksw * 2));
        ym1 = M(y);
        if (isnan(csum)) break;  // y = csum / l_sc / 2 stays NaN in every later pass
        if (ym1 == ym1_in && t == t_in && csum == csum_in) break;
    }
    c.y_sc = M(y);
    c.inf_sc = 0.0;
    c.inf_sc_event_csum += M(c.inf_sc);
    c.z0 += M(-c.inf_sc);
    c.z0 = M((c.z0 < 0) ? 0.0 : c.z0);
    STS(inf_sc_event_csum); STF(z_sc_non_sat); STF(inf_sc);
    if (p.cond4) {  // reset_event_vars :2056-2144
        io.sts(RF_y_sc, 0.0); io.sts(RF_z_sc, 0.0); io.sts(RF_z_wf_t1, 0.0);
        const double td = theta_deficit<MASK1>(c);
        io.sts(RF_theta_d, M(td)); io.sts(RF_theta_d_t0, M(td));
        c.z_wf = 0;
    } else {
        STS(y_sc); STS(z_sc); STS(z_wf_t1); STS(theta_d); STS(theta_d_t0);
    }
}

template <bool FLUX, bool MASK1>
RD void infiltration_tail(Cell& c, const StepPlan& p, const Io& io) {
    // calc_inf_rz :1321-1417
    c.inf_mat_rz = M(c.inf_mat);
    c.inf_sc_rz = M(c.inf_sc);
    c.inf_rz = M(c.inf_mat_rz + c.inf_mp_rz + c.inf_sc_rz);
    c.S_fp_rz += M(c.inf_rz);
    bool ex = c.S_fp_rz > c.S_ufc_rz;
    c.S_lp_rz += M(ex ? (c.S_fp_rz - c.S_ufc_rz) : 0.0);
    c.S_fp_rz = M(ex ? c.S_ufc_rz : c.S_fp_rz);
    ex = c.S_lp_rz > c.S_ac_rz;
    c.inf_mp_rz += M(ex ? -(c.S_lp_rz - c.S_ac_rz) : 0.0);
    c.inf_mp_rz = M((c.inf_mp_rz < 0) ? 0.0 : c.inf_mp_rz);
    c.z0 += M(ex ? c.S_lp_rz - c.S_ac_rz : 0.0);
    c.S_lp_rz = M(ex ? c.S_ac_rz : c.S_lp_rz);
    c.inf_mp = c.inf_mp_rz + M(c.inf_mp_ss);  // :1399
    c.inf_rz = M(c.inf_mat_rz + c.inf_mp_rz + c.inf_sc_rz);
    c.inf = M(c.inf_rz + c.inf_ss);  // :1519-1532
    // calc_hof_and_sof :1420-1476
    double hof = M(c.z0);
    hof = M((hof < 0) ? 0.0 : hof);
    c.q_hof = hof;
    bool sat = ((c.S_lp_rz + c.S_fp_rz) > (c.S_ac_rz + c.S_ufc_rz)) && ((c.S_lp_ss + c.S_fp_ss) >= (c.S_ac_ss + c.S_ufc_ss));
    c.q_sof = M(sat ? (c.S_lp_rz + c.S_fp_rz) - (c.S_ac_rz + c.S_ufc_rz) : 0.0);
    bool so = c.q_sof > 0;
    c.S_fp_rz = M(so ? c.S_ufc_rz : c.S_fp_rz);
    c.S_lp_rz = M(so ? c.S_ac_rz : c.S_lp_rz);
    // calc_surface_runoff :1479-1516
    c.z0 += M(-hof);
    c.z0 = M((c.z0 < 0) ? 0.0 : c.z0);
    c.q_sur = 0.0 + M(c.q_hof + c.q_sof);
    c.q_sur += (c.maskRiverLake != 0) ? c.prec : 0.0;
    STS(inf_mat_rz); STS(z0); STS(z_wf);
    STF(inf_mp); STF(inf_mp_rz); STF(inf_mp_ss); STF(inf_sc_rz); STF(inf_rz); STF(inf); STF(q_hof); STF(q_sof); STF(q_sur);
    c.S_sur = M(c.S_sur + c.z0);
    STS(S_sur);
}

// ---- subsurface_runoff.py / capillary_rise.py -------------------------------------------------
// Salvucci-type flux between two depths: ((z/(-ha*10.2))^-n - (h/ha)^-n) / (1 + (h/ha)^-n + (n-1)(z/(-ha*10.2))^-n)
RD double salvucci(double pa, double pb, double n_salv) { return (pa - pb) / (1 + pb + (n_salv - 1) * pa); }

template <bool LAT, bool FLUX, bool MASK1>
RD void subsurface(Cell& c, const StepPlan& p, const Io& io) {
    const double dt = p.dt;
    const double zsub = c.z_soil - c.z_root;
    // calc_rise_of_saturation_water_table :693-765
    {
        double lmpv_ss = c.lmpv - M(c.z_root);  // :705
        lmpv_ss = M((c.lmpv < c.z_root) ? 0.0 : lmpv_ss);
        const double lp_ss_ac = dv(c.S_lp_ss, c.theta_ac);
        double top = M((c.S_lp_ss < c.theta_ac) ? lp_ss_ac : c.S_lp_rz + lp_ss_ac);
        double nomp = zsub - lmpv_ss - M(c.z_sat);  // :733
        nomp = (nomp < 0) ? 0.0 : nomp;
        double v = lp_ss_ac;
        if (top > nomp) {
            if ((c.S_fp_ss >= c.S_ufc_ss) && (((c.S_lp_ss + 1e-6) / c.theta_ac) < zsub)) v = lp_ss_ac;
            else if ((c.S_fp_rz >= c.S_ufc_rz) && (c.S_lp_ss + 1e-6 >= c.S_ac_ss)) v = dv(c.S_lp_rz, c.theta_ac) + lp_ss_ac;
        }
        c.z_sat = M(v);
    }
    // calc_S_zsat :7-48
    c.S_zsat = M((c.z_sat <= c.z_soil) ? c.z_sat * c.theta_ac : c.z_soil * c.theta_ac);
    c.S_zsat_ss = M((c.z_sat <= zsub) ? c.S_zsat : zsub * c.theta_ac);
    c.S_zsat_rz = M((c.z_sat > zsub) ? (c.z_sat - zsub) * c.theta_ac : 0.0);
    if (LAT) {  // calc_z_sat_layer :52-245
#define ROGER_LAYER(k, name, clamp_top)                                  \
    {                                                                    \
        double z = (k == 0) ? M(c.z_sat) : c.z_sat - M(200.0 * k);       \
        if (clamp_top) z = M((z > 200) ? 200.0 : z);                     \
        z = M((z <= 0) ? 0.0 : z);                                       \
        c.name = z;                                                      \
    }
        ROGER_LAYER(0, z_sat_layer_1, true) ROGER_LAYER(1, z_sat_layer_2, true) ROGER_LAYER(2, z_sat_layer_3, true)
        ROGER_LAYER(3, z_sat_layer_4, true) ROGER_LAYER(4, z_sat_layer_5, true) ROGER_LAYER(5, z_sat_layer_6, true)
        ROGER_LAYER(6, z_sat_layer_7, true) ROGER_LAYER(7, z_sat_layer_8, false)
#undef ROGER_LAYER
    }
    const double mn = -c.n_salv;
    const double haz = -c.ha * 10.2;
    double pb_rz = 0.0;  // (h_rz/ha)^-n, shared by root-zone percolation and capillary rise (h_rz is unchanged in between)
    bool have_pb_rz = false;
    // calc_potential_percolation_rz :768-888
    {
        bool m1 = (c.z_wf < c.z_root) && (c.z_sat <= 0);
        bool m2 = (c.z_wf >= c.z_root) && (c.z_sat <= 0);
        bool m3 = (c.z_sat > 0) && (c.z_root < c.z_soil - c.z_sat);
        double perc = M(m1 ? c.k_rz * dt : 0.0);
        perc = M(m2 ? c.k_rz * dt : perc);
        if (m3) {
            double z = zsub - c.z_sat;
            double pa = powr(z / haz, mn);
            pb_rz = powr(-c.h_rz / -c.ha, mn);
            have_pb_rz = true;
            perc = salvucci(pa, pb_rz, c.n_salv) * dt * c.ks * (-1);
        }
        perc = M(perc);
        perc = M((perc < 0) ? 0.0 : perc);
        const double lim = c.z_soil - c.z_sat;
        const bool gate = c.z_root < lim;  // z_root[taum1] == z_root[tau]
        const double avail = c.S_lp_rz + c.S_fp_rz;
        bool m4 = (perc > 0) && (avail >= perc) && gate;
        bool m5 = (perc > 0) && (avail < perc) && gate;
        double q = 0.0;
        q = M(m4 ? perc : q);
        q = M(m5 ? c.S_fp_rz + c.S_lp_rz : q);
        double room = (c.S_ac_ss + c.S_ufc_ss) - (c.S_lp_ss + c.S_fp_ss);
        bool m6 = (q > 0) && (room > 0) && (q > room) && gate;
        q = M(m6 ? room : q);
        bool m7 = (c.S_lp_ss >= c.S_ac_ss - 1e-6) && (c.S_fp_ss >= c.S_ufc_ss - 1e-6);
        q = M(m7 ? 0.0 : q);
        q = M((c.z_root >= lim) ? 0.0 : q);
        c.q_pot_rz = q;
    }
    // calc_percolation_rz :891-968
    {
        bool below = c.z_sat < zsub;
        bool m1 = (c.S_lp_rz < c.q_pot_rz) && below;
        bool m2 = (c.S_lp_rz >= c.q_pot_rz) && below;
        double q = M(c.q_pot_rz);
        q = M((c.z_sat >= zsub) ? 0.0 : q);
        c.S_fp_rz += M(m1 ? -(q - c.S_lp_rz) : 0.0);
        c.S_lp_rz = M(m1 ? 0.0 : c.S_lp_rz);
        c.S_lp_rz += M(m2 ? -q : 0.0);
        c.S_fp_ss += M(q);
        bool ex = c.S_fp_ss > c.S_ufc_ss;
        c.S_lp_ss += M(ex ? c.S_fp_ss - c.S_ufc_ss : 0.0);
        c.S_fp_ss = M(ex ? c.S_ufc_ss : c.S_fp_ss);
        ex = c.S_lp_ss > c.S_ac_ss;
        q += M(ex ? -(c.S_lp_ss - c.S_ac_ss) : 0.0);
        c.S_lp_rz += M(ex ? c.S_lp_ss - c.S_ac_ss : 0.0);
        c.S_lp_ss = M(ex ? c.S_ac_ss : c.S_lp_ss);
        c.q_rz = q;
    }
    if (LAT) {
        // calc_potential_lateral_subsurface_runoff :249-362
        const double geom = (1 / (p.dx * (c.z_soil / 1000)));
        double mat = M((c.ks * c.slope * c.z_sat * p.dx * 1000 * dt) * 1e-6 * geom);
        mat = M((c.z_sat <= 0) ? 0.0 : mat);
        const double mpf = p.r_mp2;
#define ROGER_TERM(n) (c.z_sat_layer_##n * c.v_mp_layer_##n * dt * p.dx * 1000 * c.dmph * 1e-6 * mpf * p.pi * 1e-6)
        double sum = ROGER_TERM(1);
        sum = sum + ROGER_TERM(2);
        sum = sum + ROGER_TERM(3);
        sum = sum + ROGER_TERM(4);
        sum = sum + ROGER_TERM(5);
        sum = sum + ROGER_TERM(6);
        sum = sum + ROGER_TERM(7);
        sum = sum + ROGER_TERM(8);
#undef ROGER_TERM
        double mp = M(sum * geom);
        mp = M((mp < 0) ? 0.0 : mp);
        mp = M((c.z_sat <= 0) ? 0.0 : mp);
        double pot = M(mp + mat);
        double sh_mat = M(dv(mat, pot));
        sh_mat = M((pot == 0) ? 0.0 : sh_mat);
        double sh_mp = M(dv(mp, pot));
        sh_mp = M((pot == 0) ? 0.0 : sh_mp);
        c.q_sub_mat_share = sh_mat;
        c.q_sub_mp_share = sh_mp;
        pot = M((pot > c.S_lp_rz + c.S_lp_ss) ? c.S_lp_rz + c.S_lp_ss : pot);
        c.q_sub_pot = pot;
        c.q_sub_mat_pot = M(pot * sh_mat);
        c.q_sub_mp_pot = M(pot * sh_mp);
        // calc_lateral_subsurface_runoff_rz :365-457
        double rz_share = M((c.z_sat > 0) ? ((c.z_sat - zsub) / c.z_sat) : 0.0);
        bool off = (c.z_sat <= zsub) || (c.S_lp_rz <= 0);
        rz_share = M(off ? 0.0 : rz_share);
        rz_share = M(isnan(rz_share) ? 0.0 : rz_share);
        c.S_zsat_rz = M((c.z_sat * rz_share) * c.theta_ac);
        double want = c.q_sub_pot * rz_share;
        c.q_sub_rz = M((want < c.S_zsat_rz) ? want : c.S_zsat_rz);
        c.q_sub_mat_rz = M(c.q_sub_rz * sh_mat);
        c.q_sub_mp_rz = M(c.q_sub_rz * sh_mp);
        c.q_sub_mp_pot_rz = M(c.q_sub_mp_pot * rz_share);
        c.z_sat += M(dv(-c.q_sub_rz, c.theta_ac));
        c.S_lp_rz += M(-c.q_sub_rz);
        // calc_potential_lateral_subsurface_runoff_ss :461-517
        double ss_share = M(dv(zsub, c.z_sat));
        bool one = (c.z_sat <= zsub) || (c.S_lp_rz <= 0);
        bool zero = c.z_sat <= 0;
        bool nan_ = isnan(ss_share);
        ss_share = M(one ? 1.0 : ss_share);
        ss_share = M(zero ? 0.0 : ss_share);
        ss_share = M(nan_ ? 0.0 : ss_share);
        c.q_sub_mat_pot_ss = M(c.q_sub_mat_pot * ss_share);
        c.q_sub_mp_pot_ss = M(c.q_sub_mp_pot * ss_share);
        c.q_sub_pot_ss = M(c.q_sub_mat_pot_ss + c.q_sub_mp_pot_ss);
    }
    // calc_potential_percolation_ss :971-1099 (the z_gw <= 10 assignment overwrites the z_gw > 10 one)
    {
        const double zgw = c.z_gw * 1000;
        double z = (zgw - c.z_soil) + M(zsub / 2);
        double pa = powr(z / haz, mn);
        double pb = powr(-c.h_ss / -c.ha, mn);
        const double kfdt = c.kf * dt;
        double sal = salvucci(pa, pb, c.n_salv);
        double perc;
        if ((c.z_gw <= 10) && (zgw > c.z_soil) && (c.z_sat > 0)) { LDS(k_ss); perc = fmin(fmin(kfdt, c.ks_ss * dt), c.k_ss * dt); }
        else perc = fmin(kfdt, sal * dt * c.ks_ss * (-1));
        perc = M(perc);
        const bool above = c.z_soil < zgw;
        const double avail = c.S_fp_ss + c.S_lp_ss;
        bool m1 = (perc > 0) && above && (perc <= avail);
        bool m2 = (perc > 0) && above && (perc > avail);
        double q = 0.0;
        q = M(m1 ? perc : q);
        q = M(m2 ? avail : q);
        double cpr = M(sal * dt * c.ks_ss);
        cpr = M(((perc > 0) && above) ? 0.0 : cpr);
        cpr = M((zgw - c.z_soil > 10000) ? 0.0 : cpr);
        q = M((cpr > 0) ? 0.0 : q);
        c.q_pot_ss = q;
    }
    if (!LAT) {  // calc_percolation_ss :1102-1154
        c.q_ss = M(c.q_pot_ss);
        c.z_sat += M((c.z_sat > 0) ? dv(-c.q_ss, c.theta_ac) : 0.0);
        c.z_sat = M((c.z_sat < 0) ? 0.0 : c.z_sat);
        c.S_zsat_ss = M(c.z_sat * c.theta_ac);
        bool m1 = c.S_lp_ss < c.q_pot_ss;
        bool m2 = c.S_lp_ss >= c.q_pot_ss;
        c.S_fp_ss += M(m1 ? -(c.q_ss - c.S_lp_ss) : 0.0);
        c.S_lp_ss = M(m1 ? 0.0 : c.S_lp_ss);
        c.S_lp_ss += M(m2 ? -c.q_ss : 0.0);
    } else {  // calc_lateral_subsurface_runoff_ss :521-662, calc_lateral_subsurface_runoff :666-690
        double q = M((c.z_sat <= 0) ? c.q_pot_ss : 0.0);
        double tot = c.q_pot_ss + c.q_sub_pot_ss;
        double fv = M((tot > 0) ? dv(c.q_pot_ss, tot) : 0.0);
        double fl = M((tot > 0) ? dv(c.q_sub_pot_ss, tot) : 0.0);
        bool fits = tot <= c.S_zsat_ss;
        double q_sat = M(fits ? tot * fv : c.S_zsat_ss * fv);
        q = (c.z_sat > 0) ? q_sat : q;
        c.q_ss = q;
        c.q_sub_ss = M(fits ? tot * fl : c.S_zsat_ss * fl);
        c.q_sub_mat_ss = M(c.q_sub_ss * c.q_sub_mat_share);
        c.q_sub_mp_ss = M(c.q_sub_ss * c.q_sub_mp_share);
        bool m1 = c.S_lp_ss < q;
        bool m2 = c.S_lp_ss >= q;
        c.S_fp_ss += M(m1 ? -(q - c.S_lp_ss) : 0.0);
        c.S_lp_ss = M(m1 ? 0.0 : c.S_lp_ss);
        c.S_lp_ss += M(m2 ? -q : 0.0);
        c.S_lp_ss += M((c.z_sat > 0) ? -c.q_sub_ss : 0.0);
        c.z_sat += M(-dv(c.q_sub_ss + q, c.theta_ac));
        c.z_sat = M((c.z_sat < 0) ? 0.0 : c.z_sat);
        c.S_zsat = M(c.z_sat * c.theta_ac);
        c.q_sub_mat = M(c.q_sub_mat_rz + c.q_sub_mat_ss);
        c.q_sub_mp = M(c.q_sub_mp_rz + c.q_sub_mp_ss);
        c.q_sub = M(c.q_sub_rz + c.q_sub_ss);
    }
    // calc_cpr_rz capillary_rise.py:7-173
    {
        double z = M((c.z_root + zsub / 2) - c.z_root / 2);
        double pa = powr(z / haz, mn);
        double pb = have_pb_rz ? pb_rz : powr(-c.h_rz / -c.ha, mn);
        double cpr = M(salvucci(pa, pb, c.n_salv) * dt * c.ks);
        cpr = M((cpr < 0) ? 0.0 : cpr);
        cpr = M(isnan(cpr) ? 0.0 : cpr);
        cpr = M((c.S_lp_rz > 0) ? 0.0 : cpr);
        cpr = M((c.h_rz > c.h_ss) ? 0.0 : cpr);
        double avail = c.S_fp_ss + c.S_lp_ss;
        cpr = M((cpr > avail) ? avail : cpr);
        double room = c.S_ufc_rz - c.S_fp_rz;
        cpr = M(((cpr > room) && (room > 0)) ? room : cpr);
        c.cpr_rz = cpr;
        bool gate = (c.z_wf < c.z_root) || (c.z_sat < zsub);
        bool m1 = (cpr > 0) && (c.S_lp_ss <= 0) && gate;
        bool m2 = (cpr > 0) && (c.S_lp_ss > 0) && (cpr <= c.S_lp_ss) && gate;
        bool m3 = (cpr > 0) && (c.S_lp_ss > 0) && (cpr > c.S_lp_ss) && gate;
        c.S_fp_rz += M(m1 ? cpr : 0.0);
        c.S_fp_ss += M(m1 ? -cpr : 0.0);
        c.S_fp_rz += M(m2 ? cpr : 0.0);
        c.S_lp_ss += M(m2 ? -cpr : 0.0);
        c.S_fp_rz += M(m3 ? cpr : 0.0);
        c.S_fp_ss += M(m3 ? -(cpr - c.S_lp_ss) : 0.0);
        c.S_lp_ss = M(m3 ? 0.0 : c.S_lp_ss);
        bool ex = c.S_fp_rz > c.S_ufc_rz;
        c.S_lp_rz += M(ex ? (c.S_fp_rz - c.S_ufc_rz) : 0.0);
        c.S_fp_rz = M(ex ? c.S_ufc_rz : c.S_fp_rz);
    }
    STS(z_sat);
    STF(S_zsat); STF(S_zsat_rz); STF(S_zsat_ss); STF(q_pot_rz); STF(q_rz); STF(q_pot_ss); STF(q_ss); STF(cpr_rz);
    if (LAT) {
        STF(q_sub_mat_pot); STF(q_sub_mp_pot); STF(q_sub_pot); STF(q_sub_mat_share); STF(q_sub_mp_share); STF(q_sub_rz);
        STF(q_sub_mat_rz); STF(q_sub_mp_rz); STF(q_sub_mp_pot_rz); STF(q_sub_mat_pot_ss); STF(q_sub_mp_pot_ss);
        STF(q_sub_pot_ss); STF(q_sub_ss); STF(q_sub_mat_ss); STF(q_sub_mp_ss); STF(q_sub_mat); STF(q_sub_mp); STF(q_sub);
        STF(z_sat_layer_1); STF(z_sat_layer_2); STF(z_sat_layer_3); STF(z_sat_layer_4); STF(z_sat_layer_5);
        STF(z_sat_layer_6); STF(z_sat_layer_7); STF(z_sat_layer_8);
    }
}

// ---- groundwater boundary condition (settings.enable_groundwater_boundary) ----------------------------------------
// calc_cpr_ss capillary_rise.py:176-324 (no film flow): capillary rise from the water table into the subsoil plus the water a
// rising water table brings into it; runs right after calc_cpr_rz (capillary_rise.py:346-358)
template <bool FLUX, bool MASK1>
RD void capillary_rise_subsoil(Cell& c, const StepPlan& p, const Io& io) {
    LDS(z_gw_m1);
    const double dt = p.dt;
    const bool mask1 = (c.z_sat > 0) && (c.z_gw * 1000 < c.z_soil);
    const bool mask2 = c.z_gw * 1000 < c.z_soil;
    const double z = (c.z_gw * 1000 - c.z_soil) + M((c.z_soil - c.z_root) / 2);
    const double mn = -c.n_salv;
    const double pa = powr(z / (-c.ha * 10.2), mn);
    const double pb = powr(-c.h_ss / -c.ha, mn);
    double cpr = M((pa - pb) / (1 + pb + (c.n_salv - 1) * pa) * dt * c.ks_ss);
    cpr = M((cpr < 0) ? 0.0 : cpr);
    cpr = M(isnan(cpr) ? 0.0 : cpr);
    cpr = M((cpr < 0) ? 0.0 : cpr);
    const double S_ufc_ss = c.S_ufc_ss;
    cpr = M((cpr > (S_ufc_ss - c.S_fp_ss)) ? S_ufc_ss - c.S_fp_ss : cpr);
    cpr = M(mask1 ? 0.0 : cpr);
    cpr = M((cpr > c.S_lp_ss) ? cpr - c.S_lp_ss : cpr);
    cpr = M((c.z_gw * 1000 - c.z_soil > 10000) ? 0.0 : cpr);
    const double zgw_soil = M(((c.z_gw * 1000 < c.z_soil) && (c.z_gw_m1 * 1000 >= c.z_soil)) ? c.z_soil : c.z_gw_m1 * 1000);
    double gw_rise = M(-(c.z_gw * 1000 - zgw_soil) * c.theta_ac);
    gw_rise = M((gw_rise < 0) ? 0.0 : gw_rise);
    gw_rise = M(mask2 ? gw_rise : 0.0);
    cpr += M(gw_rise);
    c.cpr_ss = cpr;
    c.S_fp_ss += M(c.cpr_ss);
    const bool ex = c.S_fp_ss > S_ufc_ss;
    c.S_lp_ss += M(ex ? (c.S_fp_ss - S_ufc_ss) : 0.0);
    c.S_fp_ss = M(ex ? S_ufc_ss : c.S_fp_ss);
    STF(cpr_ss);
}

// calc_q_re groundwater_flow.py:99-160: recharge through the vadose zone below the soil, a linear reservoir that spills when full;
// runs after the storage kernels (roger.py:436-438).  Then vs.z_gw[taum1] = vs.z_gw[tau] (the setups' after_timestep).
template <bool FLUX, bool MASK1>
RD void groundwater_recharge(Cell& c, const StepPlan& p, const Io& io) {
    LDP(n0);
    LDS(S_vad);
    c.S_vad_tot = M((c.z_gw * 1000 > c.z_soil) ? (c.z_gw * 1000 - c.z_soil) * c.n0 : 0.0);
    c.S_vad += M(c.q_ss);
    const double k = M(c.kf / p.kf_max);
    c.q_re = M(k * c.S_vad);
    c.S_vad += M(-c.q_re);
    const double spill = M((c.S_vad > c.S_vad_tot) ? (c.S_vad - c.S_vad_tot) : 0.0);
    c.q_re += M(spill);
    c.S_vad += M(-spill);
    c.z_gw_m1 = c.z_gw;
    STS(S_vad); STS(S_vad_tot); STS(z_gw_m1);
    STF(q_re);
}

// Brooks-Corey conductivity and matric potential of one layer (root_zone.py:36-60, subsoil.py:22-48,
// soil.py:9-42): k = ks / (1 + r**-m), h = ha / r**(1/lambda) with r = theta / theta_sat.
template <bool MASK1>
RD void brooks_corey(const Cell& c, double theta, double mm, double inv_lambda, double& k, double& h) {
    const double r = theta / c.theta_sat;
#if ROGER_FAST_MATH
    const double lr = log(r);  // one logarithm for both powers
    k = M(c.ks / (1 + exp(mm * lr)));
    h = M(c.ha / exp(inv_lambda * lr));
#else
    double pk, ph;  // r**mm and r**inv_lambda, one logarithm (bit-identical to two separate powers)
    roger_pow2(r, mm, inv_lambda, roger_pow_tables{g_svat_pow_log_tab, g_svat_pow_exp_tab}, &pk, &ph);
    k = M(c.ks / (1 + pk));
    h = M(c.ha / ph);
#endif
}

// ---- storages (surface.py:8-28, root_zone.py:36-163, subsoil.py:6-130, soil.py:9-140, numerics.py:125-152)
template <bool FLUX, bool MASK1>
RD void storages(Cell& c, const StepPlan& p, const Io& io) {
    const double inv_lambda = 1 / c.lambda_bc;
    const double mm = -c.m_bc;
    c.S_rz = M(c.S_pwp_rz + c.S_fp_rz + c.S_lp_rz);
    c.dS_rz = M(c.S_rz - c.S_rz_m1);
    c.theta_rz = M((c.S_fp_rz + c.S_lp_rz) / c.z_root + c.theta_pwp);
    if (p.irrigation_month) {  // root_zone.py:7-33 with vs.theta_irr == 0 in the SVAT / oneD setups
        double d = 0.0 - c.theta_rz;
        d = (d <= 0) ? 0.0 : d;
        c.irr_demand = d * c.z_root;
    } else {
        c.irr_demand = 0.0;
    }
    brooks_corey<MASK1>(c, c.theta_rz, mm, inv_lambda, c.k_rz, c.h_rz);
    c.S_ss = M(c.S_pwp_ss + c.S_fp_ss + c.S_lp_ss);
    c.dS_ss = M(c.S_ss - c.S_ss_m1);
    c.theta_ss = M((c.S_fp_ss + c.S_lp_ss) / (c.z_soil - c.z_root) + c.theta_pwp);
    brooks_corey<MASK1>(c, c.theta_ss, mm, inv_lambda, c.k_ss, c.h_ss);
    c.S_fp_s = M(c.S_fp_rz + c.S_fp_ss);
    c.S_lp_s = M(c.S_lp_rz + c.S_lp_ss);
    c.S_s = M(c.S_pwp_s + c.S_fp_s + c.S_lp_s);
    c.dS_s = M(c.S_s - c.S_s_m1);
    c.theta = M((c.S_fp_s + c.S_lp_s) / c.z_soil + c.theta_pwp);
    // k, h of the whole soil column (soil.py:9-42) are results only -- no process reads them -- so the state-only kernels
    // leave them to soil_kh_refresh() below, which the host runs before anybody looks at those two planes
    if (FLUX) brooks_corey<MASK1>(c, c.theta, mm, inv_lambda, c.k, c.h);
    c.S = c.S_sur + M(c.S_s);       // numerics.py:134
    c.dS = c.S - M(c.S_m1);         // numerics.py:147
    STS(S_rz); STS(theta_rz); STS(k_rz); STS(h_rz); STS(S_ss); STS(theta_ss); STS(k_ss); STS(h_ss);
    STS(S_s); STS(theta); STS(S);
    if (FLUX) { STS(k); STS(h); }
    STF(S_fp_s); STF(S_lp_s); STF(dS_rz); STF(dS_ss); STF(dS_s); STF(dS); STF(irr_demand);
}

// k, h of the soil column from the stored theta: the statements storages<> skips in the state-only kernels, evaluated with
// the mask multiplications spelled out (x * 1.0 == x, so cells inside the catchment get the bits of the MASK1 path)
RD void soil_kh_refresh(const Io& io) {
    Cell c;
    c.m = io.ldp(RF_maskCatch);
    LDP(theta_ac); LDP(theta_ufc); LDP(theta_pwp); LDP(ks); LDP(ha); LDP(m_bc); LDP(lambda_bc);
    LDS(theta);
    constexpr bool MASK1 = false;
    c.theta_sat = M(c.theta_ac + c.theta_ufc + c.theta_pwp);
    const double inv_lambda = 1 / c.lambda_bc;
    const double mm = -c.m_bc;
    brooks_corey<MASK1>(c, c.theta, mm, inv_lambda, c.k, c.h);
    STS(k); STS(h);
}

// numerics.py:217-346 calc_dS_num_error and :742-800 / :978-1016 sanity_check; returns the cell's verdict
template <bool LAT, bool FLUX, bool GWB = false>
RD bool balance_check(Cell& c, const StepPlan& p, const Io& io) {
    const double lhs = c.S - c.S_m1;
    double rhs;
    if (GWB) {
        // the balance error (numerics.py:271-290 with lateral flow, :396-420 without) and the closure the sanity check tests
        // (:822-841, :902-918) are different expressions when lateral flow is on: q_re there, q_ss here
        if (LAT) {
            c.dS_num_error = fabs(lhs - (c.prec - c.q_sur - c.aet - c.q_re - c.q_sub + c.cpr_ss));
            rhs = c.prec - c.q_sur - c.aet - c.q_ss - c.q_sub + c.cpr_ss;
        } else {
            rhs = c.prec - c.q_sur - c.aet - c.q_ss + c.cpr_ss;
            c.dS_num_error = fabs(lhs - rhs);
        }
    } else if (LAT) {
        rhs = c.prec - c.q_sur - c.aet - c.q_ss - c.q_sub;
        c.dS_num_error = fabs(lhs - rhs);
    } else {
        rhs = c.prec - c.q_sur - c.aet - c.q_ss;
        c.dS_num_error = fabs(lhs - rhs);
        c.dS_rz_num_error = fabs((c.S_rz - c.S_rz_m1) -
                                 (c.inf_mat_rz + c.inf_mp_rz + c.inf_sc_rz + c.cpr_rz - c.transp - c.evap_soil - c.q_rz));
        c.dS_ss_num_error = fabs((c.S_ss - c.S_ss_m1) - (c.inf_mp_ss + c.q_rz - c.q_ss - c.cpr_rz));
    }
    bool ok = true;
    if (c.m != 0) {
        bool close = fabs(lhs - rhs) <= (p.atol + p.rtol * fabs(rhs));  // npx.isclose
        if (!(isfinite(lhs) && isfinite(rhs))) close = (lhs == rhs);
        ok = close;
    }
    auto nz = [](double x) { return isnan(x) ? 0.0 : x; };
    const double a = p.atol;
    ok = ok && (nz(c.S_fp_rz) > -a) && (nz(c.S_lp_rz) > -a) && (nz(c.S_fp_ss) > -a) && (nz(c.S_lp_ss) > -a);
    ok = ok && (nz(c.S_fp_rz) - a <= nz(c.S_ufc_rz)) && (nz(c.S_lp_rz) - a <= nz(c.S_ac_rz)) &&
         (nz(c.S_fp_ss) - a <= nz(c.S_ufc_ss)) && (nz(c.S_lp_ss) - a <= nz(c.S_ac_ss));
    STF(dS_num_error);
    if (!LAT && !GWB) { STF(dS_rz_num_error); STF(dS_ss_num_error); }
    return ok;
}

// Parameters the reference derives once at setup from the primary ones (soil.py:143-560); two flops each, so
// they are recomputed instead of read (80 B per cell-step less HBM traffic).  Same expressions, same rounding.
template <bool MASK1>
RD void derived_parameters(Cell& c, const Io& io) {
    LDP(theta_ac); LDP(theta_ufc); LDP(theta_pwp); LDP(z_root); LDP(z_soil);
    c.theta_sat = M(c.theta_ac + c.theta_ufc + c.theta_pwp);
    c.theta_fc = M(c.theta_ufc + c.theta_pwp);
    c.S_ac_rz = M(c.theta_ac * c.z_root);
    c.S_ufc_rz = M(c.theta_ufc * c.z_root);
    c.S_pwp_rz = M(c.theta_pwp * c.z_root);
    const double zsub = c.z_soil - c.z_root;
    c.S_ac_ss = M(c.theta_ac * zsub);
    c.S_ufc_ss = M(c.theta_ufc * zsub);
    c.S_pwp_ss = M(c.theta_pwp * zsub);
    c.S_pwp_s = M(c.z_soil * c.theta_pwp);
}

// ---- plane loads, grouped by the process that first needs them.  cell_step issues each group one process
// ahead of its use (software prefetch): with 16 resident warps per SM the ~600-cycle HBM latency of a group
// is then covered by the arithmetic of the process running in between.
RD void load_interception(Cell& c, const Io& io) {
    LDP(throughfall_coeff_top); LDP(throughfall_coeff_ground); LDP(S_int_top_tot); LDP(S_int_ground_tot); LDP(lu_id);
    LDS(swe_top); LDS(S_int_top); LDS(S_int_ground); LDS(S_snow); LDS(z0); LDS(prec_event_csum); LDS(swe_top_tot);
    LDS(swe_ground);
}
RD void load_evapotranspiration(Cell& c, const Io& io) {
    LDS(S_dep); LDS(de); LDP(rew); LDP(tew); LDP(basal_evap_coeff); LDS(S_fp_rz); LDS(swe); LDS(theta_rz);
    LDP(basal_transp_coeff); LDS(S_lp_rz); LDP(z_evap);
}
RD void load_event_state(Cell& c, const StepPlan& p, const Io& io) {
    LDP(sealing);
    if (!p.cond1) {  // (set_event_vars overwrites every one of these)
        LDS(no_wf); LDS(z_wf); LDS(z_wf_t0); LDS(z_wf_t1); LDS(z_wf_fc); LDS(inf_mat_event_csum);
        LDS(inf_mat_pot_event_csum); LDS(inf_mp_event_csum); LDS(y_mp); LDS(inf_sc_event_csum); LDS(y_sc);
        LDS(theta_d); LDS(theta_d_rel); LDS(theta_d_t0); LDS(theta_d_rel_t0); LDS(theta_d_fp); LDS(t_event_csum);
        LDS(z_sc);
    } else {
        LDP(theta_4); LDP(theta_27); LDP(z_sc_max);
    }
}
RD void load_infiltration(Cell& c, const Io& io) {
    LDP(ks); LDP(wfs); LDS(t_sat);
    LDP(lmpv); LDP(dmpv); LDP(mp_drain_area); LDS(inf_mat_rz); LDS(S_lp_ss); LDS(S_fp_ss);
    LDP(maskRiverLake);
}
template <bool LAT>
RD void load_subsurface(Cell& c, const Io& io) {
    LDS(z_sat); LDS(k_rz); LDS(h_rz); LDS(h_ss); LDP(n_salv); LDP(ha); LDP(z_gw); LDP(kf); LDP(ks_ss);
    if (LAT) {
        LDP(dmph); LDP(slope); LDP(v_mp_layer_1); LDP(v_mp_layer_2); LDP(v_mp_layer_3); LDP(v_mp_layer_4);
        LDP(v_mp_layer_5); LDP(v_mp_layer_6); LDP(v_mp_layer_7); LDP(v_mp_layer_8);
    }
}
template <bool FLUX>
RD void load_storages(Cell& c, const Io& io) {
    LDP(lambda_bc); LDP(m_bc); LDS(S);
    c.S_m1 = c.S;
    if (FLUX) { c.S_rz_m1 = io.lds(RF_S_rz); c.S_ss_m1 = io.lds(RF_S_ss); c.S_s_m1 = io.lds(RF_S_s); }
    else { c.S_rz_m1 = 0; c.S_ss_m1 = 0; c.S_s_m1 = 0; }
}

#ifndef ROGER_PREFETCH
#define ROGER_PREFETCH 0
#endif
// Register-free alternative: ask the memory system to pull the planes of the later processes into L2 (or L1)
// while the first processes compute.
template <bool LAT>
RD void prefetch_later_groups(const StepPlan& p, const Io& io) {
#define PF(name) io.pf(RF_##name)
    PF(sealing);
    if (!p.cond1) {
        PF(no_wf); PF(z_wf); PF(z_wf_t0); PF(z_wf_t1); PF(z_wf_fc); PF(inf_mat_event_csum); PF(inf_mat_pot_event_csum);
        PF(inf_mp_event_csum); PF(y_mp); PF(inf_sc_event_csum); PF(y_sc); PF(theta_d); PF(theta_d_rel); PF(theta_d_t0);
        PF(theta_d_rel_t0); PF(theta_d_fp); PF(t_event_csum); PF(z_sc);
    }
    PF(ks); PF(wfs); PF(t_sat); PF(lmpv); PF(dmpv); PF(mp_drain_area); PF(inf_mat_rz); PF(S_lp_ss); PF(S_fp_ss); PF(maskRiverLake);
    PF(z_sat); PF(k_rz); PF(h_rz); PF(h_ss); PF(n_salv); PF(ha); PF(z_gw); PF(kf); PF(ks_ss);
    if (LAT) {
        PF(dmph); PF(slope); PF(v_mp_layer_1); PF(v_mp_layer_2); PF(v_mp_layer_3); PF(v_mp_layer_4); PF(v_mp_layer_5);
        PF(v_mp_layer_6); PF(v_mp_layer_7); PF(v_mp_layer_8);
    }
    PF(lambda_bc); PF(m_bc); PF(S);
#undef PF
}

template <bool LAT, bool FLUX, bool MASK1, bool GWB = false>
RD bool cell_step(Cell& c, const StepPlan& p, const Io& io) {
    derived_parameters<MASK1>(c, io);
    c.pet_res = c.pet;  // adaptive_time_stepping.py:367-368
    load_interception(c, io);
    if (ROGER_PREFETCH) load_evapotranspiration(c, io);
    if (ROGER_L2_PREFETCH) prefetch_later_groups<LAT>(p, io);
    interception<FLUX, MASK1>(c, p, io);
    if (!ROGER_PREFETCH) load_evapotranspiration(c, io);
    if (ROGER_PREFETCH) load_event_state(c, p, io);
    evapotranspiration<FLUX, MASK1>(c, p, io);
    snow<FLUX, MASK1>(c, p, io);
    if (!ROGER_PREFETCH) load_event_state(c, p, io);
    if (ROGER_PREFETCH) load_infiltration(c, io);
    event_bookkeeping_start<FLUX, MASK1>(c, p, io);
    if (!ROGER_PREFETCH) load_infiltration(c, io);
    green_ampt_and_matrix<FLUX, MASK1>(c, p, io);
    if (ROGER_PREFETCH) load_subsurface<LAT>(c, io);
    macropores<FLUX, MASK1>(c, p, io);
    shrinkage_cracks<FLUX, MASK1>(c, p, io);
    infiltration_tail<FLUX, MASK1>(c, p, io);
    if (!ROGER_PREFETCH) load_subsurface<LAT>(c, io);
    if (ROGER_PREFETCH) load_storages<FLUX>(c, io);
    subsurface<LAT, FLUX, MASK1>(c, p, io);
    if (GWB) capillary_rise_subsoil<FLUX, MASK1>(c, p, io);
    if (!ROGER_PREFETCH) load_storages<FLUX>(c, io);
    storages<FLUX, MASK1>(c, p, io);
    if (GWB) groundwater_recharge<FLUX, MASK1>(c, p, io);
    bool ok = balance_check<LAT, FLUX, GWB>(c, p, io);
    if (p.clamp) {  // models/svat/svat.py:333-352
        c.S_fp_rz = ((c.S_fp_rz > -1e-6) && (c.S_fp_rz < 0)) ? 0.0 : c.S_fp_rz;
        c.S_lp_rz = ((c.S_lp_rz > -1e-6) && (c.S_lp_rz < 0)) ? 0.0 : c.S_lp_rz;
        c.S_fp_ss = ((c.S_fp_ss > -1e-6) && (c.S_fp_ss < 0)) ? 0.0 : c.S_fp_ss;
        c.S_lp_ss = ((c.S_lp_ss > -1e-6) && (c.S_lp_ss < 0)) ? 0.0 : c.S_lp_ss;
    }
    STS(S_fp_rz); STS(S_lp_rz); STS(S_fp_ss); STS(S_lp_ss);
    return ok;
}

#undef M
}  // namespace roger
