/* svat_oracle.c -- CPU restatement of the reference's per-cell SVAT / oneD time step.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, load or call this file.  The product path
 * (roger_b200/csrc/ -> libroger_b200.so) never links it and has no CPU fallback.
 *
 * Parity status: PINNED against outputs of the reference itself.  tests/golden/make_svat_golden.py
 * imports the unmodified NumPy backend from /root/reference in the build container, runs seeded
 * synthetic cases and commits per-step dumps (tests/golden/svat_*.npz); tests/test_oracle_golden.py
 * replays them through this file.  The reference's own tests hold no golden vectors for this path
 * (test/setup_test.py asserts nothing numeric).
 *
 * Structure follows roger/roger.py:389-485 (RogerSetup.step) kernel by kernel; every function
 * cites the reference lines it restates.  Scalar C, one cell at a time, with the grid-wide
 * reductions of adaptive_time_stepping done in a pre-pass.  All `where` chains, mask-capture
 * points, operator-precedence slips and NaN semantics of the reference are kept on purpose.
 *
 * Build: see oracle/Makefile (gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/roger_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif
/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline legs of bench.py ask for all host cores */
void roger_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Which pow the physics calls (see sas_oracle.c): 0 the C library's pow -- the mode the oracle is pinned in against the
 * reference fixtures -- 1 roger_pow(), the fixed IEEE operation sequence the CUDA kernels evaluate, for bit-for-bit
 * comparisons with the GPU. */
extern int g_roger_oracle_pow_mode;
double roger_pow_host(double x, double y);
static double opow(double x, double y) { return g_roger_oracle_pow_mode ? roger_pow_host(x, y) : pow(x, y); }

#define NSLOT ROGER_SLOTS_PER_DAY
#define V(name) f[RF_##name]

/* ---------------------------------------------------------------------------------- */
/* numpy's float64 add.reduce over a contiguous axis: pairwise summation with 8-lane   */
/* unrolling (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum); used by    */
/* npx.sum / npx.nanmean in adaptive_time_stepping.py:405-435.                          */
static double np_pairwise_sum(const double *a, int n)
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
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}
static double np_sum144(const double *a) { return 0.0 + np_pairwise_sum(a, NSLOT); }

/* adaptive_time_stepping.py:396-426 calc_prec_ta_pet_hourly: sum / nanmean of the slots
 * [itt_day, itt_day+6) selected with npx.where(..., x, 0) over the whole 144-slot row */
static double window_sum(const double *row, int64_t itt_day)
{
    double tmp[NSLOT];
    for (int k = 0; k < NSLOT; k++) tmp[k] = (k >= itt_day && k < itt_day + 6) ? row[k] : 0.0;
    return np_sum144(tmp);
}
static double window_nanmean(const double *row, int64_t itt_day)
{
    double tmp[NSLOT];
    int64_t cnt = 0;
    for (int k = 0; k < NSLOT; k++) {
        double v = (k >= itt_day && k < itt_day + 6) ? row[k] : NAN;
        if (isnan(v)) v = 0.0; else cnt++;
        tmp[k] = v;
    }
    return np_sum144(tmp) / (double)cnt;
}
static double day_nanmean(const double *row)
{
    double tmp[NSLOT];
    int64_t cnt = 0;
    for (int k = 0; k < NSLOT; k++) {
        double v = row[k];
        if (isnan(v)) v = 0.0; else cnt++;
        tmp[k] = v;
    }
    return np_sum144(tmp) / (double)cnt;
}

/* per-step context shared by all cells */
typedef struct {
    const roger_svat_settings *st;
    double dt;
    int cond1, cond2_scalar, cond3_scalar, cond4, cond5; /* infiltration.py:2155-2167 */
    int substeps;
    int irrigation_month;
} step_ctx;

/* ================================================================================== */
/* interception.py                                                                    */
/* ================================================================================== */
static void calc_rain_int_top(double *f, const step_ctx *c) /* interception.py:6-71 */
{
    const double m = V(maskCatch), ta_fm = c->st->ta_fm;
    int mask_rain = V(ta) > ta_fm;
    V(rain_top) = (mask_rain ? V(prec) : 0) * m;
    double wtmx = (10000. / (100 - c->st->rmax) / 100.) * V(swe_top);
    double S_int_top_tot = ((V(S_int_top_tot) < wtmx) ? wtmx : V(S_int_top_tot)) * m;
    double int_top_free = ((V(S_int_top) < S_int_top_tot) ? S_int_top_tot - V(S_int_top) : 0) * m;
    int mask1 = (int_top_free >= V(prec) * (1. - V(throughfall_coeff_top))) && (V(ta) > ta_fm) && (int_top_free > 0);
    int mask2 = (int_top_free < V(prec) * (1. - V(throughfall_coeff_top))) && (V(ta) > ta_fm) && (int_top_free > 0);
    V(int_rain_top) = 0;
    V(int_rain_top) += V(prec) * (1. - V(throughfall_coeff_top)) * mask1 * m;
    V(int_rain_top) = (mask2 ? int_top_free : V(int_rain_top)) * m;
    V(S_int_top) += V(int_rain_top) * m;
}

static void calc_rain_int_ground(double *f, const step_ctx *c) /* interception.py:74-151 */
{
    const double m = V(maskCatch), ta_fm = c->st->ta_fm;
    int mask_rain = V(ta) > ta_fm;
    double rain = (V(prec) - V(int_rain_top)) * mask_rain * m;
    double int_ground_free =
        (((V(S_int_ground) < V(S_int_ground_tot)) && (V(S_snow) <= 0)) ? V(S_int_ground_tot) - V(S_int_ground) : 0) * m;
    int mask1 = (int_ground_free >= rain * (1. - V(throughfall_coeff_ground))) && (V(ta) > ta_fm) && (int_ground_free > 0);
    int mask2 = (int_ground_free < rain * (1. - V(throughfall_coeff_ground))) && (V(ta) > ta_fm) && (int_ground_free > 0);
    V(int_rain_ground) = 0;
    V(int_rain_ground) += rain * (1. - V(throughfall_coeff_ground)) * mask1 * m;
    V(int_rain_ground) = (mask2 ? int_ground_free : V(int_rain_ground)) * m;
    V(int_rain_ground) = ((V(lu_id) == 599) ? 0 : V(int_rain_ground)) * m;
    V(S_int_ground) += V(int_rain_ground) * m;
    V(rain_ground) = (V(rain_top) - V(int_rain_top) - V(int_rain_ground)) * m;
    V(z0) += ((V(S_snow) > 0) ? 0 : V(rain_ground)) * m;
    V(prec_event_csum) += ((V(S_snow) > 0) ? 0 : V(rain_ground)) * m;
}

static void calc_snow_int_top(double *f, const step_ctx *c) /* interception.py:154-245 */
{
    const double m = V(maskCatch), ta_fm = c->st->ta_fm;
    const double ta = V(ta), lu = V(lu_id);
    int mask_snow = ta <= ta_fm;
    V(snow_top) = (mask_snow ? V(prec) : 0) * m;
    V(swe_top_tot) = (((ta > -1) && (lu == 10)) ? 9 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta > -1) && (lu == 11)) ? 15 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta > -1) && (lu == 12)) ? 25 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta >= -3) && (ta <= -1) && (lu == 10)) ? 2.5 + 0.5 * ta * 9 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta >= -3) && (ta <= -1) && (lu == 11)) ? 2.5 + 0.5 * ta * 15 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta >= -3) && (ta <= -1) && (lu == 12)) ? 2.5 + 0.5 * ta * 25 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta < -3) && (lu == 10)) ? 18 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta < -3) && (lu == 11)) ? 30 : V(swe_top_tot)) * m;
    V(swe_top_tot) = (((ta < -3) && (lu == 12)) ? 50 : V(swe_top_tot)) * m;
    double int_top_free = ((V(swe_top) >= V(swe_top_tot)) ? 0 : V(swe_top_tot) - V(swe_top)) * m;
    int mask1 = (int_top_free >= V(prec) * (1. - V(throughfall_coeff_top))) && (ta <= ta_fm) && (int_top_free > 0);
    int mask2 = (int_top_free < V(prec) * (1. - V(throughfall_coeff_top))) && (ta <= ta_fm) && (int_top_free > 0);
    V(int_snow_top) = 0;
    V(int_snow_top) += V(prec) * (1. - V(throughfall_coeff_top)) * mask1 * m;
    V(int_snow_top) = (mask2 ? int_top_free : V(int_snow_top)) * m;
    V(S_int_top) += V(int_snow_top) * m;
    V(swe_top) += V(int_snow_top) * m;
}

static void calc_snow_int_ground(double *f, const step_ctx *c) /* interception.py:248-318 */
{
    const double m = V(maskCatch), ta_fm = c->st->ta_fm;
    int mask_snow = V(ta) <= ta_fm;
    double snow = (V(prec) - V(int_snow_top)) * mask_snow * m;
    double int_ground_free = ((V(S_int_ground) >= V(S_int_ground_tot)) ? 0 : V(S_int_ground_tot) - V(S_int_ground)) * m;
    int mask1 = (int_ground_free >= snow * (1. - V(throughfall_coeff_ground))) && (V(ta) <= ta_fm) && (int_ground_free > 0);
    int mask2 = (int_ground_free < snow * (1. - V(throughfall_coeff_ground))) && (V(ta) <= ta_fm) && (int_ground_free > 0);
    V(int_snow_ground) = 0;
    V(int_snow_ground) += snow * (1. - V(throughfall_coeff_ground)) * mask1 * m;
    V(int_snow_ground) = (mask2 ? int_ground_free : V(int_snow_ground)) * m;
    V(int_snow_ground) = ((V(lu_id) == 599) ? 0 : V(int_snow_ground)) * m;
    V(S_int_ground) += V(int_snow_ground) * m;
    V(swe_ground) += V(int_snow_ground) * m;
    V(snow_ground) = (V(snow_top) - V(int_snow_top) - V(int_snow_ground)) * m;
    V(prec_event_csum) += V(snow_ground) * m;
}

static void calc_int(double *f) /* interception.py:321-343 */
{
    const double m = V(maskCatch);
    V(int_top) = (V(int_rain_top) + V(int_snow_top)) * m;
    V(int_ground) = (V(int_rain_ground) + V(int_snow_ground)) * m;
    V(int_prec) = (V(int_rain_top) + V(int_rain_ground) + V(int_snow_top) + V(int_snow_ground)) * m;
}

/* ================================================================================== */
/* evapotranspiration.py                                                              */
/* ================================================================================== */
static void calc_evap_int_top(double *f) /* evapotranspiration.py:9-66 */
{
    const double m = V(maskCatch);
    int mask1 = (V(S_int_top) <= V(S_int_top_tot)) && (V(pet_res) <= V(S_int_top)) && (V(S_int_top_tot) > 0) && (V(S_int_top) > 0);
    int mask2 = (V(S_int_top) <= V(S_int_top_tot)) && (V(pet_res) > V(S_int_top)) && (V(S_int_top_tot) > 0) && (V(S_int_top) > 0);
    V(evap_int_top) = 0;
    V(evap_int_top) += V(pet_res) * mask1 * m;
    V(pet_res) = (mask1 ? 0 : V(pet_res)) * m;
    V(evap_int_top) += V(S_int_top) * mask2 * m;
    V(pet_res) += -V(S_int_top) * mask2 * m;
    V(S_int_top) += -V(evap_int_top) * m;
}

static void calc_evap_int_ground(double *f) /* evapotranspiration.py:69-134 */
{
    const double m = V(maskCatch);
    int mask1 = (V(S_int_ground) <= V(S_int_ground_tot)) && (V(pet_res) <= V(S_int_ground)) && (V(S_int_ground_tot) > 0) && (V(S_int_ground) > 0);
    int mask2 = (V(S_int_ground) <= V(S_int_ground_tot)) && (V(pet_res) > V(S_int_ground)) && (V(S_int_ground_tot) > 0) && (V(S_int_ground) > 0);
    V(evap_int_ground) = 0;
    V(evap_int_ground) += V(pet_res) * mask1 * m;
    V(pet_res) = (mask1 ? 0 : V(pet_res)) * m;
    V(evap_int_ground) += V(S_int_ground) * mask2 * m;
    V(pet_res) += -V(S_int_ground) * mask2 * m;
    V(S_int_ground) += -V(evap_int_ground) * m;
    V(evap_int) = V(evap_int_ground) + V(evap_int_top) * m; /* :129 mask binds to last term */
}

static void calc_evap_dep(double *f) /* evapotranspiration.py:137-194 */
{
    const double m = V(maskCatch);
    int mask1 = (V(S_dep) <= V(pet_res)) && (V(S_dep) > 0) && (V(pet_res) > 0) && (V(prec) <= 0);
    int mask2 = (V(S_dep) > V(pet_res)) && (V(S_dep) > 0) && (V(pet_res) > 0) && (V(prec) <= 0);
    V(evap_dep) = 0;
    V(evap_dep) += V(S_dep) * mask1 * m;
    V(pet_res) += -V(S_dep) * mask1 * m;
    V(evap_dep) += V(pet_res) * mask2 * m;
    V(pet_res) = (mask2 ? 0 : V(pet_res)) * m;
    int mask3 = (V(S_dep) > 0) && (V(evap_dep) > 0);
    V(S_dep) += -V(evap_dep) * mask3 * m;
}

static void calc_evap_sur(double *f) /* evapotranspiration.py:197-212 */
{
    V(evap_sur) = V(evap_int_top) + V(evap_int_ground) + V(evap_dep) * V(maskCatch);
}

static void calc_evap_soil(double *f) /* evapotranspiration.py:215-345 */
{
    const double m = V(maskCatch);
    int mask3 = V(de) <= V(rew);
    int mask4 = (V(de) > V(rew)) && (V(de) <= V(tew));
    int mask5 = V(de) > V(tew);
    V(k_stress_evap) = (mask3 ? 1 : V(k_stress_evap)) * m;
    V(k_stress_evap) = (mask4 ? (V(tew) - V(de)) / (V(tew) - V(rew)) : V(k_stress_evap)) * m;
    V(k_stress_evap) = (mask5 ? 0 : V(k_stress_evap)) * m;
    V(evap_coeff) = V(basal_evap_coeff) * V(k_stress_evap) * m;
    double pevap_soil = V(pet_res) * V(evap_coeff) * m;
    V(pevap_soil) = V(pet_res) * V(evap_coeff) * m;
    int mask1 = (V(S_fp_rz) > 0) && (pevap_soil <= V(S_fp_rz)) && (pevap_soil > 0) && (V(swe) <= 0) && (V(prec) <= 0);
    int mask2 = (V(S_fp_rz) > 0) && (pevap_soil > V(S_fp_rz)) && (pevap_soil > 0) && (V(swe) <= 0) && (V(prec) <= 0);
    double evap_fp = 0;
    evap_fp += pevap_soil * mask1 * m;
    V(pet_res) += -pevap_soil * mask1 * m;
    V(pet_res) = ((V(pet_res) < 0) ? 0 : V(pet_res)) * m;
    evap_fp += V(S_fp_rz) * mask2 * m;
    V(pet_res) += -V(S_fp_rz) * mask2 * m;
    V(pet_res) = ((V(pet_res) < 0) ? 0 : V(pet_res)) * m;
    V(evap_soil) = evap_fp * m;
    V(S_fp_rz) += -V(evap_soil) * m;
}

static int is_crop(double lu) { return lu >= 500 && lu < 600 && lu == floor(lu); }
static int is_tree_lu(double lu) { return lu == 10 || lu == 11 || lu == 12 || lu == 15 || lu == 16 || lu == 17; }

static void calc_transp(double *f, const step_ctx *c) /* evapotranspiration.py:348-543 */
{
    const double m = V(maskCatch);
    double theta_water_stress = c->st->transp_water_stress * V(theta_ufc) + V(theta_pwp) * m; /* :360 */
    int mask_crops = is_crop(V(lu_id));
    V(k_stress_transp) = (mask_crops ? V(k_stress_transp)
                                     : ((V(theta_rz) - V(theta_pwp)) / (theta_water_stress - V(theta_pwp)))) * m;
    V(k_stress_transp) = (V(k_stress_transp) > 1) ? 1 : V(k_stress_transp);
    V(transp_coeff) = V(basal_transp_coeff) * V(k_stress_transp) * m;
    int mask_anoxia = (V(lu_id) > 500) && (V(lu_id) < 599) && (V(theta_rz) >= 0.8 * V(theta_sat));
    if (mask_anoxia) {
        double r = V(S_lp_rz) / V(S_ac_rz);
        V(transp_coeff) = (((r >= 0) && (r <= 1)) ? 1 - opow(r, 1.5) : 1) * m;
    } else {
        V(transp_coeff) = V(transp_coeff) * m;
    }
    double _pt = ((V(pevap_soil) < V(pet)) ? V(pet) - V(pevap_soil) : 0) * m;
    double _ptransp = ((V(evap_soil) < V(pet)) ? V(pet) - V(evap_soil) : 0) * m;
    V(pt) = _pt * V(basal_transp_coeff) * m;
    V(ptransp) = _ptransp * V(transp_coeff) * m;
    V(ptransp) = (is_tree_lu(V(lu_id)) ? V(pet) * V(transp_coeff) : V(ptransp)) * m;
    V(ptransp_res) = V(ptransp) * m;
    double transp_lp = 0, transp_fp = 0;
    int mask1 = (V(S_lp_rz) > 0) && (V(ptransp_res) <= V(S_lp_rz)) && (V(ptransp) > 0) && (V(prec) <= 0);
    transp_lp += (mask1 ? V(ptransp_res) : 0) * m;
    V(ptransp_res) = (mask1 ? 0 : V(ptransp_res)) * m;
    int mask2 = (V(S_lp_rz) > 0) && (V(ptransp_res) > V(S_lp_rz)) && (V(ptransp) > 0) && (V(prec) <= 0);
    transp_lp += (mask2 ? V(S_lp_rz) : 0) * m;
    V(ptransp_res) += (mask2 ? -V(S_lp_rz) : 0) * m;
    int mask3 = (V(S_fp_rz) > 0) && (V(ptransp_res) <= V(S_fp_rz)) && (V(S_lp_rz) <= 0) && (V(ptransp) > 0) && (V(prec) <= 0);
    transp_fp += (mask3 ? V(ptransp_res) : 0) * m;
    V(ptransp_res) = (mask3 ? 0 : V(ptransp_res)) * m;
    int mask4 = (V(S_fp_rz) > 0) && (V(ptransp_res) > V(S_fp_rz)) && (V(S_lp_rz) <= 0) && (V(ptransp) > 0) && (V(prec) <= 0);
    transp_fp += (mask4 ? V(S_fp_rz) : 0) * m;
    V(ptransp_res) += (mask4 ? -V(S_fp_rz) : 0) * m;
    V(ptransp_res) = ((V(ptransp_res) < 0) ? 0 : V(ptransp_res)) * m;
    V(S_lp_rz) += -transp_lp * m;
    V(S_fp_rz) += -transp_fp * m;
    V(transp) = (transp_fp + transp_lp) * m;
}

static void calc_et_tail(double *f) /* evapotranspiration.py:546-599 */
{
    const double m = V(maskCatch);
    V(de) += V(evap_soil) + V(transp) * (V(z_evap) / V(z_root)) * m; /* :557 */
    V(aet_soil) = (V(evap_soil) + V(transp)) * m;
    V(aet) = (V(evap_int_top) + V(evap_int_ground) + V(evap_dep) + V(evap_soil) + V(transp)) * m;
}

/* ================================================================================== */
/* snow.py                                                                            */
/* ================================================================================== */
static void calc_snow(double *f, const step_ctx *c) /* snow.py:6-290 */
{
    const double m = V(maskCatch), ta_fm = c->st->ta_fm, sf = c->st->sf, dt = c->dt;
    const double rf = (10000. / (100 - c->st->rmax) / 100.);
    /* calc_snow_accumulation :6-26 */
    {
        int mask1 = V(ta) <= ta_fm;
        V(S_snow) += V(snow_ground) * mask1 * m;
        V(swe) += V(snow_ground) * mask1 * m;
    }
    /* calc_rain_on_snow :29-44 */
    {
        int mask1 = (V(swe) > 0) && (V(ta) > ta_fm);
        V(S_snow) += V(rain_ground) * mask1 * m;
    }
    /* calc_snow_melt_int_top :47-133 */
    {
        double pot = (sf * (V(ta) - ta_fm) * dt) * m;
        int mask1 = (pot > 0) && (pot <= V(swe_top)) && (V(swe_top) > 0);
        int mask2 = (pot > 0) && (pot > V(swe_top)) && (V(swe_top) > 0);
        V(snow_melt_top) = 0;
        V(snow_melt_top) = (mask1 ? pot : V(snow_melt_top)) * m;
        V(snow_melt_top) = (mask2 ? V(swe_top) : V(snow_melt_top)) * m;
        int mask4 = (V(snow_melt_top) > 0) && (V(snow_melt_top) <= V(swe_top));
        int mask5 = (V(snow_melt_top) > 0) && (V(snow_melt_top) > V(swe_top));
        V(pet_res) += -V(snow_melt_top) * mask4 * m;
        V(swe_top) += -V(snow_melt_top) * mask4 * m;
        V(pet_res) += -V(swe_top) * mask5 * m;
        V(swe_top) += (mask5 ? 0 : -V(swe_top)) * m; /* :98-101 zeroes swe_top unless mask5 */
        V(pet_res) = ((V(pet_res) < 0) ? 0 : V(pet_res)) * m;
        double wtmx = rf * V(swe_top);
        double q_ret = ((V(S_int_top) > V(S_int_top_tot)) ? V(S_int_top) - V(swe_top) : 0) * m;
        V(snow_melt_drip) = ((q_ret > wtmx) ? q_ret - wtmx
                             : (((wtmx <= 0) && (V(S_int_top_tot) < V(S_int_top))) ? V(S_int_top) - V(S_int_top_tot) : 0)) * m;
        int mask6 = V(S_int_top_tot) < V(S_int_top);
        V(S_snow) += (mask6 ? V(snow_melt_drip) : 0) * m;
        V(S_int_top) += (mask6 ? -V(snow_melt_drip) : 0) * m;
    }
    /* calc_snow_melt_ground_int :136-193 */
    {
        double pot = (sf * (V(ta) - ta_fm) * dt) * m;
        int mask1 = (pot > 0) && (pot <= V(swe_ground)) && (V(swe_ground) > 0);
        int mask2 = (pot > 0) && (pot > V(swe_ground)) && (V(swe_ground) > 0);
        V(snow_melt_ground) = 0;
        V(snow_melt_ground) = (mask1 ? pot : V(snow_melt_ground)) * m;
        V(snow_melt_ground) = (mask2 ? V(swe_ground) : V(snow_melt_ground)) * m;
        int mask4 = (V(snow_melt_ground) > 0) && (V(snow_melt_ground) <= V(swe_ground));
        int mask5 = (V(snow_melt_ground) > 0) && (V(snow_melt_ground) > V(swe_ground));
        V(pet_res) += -V(snow_melt_ground) * mask4 * m;
        V(swe_ground) += -V(snow_melt_ground) * mask4 * m;
        V(pet_res) += -V(swe_ground) * mask5 * m;
        V(swe_ground) += (mask5 ? 0 : -V(swe_ground)) * m;
    }
    /* calc_snow_melt :196-290 */
    {
        double pot = (sf * (V(ta) - ta_fm) * dt) * m;
        int mask1 = (pot > 0) && (pot <= V(swe)) && (V(swe) > 0);
        int mask2 = (pot > 0) && (pot > V(swe)) && (V(swe) > 0);
        V(snow_melt) = 0;
        V(snow_melt) = (mask1 ? pot : V(snow_melt)) * m;
        V(snow_melt) = (mask2 ? V(swe) : V(snow_melt)) * m;
        int mask5 = (V(snow_melt) > 0) && (V(snow_melt) <= V(swe));
        int mask6 = (V(snow_melt) > 0) && (V(snow_melt) > V(swe));
        V(pet_res) += -V(snow_melt) * mask5 * m;
        V(swe) += -V(snow_melt) * mask5 * m;
        V(pet_res) += -V(swe) * mask6 * m;
        V(swe) = (mask6 ? 0 : V(swe)) * m;
        V(pet_res) = ((V(pet_res) < 0) ? 0 : V(pet_res)) * m;
        double wtmx = rf * V(swe);
        double q_ret = ((V(S_snow) > 0) ? V(S_snow) - V(swe) : 0) * m;
        V(q_snow) = 0;
        V(q_snow) = ((q_ret > wtmx) ? q_ret - wtmx : ((wtmx <= 0) ? V(S_snow) : 0)) * m;
        V(S_snow) += -V(q_snow) * m;
        V(z0) += V(q_snow) * m;
        V(prec_event_csum) += V(q_snow) * m;
    }
}

/* ================================================================================== */
/* infiltration.py                                                                    */
/* ================================================================================== */
static double calc_theta_d(const double *f) /* infiltration.py:1563-1594 */
{
    const double m = V(maskCatch);
    double theta_d = 0;
    theta_d = ((V(z_root) > 0) ? (V(theta_sat) - V(theta_rz)) * (1 - V(sealing) / 1) : theta_d) * m;
    theta_d = ((V(z_soil) <= 0) ? 0.01 : theta_d) * m;
    theta_d = ((theta_d <= 0) ? 0.01 : theta_d) * m;
    return theta_d;
}
static double calc_theta_d_rel(const double *f) /* infiltration.py:1597-1632 */
{
    const double m = V(maskCatch);
    double t = 0;
    t = ((V(z_root) > 0) ? ((V(theta_sat) - V(theta_rz)) / (V(theta_sat) - V(theta_pwp))) * (1 - V(sealing) / 1) : t) * m;
    t = ((V(z_soil) <= 0) ? 0.01 : t) * m;
    t = ((t <= 0) ? 0.01 : t) * m;
    return t;
}
static double calc_theta_d_fp(const double *f) /* infiltration.py:1635-1666 */
{
    const double m = V(maskCatch);
    double t = 0;
    t = ((V(z_soil) > 0) ? (V(theta_fc) - V(theta_rz)) * (1 - V(sealing) / 1) : t) * m;
    t = ((V(z_soil) <= 0) ? 0.01 : t) * m;
    t = ((t <= 0) ? 0.01 : t) * m;
    return t;
}

static void calc_depth_shrinkage_cracks(double *f) /* infiltration.py:1767-1826 */
{
    const double m = V(maskCatch), th = V(theta_rz);
    V(z_sc) = ((th < V(theta_4)) ? V(z_sc_max)
               : (((th >= V(theta_4)) && (th < V(theta_27))) ? (th - V(theta_4)) / (V(theta_27) - V(theta_4)) : 0) * V(z_sc_max)) * m;
    V(z_sc) = ((th < V(theta_4)) ? V(z_sc_max) : V(z_sc)) * m;
    V(z_sc) = ((th > V(theta_27)) ? 0 : V(z_sc)) * m;
    V(z_sc) = ((1 - V(sealing) / 1) * V(z_sc)) * m;
    V(z_sc) = ((V(z_sc) > V(z_root)) ? V(z_root) : V(z_sc)) * m;
    V(z_sc) = ((V(lu_id) == 13) ? 0 : V(z_sc)) * m;
}

/* per-cell [taum1] companions that the step reads (the reference copies tau->taum1 after every step) */
typedef struct {
    double z_wf, z_wf_t0, z_wf_t1, y_mp, y_sc, prec, S, S_rz, S_ss, S_s, z_root;
} cell_m1;

static void set_event_vars(double *f, cell_m1 *p) /* infiltration.py:1829-1976 */
{
    const double m = V(maskCatch);
    V(no_wf) = 1;
    V(z_wf) = 0; p->z_wf = 0;
    V(z_wf_t0) = 0; p->z_wf_t0 = 0;
    V(z_wf_t1) = 0; p->z_wf_t1 = 0;
    V(z_wf_fc) = 0;
    V(inf_mat_event_csum) = 0;
    V(inf_mat_pot_event_csum) = 0;
    V(inf_mp_event_csum) = 0;
    V(y_mp) = 0; p->y_mp = 0;
    V(inf_sc_event_csum) = 0;
    V(y_sc) = 0; p->y_sc = 0;
    double theta_d = calc_theta_d(f);
    V(theta_d) = theta_d * m;
    double theta_d_rel = calc_theta_d_rel(f);
    V(theta_d_rel) = theta_d_rel * m;
    V(theta_d_t0) = theta_d * m;
    V(theta_d_rel_t0) = theta_d_rel * m;
    V(theta_d_fp) = calc_theta_d_fp(f) * m;
    V(prec_event_csum) = 0;
    V(t_event_csum) = 0;
    V(de) = 0;
}

static void set_event_vars_start_rainfall_pause(double *f, const cell_m1 *p) /* infiltration.py:1979-1995, :1535-1560 */
{
    const double m = V(maskCatch);
    double z_wf_fc = ((V(theta_d_fp) > 0) ? V(inf_mat_event_csum) / V(theta_d_fp) : V(z_wf)) * m;
    z_wf_fc = ((z_wf_fc > V(z_soil)) ? V(z_soil) : z_wf_fc) * m;
    int mask = (V(prec) == 0) && (p->prec != 0);
    V(z_wf_fc) = (mask ? z_wf_fc : V(z_wf_fc)) * m;
}

static void set_event_vars_end_rainfall_pause(double *f, cell_m1 *p) /* infiltration.py:1998-2053 */
{
    const double m = V(maskCatch);
    int mask = (V(prec) != 0) && (p->prec == 0);
    V(no_wf) = mask ? 2 : V(no_wf);
    double theta_d = calc_theta_d(f);
    V(theta_d) = (mask ? theta_d : V(theta_d)) * m;
    double theta_d_rel = calc_theta_d_rel(f);
    V(theta_d_rel) = (mask ? theta_d_rel : V(theta_d_rel)) * m;
    if (mask) { V(z_wf_t1) = 0; p->z_wf_t1 = 0; }
    V(prec_event_csum) = mask ? 0 : V(prec_event_csum);
    V(t_event_csum) = mask ? 0 : V(t_event_csum);
}

static void reset_event_vars(double *f, cell_m1 *p) /* infiltration.py:2056-2144 */
{
    const double m = V(maskCatch);
    V(z_wf) = 0; p->z_wf = 0;
    V(z_wf_t0) = 0; p->z_wf_t0 = 0;
    V(z_wf_t1) = 0; p->z_wf_t1 = 0;
    V(y_mp) = 0; /* [tau] only, :2082-2086 */
    V(y_sc) = 0; p->y_sc = 0;
    double theta_d = calc_theta_d(f);
    V(theta_d) = theta_d * m;
    V(theta_d_t0) = theta_d * m;
    V(pi_gr) = 0;
    V(pi_m) = 0;
    V(t_sat) = 0;
    V(Fs) = 0;
    V(z_sc) = 0;
}

static void calc_green_ampt_params(double *f, const step_ctx *c) /* infiltration.py:7-48, :1669-1764 */
{
    const double m = V(maskCatch), dt = c->dt;
    double pi_gr = V(ks) * (((V(theta_d) * V(wfs)) / (V(prec_event_csum) + 1)) + 1);
    V(pi_gr) = pi_gr * m;
    double pi_m = V(ks) * V(theta_d) * V(wfs) * m;
    V(pi_m) = pi_m * m;
    /* calc_sat_time :1706-1740 (reads the freshly stored vs.pi_m / vs.pi_gr) */
    int mask1 = (V(pi_m) <= V(prec_event_csum)) && (V(pi_m) > V(pi_gr)) && (V(t_sat) == 0);
    int mask2 = ((V(prec) * (1 / dt) - V(ks)) * V(prec_event_csum) > V(ks) * V(theta_d) * V(wfs)) &&
                (V(pi_m) <= V(prec_event_csum)) && (V(pi_m) <= V(pi_gr)) && (V(t_sat) == 0);
    V(t_sat) = mask1 ? V(t_event_csum) - dt : V(t_sat);
    V(t_sat) = mask2 ? V(t_event_csum) + ((V(ks) * V(theta_d) * V(wfs)) / (V(pi_m) * (V(pi_m) * -V(ks)))) -
                           (dt / V(pi_m)) * V(prec_event_csum)
                     : V(t_sat);
    V(t_sat) = V(t_sat) * m;
    /* calc_Fs :1743-1764 */
    double Fs = ((V(ks) * V(theta_d) * V(wfs)) / (pi_m - V(ks))) * m;
    Fs = ((pi_m <= V(ks)) ? pi_m : Fs) * m;
    V(Fs) = Fs * m;
}

static void calc_inf_mat(double *f, const step_ctx *c, cell_m1 *p) /* infiltration.py:51-427 */
{
    const double m = V(maskCatch), dt = c->dt;
    const double Fs_t0 = 0.0, theta_d_t1 = 0.0; /* vs.Fs_t0 / vs.theta_d_t1 are never written by the reference */
    int mask1 = (V(pi_m) <= V(prec_event_csum)) && (V(t_event_csum) > V(t_sat)) && (V(t_sat) > 0);
    int mask2 = (V(pi_m) > V(prec_event_csum)) && (V(t_event_csum) > V(t_sat)) && (V(t_sat) > 0);
    int mask3 = (V(t_sat) > V(t_event_csum) - dt) && (V(t_sat) < V(t_event_csum));
    int mask4 = (V(pi_m) > V(prec_event_csum)) && (V(t_sat) <= 0);
    double a = V(ks) * (V(t_event_csum) - V(t_sat)) * m;
    double b = V(Fs) + 2 * V(theta_d) * V(wfs) * m; /* :77 */
    double l1 = ((V(z0) > V(ks) * dt) ? (V(ks) * dt * V(wfs) * V(theta_d)) / (V(z0) - V(ks) * dt)
                                      : (V(ks) * dt * V(wfs) * V(theta_d)) / (V(ks) * dt)) * m;
    double seal = ((1 - V(sealing)) / 1);
    V(inf_mat_pot) = V(ks) * dt;
    if (mask1)
        V(inf_mat_pot) = ((V(ks) * dt / 2) * (1 + (1 + 2 * b / a) / sqrt(1 + (4 * b / a) + (4 * (Fs_t0 * Fs_t0) / (a * a)))) * seal) * m;
    else
        V(inf_mat_pot) = V(inf_mat_pot) * m;
    if (mask2)
        V(inf_mat_pot) = (V(ks) * dt * (1 + ((V(wfs) * V(theta_d)) / l1)) * seal) * m;
    else
        V(inf_mat_pot) = V(inf_mat_pot) * m;
    double inf_mat_pot_rec = 0, inf_mat_pot_sat = 0;
    inf_mat_pot_rec = (mask3 ? (V(ks) * dt / 2) * (1 + (1 + 2 * b / a) / sqrt(1 + (4 * b / a) + (4 * (Fs_t0 * Fs_t0) / (a * a)))) : inf_mat_pot_rec) * m;
    inf_mat_pot_sat = (mask3 ? V(z0) * (V(t_sat) - (V(t_event_csum) - dt)) : inf_mat_pot_sat) * m;
    V(inf_mat_pot) = (mask3 ? inf_mat_pot_sat + inf_mat_pot_rec * seal : V(inf_mat_pot)) * m;
    V(inf_mat_pot) = (mask4 ? V(pi_gr) * seal : V(inf_mat_pot)) * m;
    /* matrix infiltration */
    int mask7 = V(z0) < V(inf_mat_pot);
    int mask8 = V(z0) >= V(inf_mat_pot);
    V(inf_mat) = (mask7 ? V(z0) : V(inf_mat)) * m;
    V(inf_mat) = (mask8 ? V(inf_mat_pot) : V(inf_mat)) * m;
    double cap = (V(S_ac_rz) + V(S_ufc_rz)) - (V(S_lp_rz) + V(S_fp_rz));
    V(inf_mat) = ((V(inf_mat) > cap) ? cap : V(inf_mat)) * m;
    V(inf_mat) = ((V(inf_mat) < 0) ? 0 : V(inf_mat)) * m;
    V(inf_mat_event_csum) += V(inf_mat) * m;
    V(inf_mat_pot_event_csum) += V(inf_mat_pot) * m;
    /* wetting front depth */
    double dz_wf = 0;
    int mask11 = V(no_wf) == 1, mask12 = V(no_wf) == 2;
    dz_wf = (mask11 ? (V(inf_mat) / V(theta_d_t0)) : dz_wf) * m;
    dz_wf = (mask12 ? V(inf_mat) / V(theta_d) : dz_wf) * m;
    V(z_wf_t0) += (isfinite(dz_wf) ? dz_wf : 0) * m;
    V(z_wf_t1) += (isfinite(dz_wf) ? dz_wf : 0) * m;
    V(z_wf_t0) = ((V(z_wf_t0) > V(z_soil)) ? V(z_soil) : V(z_wf_t0)) * m;
    V(z_wf_t1) = ((V(z_wf_t1) > V(z_soil)) ? V(z_soil) : V(z_wf_t1)) * m;
    V(z0) += -V(inf_mat) * m;
    V(z0) = ((V(z0) < 0) ? 0 : V(z0)) * m;
    double dz_wf_t0 = (((V(z_wf_fc) > 0) && (V(rain_ground) <= 0) && (V(no_wf) == 1)) ? V(inf_mat_pot) / V(theta_d_t0) : 0) * m;
    V(z_wf_t0) += (isfinite(dz_wf_t0) ? dz_wf_t0 : 0) * m;
    V(z_wf_t0) = (((V(z_wf_t0) > V(z_wf_fc)) && (V(z_wf_fc) > 0)) ? V(z_wf_fc) : V(z_wf_t0)) * m;
    V(z_wf_t0) = ((V(z_wf_t0) > V(z_soil)) ? V(z_soil) : V(z_wf_t0)) * m;
    double dz_wf_t1 = (((V(z_wf_fc) > 0) && (V(rain_ground) <= 0) && (V(no_wf) == 2)) ? V(inf_mat_pot) / V(theta_d) : 0) * m;
    V(z_wf_t1) += (isfinite(dz_wf_t1) ? dz_wf_t1 : 0) * m;
    V(z_wf_t1) = (((V(z_wf_t1) > V(z_wf_fc)) && (V(z_wf_fc) > 0)) ? V(z_wf_fc) : V(z_wf_t1)) * m;
    V(z_wf_t1) = ((V(z_wf_t1) > V(z_soil)) ? V(z_soil) : V(z_wf_t1)) * m;
    int mask14 = (V(z_wf_t0) >= V(z_wf_t1)) && (V(z_wf_t1) <= 0);
    int mask15 = (V(z_wf_t0) > V(z_wf_t1)) && (V(z_wf_t1) > 0);
    int mask20 = (V(z_wf_t0) <= V(z_wf_t1)) && (V(z_wf_t1) > 0);
    V(z_wf) = (mask14 ? V(z_wf_t0) : V(z_wf)) * m;
    V(theta_d) = (mask14 ? V(theta_d_t0) : V(theta_d)) * m;
    V(theta_d_rel) = (mask14 ? V(theta_d_rel_t0) : V(theta_d_rel)) * m;
    p->z_wf = (mask15 ? 0 : p->z_wf) * m;
    V(z_wf) = (mask15 ? V(z_wf_t1) : V(z_wf)) * m;
    V(no_wf) = mask20 ? 1 : V(no_wf);
    V(z_wf) = (mask20 ? V(z_wf_t0) : V(z_wf)) * m;
    V(theta_d) = (mask20 ? V(theta_d_t0) : V(theta_d)) * m;
    V(theta_d_rel) = (mask20 ? V(theta_d_rel_t0) : V(theta_d_rel)) * m;
    V(z_wf) = ((V(z_wf) > V(z_soil)) ? V(z_soil) : V(z_wf)) * m;
    V(theta_d) = ((theta_d_t1 <= 0) ? V(theta_d_t0) : V(theta_d)) * m; /* :411-416 always true */
}

static void calc_inf_mp(double *f, const step_ctx *c, const cell_m1 *p) /* infiltration.py:430-1077 */
{
    const double m = V(maskCatch), dt = c->dt, r_mp = c->st->r_mp, pi = c->st->pi;
    const int substeps = c->substeps;
    double z_wf, z_wf_m1, dz_wf;
    z_wf = ((V(no_wf) == 1) ? 0 : V(z_wf_t0)) * m;
    z_wf = ((V(no_wf) == 2) ? 0 : V(z_wf_t1)) * m;
    z_wf_m1 = ((V(no_wf) == 1) ? 0 : p->z_wf_t0) * m;
    z_wf_m1 = ((V(no_wf) == 2) ? 0 : p->z_wf_t1) * m;
    V(lmpv_non_sat) = V(lmpv) - z_wf * m;
    V(lmpv_non_sat) = ((V(lmpv_non_sat) < 0) ? 0 : V(lmpv_non_sat)) * m;
    dz_wf = z_wf - z_wf_m1 * m;
    int mask1 = z_wf >= V(lmpv);
    dz_wf = (mask1 ? V(lmpv_non_sat) : dz_wf) * m;
    dz_wf = ((V(lmpv_non_sat) <= 0) ? 0 : dz_wf) * m;
    dz_wf = ((dz_wf <= 0) ? 0 : dz_wf) * m;
    V(lmpv_non_sat) = V(lmpv) - V(z_wf) * m;
    V(lmpv_non_sat) = ((V(lmpv_non_sat) < 0) ? 0 : V(lmpv_non_sat)) * m;
    V(lmpv_non_sat) = ((substeps == 1) ? V(lmpv_non_sat) + dz_wf / 1.39 : V(lmpv_non_sat)) * m;

    double y1 = 0, y2, ym1, a, b1 = 0, b2 = 0, cc = 0, inf_mp_pot_di = 0, inf_mp_di = 0, z0_di = 0;
    double inf_mp = 0, inf_mp_pot = 0, inf_mp_event_csum, t = 0, y;
    y2 = (r_mp / 2) * m;
    a = V(theta_d) * (r_mp * r_mp) * m;
    y = p->y_mp * m;
    ym1 = p->y_mp * m;
    inf_mp_event_csum = V(inf_mp_event_csum) * m;
    const double sqrt6x2 = pow(6, 0.5) * 2;
    for (int i = 0; i < substeps; i++) { /* loop_body :582-777 */
        z0_di = V(z0) * (V(mp_drain_area) / substeps) * m;
        t += (dt / substeps) * m;
        cc = V(ks) * V(wfs) * t * m;
        cc = (isnan(cc) ? 0 : cc) * m;
        b1 = (sqrt6x2 * sqrt(cc * (6 * cc - a))) * m;
        b1 = (isnan(b1) ? 0 : b1) * m;
        b2 = (r_mp * (V(theta_d) * V(theta_d))) * (12 * cc - a + b1) * m;
        b2 = (isnan(b2) ? 0 : b2) * m;
        b2 = ((b2 <= 0) ? 0 : b2) * m;
        y1 = ((opow(b2, 1.0 / 3)) / V(theta_d)) * 0.5 * m;
        y2 = (a / (opow(b2, 1.0 / 3))) * 0.5 * m;
        y = (y1 + y2 + ym1) * m;
        y = ((y < r_mp) ? r_mp : y) * m;
        y = ((y < ym1) ? ym1 : y) * m;
        inf_mp_pot_di = (pi * (y * y - ym1 * ym1) * V(lmpv_non_sat) * V(theta_d) * V(dmpv) * 1e-06) * m;
        inf_mp_pot += inf_mp_pot_di * m;
        int mk = inf_mp_pot_di > z0_di;
        inf_mp_di = (mk ? z0_di : inf_mp_pot_di) * m;
        inf_mp_di = ((V(lmpv_non_sat) == 0) ? 0 : inf_mp_di) * m;
        inf_mp += inf_mp_di * m;
        inf_mp_event_csum += inf_mp_di * m;
        y = r_mp + sqrt((inf_mp_event_csum / (V(dmpv) * V(theta_d))) / pi) * m; /* :728-730 */
        y = ((y < r_mp) ? r_mp : y) * m;
        t = V(theta_d) / (V(ks) * V(wfs) * r_mp) * (opow(y, 3) / 3.0 - (y * y) * r_mp / 2.0 + (r_mp * r_mp * r_mp) / 6.0) * m;
        inf_mp = ((inf_mp < 0) ? 0 : inf_mp) * m;
        ym1 = y * m;
    }
    (void)y1; (void)y2; (void)inf_mp_pot;
    V(y_mp) = y * m;
    V(y_mp) = (isnan(V(y_mp)) ? 0 : V(y_mp)) * m;
    V(inf_mp) = inf_mp * m;
    V(inf_mp) = (isnan(V(inf_mp)) ? 0 : V(inf_mp)) * m;

    double rz_share_mp = ((V(lmpv_non_sat) > 0) ? 1.0 - (V(lmpv) - V(z_root)) / V(lmpv_non_sat) : 0) * m;
    rz_share_mp = ((V(lmpv) <= V(z_root)) ? 1 : rz_share_mp) * m;
    rz_share_mp = ((z_wf >= V(z_root)) ? 0 : rz_share_mp) * m;
    rz_share_mp = ((rz_share_mp < 0) ? 0 : rz_share_mp) * m;
    rz_share_mp = ((rz_share_mp > 1) ? 1 : rz_share_mp) * m;
    V(inf_mp_rz) = V(inf_mp) * rz_share_mp * m;
    /* NOTE vs.inf_mat_rz here is still the value of the PREVIOUS step (calc_inf_rz runs later) */
    double free_rz = (V(S_ac_rz) + V(S_ufc_rz)) - (V(inf_mat_rz) + V(S_lp_rz) + V(S_fp_rz));
    V(inf_mp_rz) = (((V(inf_mp_rz) > free_rz) && (free_rz >= 0)) ? free_rz : V(inf_mp_rz)) * m;
    V(inf_mp_rz) = ((free_rz < 0) ? 0 : V(inf_mp_rz)) * m;
    V(inf_mp_ss) = V(inf_mp) * (1 - rz_share_mp) * m;
    double free_ss = (V(S_ac_ss) + V(S_ufc_ss)) - (V(S_lp_ss) + V(S_fp_ss));
    V(inf_mp_ss) = (((V(inf_mp_ss) > free_ss) && (free_ss > 0)) ? free_ss : V(inf_mp_ss)) * m;
    V(inf_ss) = V(inf_mp_ss) * m;
    V(S_fp_ss) += V(inf_ss) * m;
    int mk = V(S_fp_ss) > V(S_ufc_ss);
    V(S_lp_ss) += (mk ? (V(S_fp_ss) - V(S_ufc_ss)) : 0) * m;
    V(S_fp_ss) = (mk ? V(S_ufc_ss) : V(S_fp_ss)) * m;
    mk = V(S_lp_ss) > V(S_ac_ss);
    V(inf_mp_ss) += (mk ? -(V(S_lp_ss) - V(S_ac_ss)) : 0) * m;
    V(inf_mp_ss) = ((V(inf_mp_ss) < 0) ? 0 : V(inf_mp_ss)) * m;
    V(S_lp_ss) = (mk ? V(S_ac_ss) : V(S_lp_ss)) * m;
    V(inf_mp) = 0;
    V(inf_mp) = V(inf_mp_rz) + V(inf_mp_ss) * m; /* :990 */
    V(inf_mp_event_csum) += V(inf_mp) * m;
    V(z0) += -V(inf_mp) * m;
    V(z0) = ((V(z0) < 0) ? 0 : V(z0)) * m;
    if (c->st->enable_macropore_lower_boundary_condition) { /* :1013-1065 */
        int mask_lbc = V(z_wf) > V(lmpv);
        double lim = V(ks) * dt * V(mp_drain_area);
        double add = (mask_lbc ? ((V(z0) > lim) ? lim : V(z0)) : 0);
        V(inf_mp) += add * m;
        double room = V(S_ac_ss) - V(S_lp_ss);
        V(inf_mp) = (((V(z_wf) > V(z_root)) && (V(inf_mp) > room) && (room > 0)) ? room : V(inf_mp)) * m;
        V(z0) += -(mask_lbc ? ((V(z0) > lim) ? lim : V(z0)) : 0) * m;
        V(z0) = ((V(z0) < 0) ? 0 : V(z0)) * m;
    }
}

static void calc_inf_sc(double *f, const step_ctx *c, const cell_m1 *p) /* infiltration.py:1080-1318 */
{
    const double m = V(maskCatch), dt = c->dt, l_sc = c->st->l_sc;
    const int substeps = c->substeps;
    double z_wf, z_wf_m1, dz_wf;
    z_wf = ((V(no_wf) == 1) ? 0 : V(z_wf_t0)) * m;
    z_wf = ((V(no_wf) == 2) ? 0 : V(z_wf_t1)) * m;
    z_wf_m1 = ((V(no_wf) == 1) ? 0 : p->z_wf_t0) * m;
    z_wf_m1 = ((V(no_wf) == 2) ? 0 : p->z_wf_t1) * m;
    V(z_sc_non_sat) = V(z_sc) - z_wf * m;
    V(z_sc_non_sat) = ((V(z_sc_non_sat) < 0) ? 0 : V(z_sc_non_sat)) * m;
    dz_wf = z_wf - z_wf_m1 * m;
    int mask1 = z_wf >= V(z_sc);
    dz_wf = (mask1 ? V(z_sc_non_sat) : dz_wf) * m;
    dz_wf = ((V(z_sc_non_sat) <= 0) ? 0 : dz_wf) * m;
    dz_wf = ((dz_wf <= 0) ? 0 : dz_wf) * m;
    V(z_sc_non_sat) = V(z_sc) - V(z_wf) * m;
    V(z_sc_non_sat) = ((V(z_sc_non_sat) < 0) ? 0 : V(z_sc_non_sat)) * m;
    V(z_sc_non_sat) = ((substeps == 1) ? V(z_sc_non_sat) + dz_wf / 1.39 : V(z_sc_non_sat)) * m;

    double y, ym1, inf_sc_pot_di = 0, inf_sc_di = 0, z0_di = 0, inf_sc_event_csum, t = 0, inf_sc = 0;
    y = p->y_sc * m;
    ym1 = p->y_sc * m;
    inf_sc_event_csum = V(inf_sc_event_csum) * m;
    for (int i = 0; i < substeps; i++) { /* loop_body :1198-1278 */
        z0_di = (V(z0) / substeps) * m;
        t += (dt / substeps) * m;
        y = (sqrt((V(ks) * V(wfs) * t * 2) / V(theta_d))) * m;
        inf_sc_pot_di = ((V(z_sc_non_sat) * V(theta_d) * l_sc) * (y - ym1) * 1e-06) * m;
        inf_sc_pot_di = ((inf_sc_pot_di <= 0) ? 0 : inf_sc_pot_di) * m;
        inf_sc_di = ((inf_sc_pot_di > z0_di) ? z0_di : inf_sc_pot_di) * m;
        inf_sc_di = ((V(z_sc_non_sat) <= 0) ? 0 : inf_sc_di) * m;
        inf_sc_di += inf_sc_di * m; /* :1248-1252 doubles it */
        inf_sc_event_csum += inf_sc_di * m;
        y = (inf_sc_event_csum / l_sc / 2) * m;
        t = (((ym1 * ym1) * V(theta_d)) / (V(ks) * V(wfs) * 2)) * m;
        ym1 = y * m;
    }
    V(y_sc) = y * m;
    V(inf_sc) = inf_sc * m; /* the loop never updates inf_sc: always 0 */
    V(inf_sc_event_csum) += V(inf_sc) * m;
    V(z0) += -V(inf_sc) * m;
    V(z0) = ((V(z0) < 0) ? 0 : V(z0)) * m;
}

static void calc_inf_rz(double *f) /* infiltration.py:1321-1417 */
{
    const double m = V(maskCatch);
    V(inf_mat_rz) = V(inf_mat) * m;
    V(inf_sc_rz) = V(inf_sc) * m;
    V(inf_rz) = (V(inf_mat_rz) + V(inf_mp_rz) + V(inf_sc_rz)) * m;
    V(S_fp_rz) += V(inf_rz) * m;
    int mk = V(S_fp_rz) > V(S_ufc_rz);
    V(S_lp_rz) += (mk ? (V(S_fp_rz) - V(S_ufc_rz)) : 0) * m;
    V(S_fp_rz) = (mk ? V(S_ufc_rz) : V(S_fp_rz)) * m;
    mk = V(S_lp_rz) > V(S_ac_rz);
    V(inf_mp_rz) += (mk ? -(V(S_lp_rz) - V(S_ac_rz)) : 0) * m;
    V(inf_mp_rz) = ((V(inf_mp_rz) < 0) ? 0 : V(inf_mp_rz)) * m;
    V(z0) += (mk ? V(S_lp_rz) - V(S_ac_rz) : 0) * m;
    V(S_lp_rz) = (mk ? V(S_ac_rz) : V(S_lp_rz)) * m;
    V(inf_mp) = 0;
    V(inf_mp) = V(inf_mp_rz) + V(inf_mp_ss) * m; /* :1399 */
    V(inf_rz) = (V(inf_mat_rz) + V(inf_mp_rz) + V(inf_sc_rz)) * m;
}

static void calc_inf_hof_sof_sur(double *f) /* infiltration.py:1519-1532, :1420-1476, :1479-1516 */
{
    const double m = V(maskCatch);
    V(inf) = (V(inf_rz) + V(inf_ss)) * m;
    V(q_hof) = 0;
    V(q_hof) = V(z0) * m;
    V(q_hof) = ((V(q_hof) < 0) ? 0 : V(q_hof)) * m;
    V(q_sof) = 0;
    int mask2 = ((V(S_lp_rz) + V(S_fp_rz)) > (V(S_ac_rz) + V(S_ufc_rz))) && ((V(S_lp_ss) + V(S_fp_ss)) >= (V(S_ac_ss) + V(S_ufc_ss)));
    V(q_sof) = (mask2 ? (V(S_lp_rz) + V(S_fp_rz)) - (V(S_ac_rz) + V(S_ufc_rz)) : V(q_sof)) * m;
    int mk = V(q_sof) > 0;
    V(S_fp_rz) = (mk ? V(S_ufc_rz) : V(S_fp_rz)) * m;
    V(S_lp_rz) = (mk ? V(S_ac_rz) : V(S_lp_rz)) * m;
    /* calc_surface_runoff (enable_routing_1D is off) */
    V(z0) += -V(q_hof) * m;
    V(z0) = ((V(z0) < 0) ? 0 : V(z0)) * m;
    V(q_sur) = 0;
    V(q_sur) += (V(q_hof) + V(q_sof)) * m;
    V(q_sur) += (V(maskRiverLake) != 0) ? V(prec) : 0;
}

static void calculate_infiltration(double *f, const step_ctx *c, cell_m1 *p) /* infiltration.py:2147-2193 */
{
    if (c->cond1) {
        calc_depth_shrinkage_cracks(f);
        set_event_vars(f, p);
    }
    /* cond2/cond3: `scalar & grid.any()` gates whose kernels re-apply the per-cell mask; a cell
       whose own mask is false is left unchanged (x * maskCatch re-multiplication is idempotent) */
    if (c->cond2_scalar) set_event_vars_start_rainfall_pause(f, p);
    if (c->cond3_scalar) set_event_vars_end_rainfall_pause(f, p);
    if (c->cond5) V(t_event_csum) += c->dt;
    calc_green_ampt_params(f, c);
    calc_inf_mat(f, c, p);
    calc_inf_mp(f, c, p);
    calc_inf_sc(f, c, p);
    calc_inf_rz(f);
    calc_inf_hof_sof_sur(f);
    if (c->cond4) reset_event_vars(f, p);
}

/* ================================================================================== */
/* subsurface_runoff.py                                                               */
/* ================================================================================== */
static void calc_rise_of_saturation_water_table(double *f) /* subsurface_runoff.py:693-765 */
{
    const double m = V(maskCatch);
    double lmpv_ss = V(lmpv) - V(z_root) * m; /* :705 */
    lmpv_ss = ((V(lmpv) < V(z_root)) ? 0 : lmpv_ss) * m;
    double z_sat_top = ((V(S_lp_ss) < V(theta_ac)) ? V(S_lp_ss) / V(theta_ac) : V(S_lp_rz) + V(S_lp_ss) / V(theta_ac)) * m;
    double z_nomp = (V(z_soil) - V(z_root)) - lmpv_ss - V(z_sat) * m; /* :733 */
    z_nomp = (z_nomp < 0) ? 0 : z_nomp;
    double v;
    if (z_sat_top > z_nomp) {
        if ((V(S_fp_ss) >= V(S_ufc_ss)) && (((V(S_lp_ss) + 1e-6) / V(theta_ac)) < (V(z_soil) - V(z_root))))
            v = V(S_lp_ss) / V(theta_ac);
        else if ((V(S_fp_rz) >= V(S_ufc_rz)) && (V(S_lp_ss) + 1e-6 >= V(S_ac_ss)))
            v = V(S_lp_rz) / V(theta_ac) + V(S_lp_ss) / V(theta_ac);
        else
            v = V(S_lp_ss) / V(theta_ac);
    } else {
        v = V(S_lp_ss) / V(theta_ac);
    }
    V(z_sat) = v * m;
}

static void calc_S_zsat(double *f) /* subsurface_runoff.py:7-48 */
{
    const double m = V(maskCatch);
    V(S_zsat) = ((V(z_sat) <= V(z_soil)) ? V(z_sat) * V(theta_ac) : V(z_soil) * V(theta_ac)) * m;
    V(S_zsat_ss) = ((V(z_sat) <= V(z_soil) - V(z_root)) ? V(S_zsat) : (V(z_soil) - V(z_root)) * V(theta_ac)) * m;
    V(S_zsat_rz) = ((V(z_sat) > V(z_soil) - V(z_root)) ? (V(z_sat) - (V(z_soil) - V(z_root))) * V(theta_ac) : 0) * m;
}

static void calc_z_sat_layer(double *f) /* subsurface_runoff.py:52-245 */
{
    const double m = V(maskCatch);
    double *L[8] = {&V(z_sat_layer_1), &V(z_sat_layer_2), &V(z_sat_layer_3), &V(z_sat_layer_4),
                    &V(z_sat_layer_5), &V(z_sat_layer_6), &V(z_sat_layer_7), &V(z_sat_layer_8)};
    for (int k = 0; k < 8; k++) {
        double z;
        if (k == 0) z = V(z_sat) * m;
        else z = V(z_sat) - (200.0 * k) * m;
        if (k < 7) z = ((z > 200) ? 200 : z) * m;
        z = ((z <= 0) ? 0 : z) * m;
        *L[k] = z;
    }
}

static void calc_potential_percolation_rz(double *f, const step_ctx *c, const cell_m1 *p) /* subsurface_runoff.py:768-888 */
{
    const double m = V(maskCatch), dt = c->dt;
    int mask1 = (V(z_wf) < V(z_root)) && (V(z_sat) <= 0);
    int mask2 = (V(z_wf) >= V(z_root)) && (V(z_sat) <= 0);
    int mask3 = (V(z_sat) > 0) && (V(z_root) < V(z_soil) - V(z_sat));
    double perc_pot = (mask1 ? V(k_rz) * dt : 0) * m;
    perc_pot = (mask2 ? V(k_rz) * dt : perc_pot) * m;
    double z = (V(z_soil) - V(z_root)) - V(z_sat);
    if (mask3) {
        double pa = opow(z / (-V(ha) * 10.2), -V(n_salv));
        double pb = opow(-V(h_rz) / -V(ha), -V(n_salv));
        perc_pot = (((pa - pb) / (1 + pb + (V(n_salv) - 1) * pa)) * dt * V(ks) * (-1)) * m;
    } else {
        perc_pot = perc_pot * m;
    }
    perc_pot = ((perc_pot < 0) ? 0 : perc_pot) * m;
    int gate = p->z_root < V(z_soil) - V(z_sat);
    int mask4 = (perc_pot > 0) && (V(S_lp_rz) + V(S_fp_rz) >= perc_pot) && gate;
    int mask5 = (perc_pot > 0) && (V(S_lp_rz) + V(S_fp_rz) < perc_pot) && gate;
    V(q_pot_rz) = 0;
    V(q_pot_rz) = (mask4 ? perc_pot : V(q_pot_rz)) * m;
    V(q_pot_rz) = (mask5 ? V(S_fp_rz) + V(S_lp_rz) : V(q_pot_rz)) * m;
    double room = (V(S_ac_ss) + V(S_ufc_ss)) - (V(S_lp_ss) + V(S_fp_ss));
    int mask6 = (V(q_pot_rz) > 0) && (room > 0) && (V(q_pot_rz) > room) && gate;
    V(q_pot_rz) = (mask6 ? room : V(q_pot_rz)) * m;
    int mask7 = (V(S_lp_ss) >= V(S_ac_ss) - 1e-6) && (V(S_fp_ss) >= V(S_ufc_ss) - 1e-6);
    V(q_pot_rz) = (mask7 ? 0 : V(q_pot_rz)) * m;
    int mask8 = p->z_root >= V(z_soil) - V(z_sat);
    V(q_pot_rz) = (mask8 ? 0 : V(q_pot_rz)) * m;
}

static void calc_percolation_rz(double *f) /* subsurface_runoff.py:891-968 */
{
    const double m = V(maskCatch);
    int mask1 = (V(S_lp_rz) < V(q_pot_rz)) && (V(z_sat) < V(z_soil) - V(z_root));
    int mask2 = (V(S_lp_rz) >= V(q_pot_rz)) && (V(z_sat) < V(z_soil) - V(z_root));
    int mask3 = V(z_sat) >= V(z_soil) - V(z_root);
    V(q_rz) = V(q_pot_rz) * m;
    V(q_rz) = (mask3 ? 0 : V(q_rz)) * m;
    V(S_fp_rz) += (mask1 ? -(V(q_rz) - V(S_lp_rz)) : 0) * m;
    V(S_lp_rz) = (mask1 ? 0 : V(S_lp_rz)) * m;
    V(S_lp_rz) += (mask2 ? -V(q_rz) : 0) * m;
    V(S_fp_ss) += V(q_rz) * m;
    int mk = V(S_fp_ss) > V(S_ufc_ss);
    V(S_lp_ss) += (mk ? V(S_fp_ss) - V(S_ufc_ss) : 0) * m;
    V(S_fp_ss) = (mk ? V(S_ufc_ss) : V(S_fp_ss)) * m;
    mk = V(S_lp_ss) > V(S_ac_ss);
    V(q_rz) += (mk ? -(V(S_lp_ss) - V(S_ac_ss)) : 0) * m;
    V(S_lp_rz) += (mk ? V(S_lp_ss) - V(S_ac_ss) : 0) * m;
    V(S_lp_ss) = (mk ? V(S_ac_ss) : V(S_lp_ss)) * m;
}

static void calc_potential_percolation_ss(double *f, const step_ctx *c) /* subsurface_runoff.py:971-1099 */
{
    const double m = V(maskCatch), dt = c->dt;
    const double zgw = V(z_gw);
    double z = (zgw * 1000 - V(z_soil)) + ((V(z_soil) - V(z_root)) / 2) * m;
    double perc_pot;
    double pa = opow(z / (-V(ha) * 10.2), -V(n_salv));
    double pb = opow(-V(h_ss) / -V(ha), -V(n_salv));
    double kmin = fmin(fmin(V(kf) * dt, V(ks_ss) * dt), V(k_ss) * dt);
    /* first assignment (z_gw > 10) :984-1003 */
    if ((zgw > 10) && (zgw * 1000 > V(z_soil)) && (V(z_sat) > 0))
        perc_pot = kmin * m;
    else
        perc_pot = fmin(V(kf) * dt, (pa / ((1 + (V(n_salv) - 1)) * pa)) * dt * V(ks_ss)) * m;
    /* second assignment (z_gw <= 10) overwrites everything :1005-1027 */
    if ((zgw <= 10) && (zgw * 1000 > V(z_soil)) && (V(z_sat) > 0))
        perc_pot = kmin * m;
    else
        perc_pot = fmin(V(kf) * dt, (pa - pb) / (1 + pb + (V(n_salv) - 1) * pa) * dt * V(ks_ss) * (-1)) * m;
    int mask1 = (perc_pot > 0) && (V(z_soil) < zgw * 1000) && (perc_pot <= V(S_fp_ss) + V(S_lp_ss));
    int mask2 = (perc_pot > 0) && (V(z_soil) < zgw * 1000) && (perc_pot > V(S_fp_ss) + V(S_lp_ss));
    V(q_pot_ss) = 0;
    V(q_pot_ss) = (mask1 ? perc_pot : V(q_pot_ss)) * m;
    V(q_pot_ss) = (mask2 ? V(S_fp_ss) + V(S_lp_ss) : V(q_pot_ss)) * m;
    double cpr_pot = (((pa - pb) / (1 + pb + (V(n_salv) - 1) * pa)) * dt * V(ks_ss)) * m;
    cpr_pot = (((perc_pot > 0) && (V(z_soil) < zgw * 1000)) ? 0 : cpr_pot) * m;
    int mask3 = (zgw * 1000 - V(z_soil) > 10000);
    cpr_pot = (mask3 ? 0 : cpr_pot) * m;
    V(q_pot_ss) = ((cpr_pot > 0) ? 0 : V(q_pot_ss)) * m;
}

static void calc_percolation_ss(double *f) /* subsurface_runoff.py:1102-1154 */
{
    const double m = V(maskCatch);
    V(q_ss) = V(q_pot_ss) * m;
    V(z_sat) += ((V(z_sat) > 0) ? -V(q_ss) / V(theta_ac) : 0) * m;
    V(z_sat) = ((V(z_sat) < 0) ? 0 : V(z_sat)) * m;
    V(S_zsat_ss) = V(z_sat) * V(theta_ac) * m;
    int mask1 = V(S_lp_ss) < V(q_pot_ss);
    int mask2 = V(S_lp_ss) >= V(q_pot_ss);
    V(S_fp_ss) += (mask1 ? -(V(q_ss) - V(S_lp_ss)) : 0) * m;
    V(S_lp_ss) = (mask1 ? 0 : V(S_lp_ss)) * m;
    V(S_lp_ss) += (mask2 ? -V(q_ss) : 0) * m;
}

/* ---- oneD lateral flow (cell-local sink) ---- */
static void calc_potential_lateral_subsurface_runoff(double *f, const step_ctx *c) /* subsurface_runoff.py:249-362 */
{
    const double m = V(maskCatch), dt = c->dt, dx = c->st->dx, r_mp = c->st->r_mp, pi = c->st->pi;
    V(q_sub_mat_pot) = ((V(ks) * V(slope) * V(z_sat) * dx * 1000 * dt) * 1e-6 * (1 / (dx * (V(z_soil) / 1000)))) * m;
    V(q_sub_mat_pot) = ((V(z_sat) <= 0) ? 0 : V(q_sub_mat_pot)) * m;
    const double *zl[8] = {&V(z_sat_layer_1), &V(z_sat_layer_2), &V(z_sat_layer_3), &V(z_sat_layer_4),
                           &V(z_sat_layer_5), &V(z_sat_layer_6), &V(z_sat_layer_7), &V(z_sat_layer_8)};
    const double *vl[8] = {&V(v_mp_layer_1), &V(v_mp_layer_2), &V(v_mp_layer_3), &V(v_mp_layer_4),
                           &V(v_mp_layer_5), &V(v_mp_layer_6), &V(v_mp_layer_7), &V(v_mp_layer_8)};
    double sum = 0;
    for (int k = 0; k < 8; k++) {
        double term = *zl[k] * *vl[k] * dt * dx * 1000 * V(dmph) * 1e-6 * (r_mp * r_mp) * pi * 1e-6;
        sum = (k == 0) ? term : sum + term;
    }
    V(q_sub_mp_pot) = (sum * (1 / (dx * (V(z_soil) / 1000)))) * m;
    V(q_sub_mp_pot) = ((V(q_sub_mp_pot) < 0) ? 0 : V(q_sub_mp_pot)) * m;
    V(q_sub_mp_pot) = ((V(z_sat) <= 0) ? 0 : V(q_sub_mp_pot)) * m;
    V(q_sub_pot) = (V(q_sub_mp_pot) + V(q_sub_mat_pot)) * m;
    V(q_sub_mat_share) = (V(q_sub_mat_pot) / V(q_sub_pot)) * m;
    V(q_sub_mat_share) = ((V(q_sub_pot) == 0) ? 0 : V(q_sub_mat_share)) * m;
    V(q_sub_mp_share) = (V(q_sub_mp_pot) / V(q_sub_pot)) * m;
    V(q_sub_mp_share) = ((V(q_sub_pot) == 0) ? 0 : V(q_sub_mp_share)) * m;
    int mask1 = V(q_sub_pot) > V(S_lp_rz) + V(S_lp_ss);
    V(q_sub_pot) = (mask1 ? V(S_lp_rz) + V(S_lp_ss) : V(q_sub_pot)) * m;
    V(q_sub_mat_pot) = V(q_sub_pot) * V(q_sub_mat_share) * m;
    V(q_sub_mp_pot) = V(q_sub_pot) * V(q_sub_mp_share) * m;
}

static void calc_lateral_subsurface_runoff_rz(double *f) /* subsurface_runoff.py:365-457 */
{
    const double m = V(maskCatch);
    double rz_share = ((V(z_sat) > 0) ? ((V(z_sat) - (V(z_soil) - V(z_root))) / V(z_sat)) : 0) * m;
    int mask1 = (V(z_sat) <= V(z_soil) - V(z_root)) || (V(S_lp_rz) <= 0);
    rz_share = (mask1 ? 0 : rz_share) * m;
    int mask2 = isnan(rz_share);
    rz_share = (mask2 ? 0 : rz_share) * m;
    V(S_zsat_rz) = ((V(z_sat) * rz_share) * V(theta_ac)) * m;
    V(q_sub_rz) = ((V(q_sub_pot) * rz_share < V(S_zsat_rz)) ? V(q_sub_pot) * rz_share : V(S_zsat_rz)) * m;
    V(q_sub_mat_rz) = V(q_sub_rz) * V(q_sub_mat_share) * m;
    V(q_sub_mp_rz) = V(q_sub_rz) * V(q_sub_mp_share) * m;
    V(q_sub_mp_pot_rz) = V(q_sub_mp_pot) * rz_share * m;
    V(z_sat) += -V(q_sub_rz) / V(theta_ac) * m;
    V(S_lp_rz) += -V(q_sub_rz) * m;
}

static void calc_potential_lateral_subsurface_runoff_ss(double *f) /* subsurface_runoff.py:461-517 */
{
    const double m = V(maskCatch);
    double ss_share = ((V(z_soil) - V(z_root)) / V(z_sat)) * m;
    int mask1 = (V(z_sat) <= V(z_soil) - V(z_root)) || (V(S_lp_rz) <= 0);
    int mask2 = V(z_sat) <= 0;
    int mask3 = isnan(ss_share);
    ss_share = (mask1 ? 1 : ss_share) * m;
    ss_share = (mask2 ? 0 : ss_share) * m;
    ss_share = (mask3 ? 0 : ss_share) * m;
    V(q_sub_mat_pot_ss) = V(q_sub_mat_pot) * ss_share * m;
    V(q_sub_mp_pot_ss) = V(q_sub_mp_pot) * ss_share * m;
    V(q_sub_pot_ss) = (V(q_sub_mat_pot_ss) + V(q_sub_mp_pot_ss)) * m;
}

static void calc_lateral_subsurface_runoff_ss(double *f) /* subsurface_runoff.py:521-662, :666-690 */
{
    const double m = V(maskCatch);
    V(q_ss) = 0;
    V(q_ss) = ((V(z_sat) <= 0) ? V(q_pot_ss) : V(q_ss)) * m;
    double tot = V(q_pot_ss) + V(q_sub_pot_ss);
    double fv = ((tot > 0) ? V(q_pot_ss) / tot : 0) * m;
    double fl = ((tot > 0) ? V(q_sub_pot_ss) / tot : 0) * m;
    double q_ss_sat = ((tot <= V(S_zsat_ss)) ? tot * fv : V(S_zsat_ss) * fv) * m;
    V(q_ss) = (V(z_sat) > 0) ? q_ss_sat : V(q_ss);
    V(q_sub_ss) = 0;
    V(q_sub_ss) = ((tot <= V(S_zsat_ss)) ? tot * fl : V(S_zsat_ss) * fl) * m;
    V(q_sub_mat_ss) = V(q_sub_ss) * V(q_sub_mat_share) * m;
    V(q_sub_mp_ss) = V(q_sub_ss) * V(q_sub_mp_share) * m;
    int mask1 = V(S_lp_ss) < V(q_ss);
    int mask2 = V(S_lp_ss) >= V(q_ss);
    V(S_fp_ss) += (mask1 ? -(V(q_ss) - V(S_lp_ss)) : 0) * m;
    V(S_lp_ss) = (mask1 ? 0 : V(S_lp_ss)) * m;
    V(S_lp_ss) += (mask2 ? -V(q_ss) : 0) * m;
    int mk = V(z_sat) > 0;
    V(S_lp_ss) += (mk ? -V(q_sub_ss) : 0) * m;
    V(z_sat) += -((V(q_sub_ss) + V(q_ss)) / V(theta_ac)) * m;
    V(z_sat) = ((V(z_sat) < 0) ? 0 : V(z_sat)) * m;
    V(S_zsat) = V(z_sat) * V(theta_ac) * m;
    /* calc_lateral_subsurface_runoff :666-690 */
    V(q_sub_mat) = (V(q_sub_mat_rz) + V(q_sub_mat_ss)) * m;
    V(q_sub_mp) = (V(q_sub_mp_rz) + V(q_sub_mp_ss)) * m;
    V(q_sub) = (V(q_sub_rz) + V(q_sub_ss)) * m;
}

/* ================================================================================== */
/* capillary_rise.py                                                                  */
/* ================================================================================== */
static void calc_cpr_rz(double *f, const step_ctx *c) /* capillary_rise.py:7-173 */
{
    const double m = V(maskCatch), dt = c->dt;
    double z = ((V(z_root) + (V(z_soil) - V(z_root)) / 2) - V(z_root) / 2) * m;
    double pa = opow(z / (-V(ha) * 10.2), -V(n_salv));
    double pb = opow(-V(h_rz) / -V(ha), -V(n_salv));
    V(cpr_rz) = (pa - pb) / (1 + pb + (V(n_salv) - 1) * pa) * dt * V(ks) * m;
    V(cpr_rz) = ((V(cpr_rz) < 0) ? 0 : V(cpr_rz)) * m;
    V(cpr_rz) = (isnan(V(cpr_rz)) ? 0 : V(cpr_rz)) * m;
    V(cpr_rz) = ((V(S_lp_rz) > 0) ? 0 : V(cpr_rz)) * m;
    V(cpr_rz) = ((V(h_rz) > V(h_ss)) ? 0 : V(cpr_rz)) * m;
    V(cpr_rz) = ((V(cpr_rz) > (V(S_fp_ss) + V(S_lp_ss))) ? V(S_fp_ss) + V(S_lp_ss) : V(cpr_rz)) * m;
    double room = V(S_ufc_rz) - V(S_fp_rz);
    V(cpr_rz) = (((V(cpr_rz) > room) && (room > 0)) ? room : V(cpr_rz)) * m;
    int gate = (V(z_wf) < V(z_root)) || (V(z_sat) < V(z_soil) - V(z_root));
    int mask1 = (V(cpr_rz) > 0) && (V(S_lp_ss) <= 0) && gate;
    int mask2 = (V(cpr_rz) > 0) && (V(S_lp_ss) > 0) && (V(cpr_rz) <= V(S_lp_ss)) && gate;
    int mask3 = (V(cpr_rz) > 0) && (V(S_lp_ss) > 0) && (V(cpr_rz) > V(S_lp_ss)) && gate;
    V(S_fp_rz) += (mask1 ? V(cpr_rz) : 0) * m;
    V(S_fp_ss) += (mask1 ? -V(cpr_rz) : 0) * m;
    V(S_fp_rz) += (mask2 ? V(cpr_rz) : 0) * m;
    V(S_lp_ss) += (mask2 ? -V(cpr_rz) : 0) * m;
    V(S_fp_rz) += (mask3 ? V(cpr_rz) : 0) * m;
    V(S_fp_ss) += (mask3 ? -(V(cpr_rz) - V(S_lp_ss)) : 0) * m;
    V(S_lp_ss) = (mask3 ? 0 : V(S_lp_ss)) * m;
    int mask4 = V(S_fp_rz) > V(S_ufc_rz);
    V(S_lp_rz) += (mask4 ? (V(S_fp_rz) - V(S_ufc_rz)) : 0) * m;
    V(S_fp_rz) = (mask4 ? V(S_ufc_rz) : V(S_fp_rz)) * m;
}

/* groundwater boundary (settings.enable_groundwater_boundary): capillary rise from the water table into the subsoil */
static void calc_cpr_ss(double *f, const step_ctx *c) /* capillary_rise.py:176-324 (no film flow) */
{
    const double m = V(maskCatch), dt = c->dt;
    const int mask1 = (V(z_sat) > 0) && (V(z_gw) * 1000 < V(z_soil));
    const int mask2 = V(z_gw) * 1000 < V(z_soil);
    const double z = (V(z_gw) * 1000 - V(z_soil)) + ((V(z_soil) - V(z_root)) / 2) * m;
    const double pa = opow(z / (-V(ha) * 10.2), -V(n_salv));
    const double pb = opow(-V(h_ss) / -V(ha), -V(n_salv));
    double cpr = (pa - pb) / (1 + pb + (V(n_salv) - 1) * pa) * dt * V(ks_ss) * m;
    cpr = ((cpr < 0) ? 0 : cpr) * m;
    cpr = (isnan(cpr) ? 0 : cpr) * m;
    cpr = ((cpr < 0) ? 0 : cpr) * m;
    cpr = ((cpr > (V(S_ufc_ss) - V(S_fp_ss))) ? V(S_ufc_ss) - V(S_fp_ss) : cpr) * m; /* room in the fine pores */
    cpr = (mask1 ? 0 : cpr) * m;                                                      /* saturated subsoil */
    cpr = ((cpr > V(S_lp_ss)) ? cpr - V(S_lp_ss) : cpr) * m;                           /* water in the large pores */
    cpr = ((V(z_gw) * 1000 - V(z_soil) > 10000) ? 0 : cpr) * m;                        /* water table deeper than 10 m */
    /* the water table rises into the subsoil */
    const double zgw_soil = (((V(z_gw) * 1000 < V(z_soil)) && (V(z_gw_m1) * 1000 >= V(z_soil))) ? V(z_soil) : V(z_gw_m1) * 1000) * m;
    double gw_rise = -(V(z_gw) * 1000 - zgw_soil) * V(theta_ac) * m;
    gw_rise = ((gw_rise < 0) ? 0 : gw_rise) * m;
    gw_rise = (mask2 ? gw_rise : 0) * m;
    cpr += gw_rise * m;
    V(cpr_ss) = cpr;
    V(S_fp_ss) += V(cpr_ss) * m;
    const int mask3 = V(S_fp_ss) > V(S_ufc_ss);
    V(S_lp_ss) += (mask3 ? (V(S_fp_ss) - V(S_ufc_ss)) : 0) * m;
    V(S_fp_ss) = (mask3 ? V(S_ufc_ss) : V(S_fp_ss)) * m;
}

/* groundwater boundary: recharge through the vadose zone below the soil, a linear reservoir that spills when full */
static void calc_q_re(double *f, const step_ctx *c) /* groundwater_flow.py:99-160 */
{
    const double m = V(maskCatch);
    const int mask1 = V(z_gw) * 1000 > V(z_soil);
    V(S_vad_tot) = (mask1 ? (V(z_gw) * 1000 - V(z_soil)) * V(n0) : 0) * m;
    V(S_vad) += V(q_ss) * m;
    const double k = (V(kf) / c->st->kf_max) * m;
    V(q_re) = k * V(S_vad) * m;
    V(S_vad) += -V(q_re) * m;
    const int mask2 = V(S_vad) > V(S_vad_tot);
    const double spill = (mask2 ? (V(S_vad) - V(S_vad_tot)) : 0) * m;
    V(q_re) += spill * m;
    V(S_vad) += -spill * m;
}

/* ================================================================================== */
/* storages: surface.py:8-28, root_zone.py:36-163, subsoil.py:6-130, soil.py:9-140,     */
/* numerics.py:125-152, :217-346                                                      */
/* ================================================================================== */
static void calc_storages(double *f, const step_ctx *c, const cell_m1 *p)
{
    const double m = V(maskCatch);
    V(S_sur) = (V(S_int_top) + V(S_int_ground) + V(S_dep) + V(S_snow) + V(z0)) * m;
    /* root zone */
    V(S_rz) = (V(S_pwp_rz) + V(S_fp_rz) + V(S_lp_rz)) * m;
    V(dS_rz) = (V(S_rz) - p->S_rz) * m;
    V(theta_rz) = ((V(S_fp_rz) + V(S_lp_rz)) / V(z_root) + V(theta_pwp)) * m;
    if (c->irrigation_month) { /* root_zone.py:7-33 (theta_irr is never set by the SVAT/oneD setups: 0) */
        double d = 0.0 - V(theta_rz);
        d = (d <= 0) ? 0 : d;
        V(irr_demand) = d * V(z_root);
    } else {
        V(irr_demand) = 0;
    }
    V(k_rz) = (V(ks) / (1 + opow(V(theta_rz) / V(theta_sat), -V(m_bc)))) * m;
    V(h_rz) = (V(ha) / opow(V(theta_rz) / V(theta_sat), 1 / V(lambda_bc))) * m;
    /* subsoil */
    V(S_ss) = (V(S_pwp_ss) + V(S_fp_ss) + V(S_lp_ss)) * m;
    V(dS_ss) = (V(S_ss) - p->S_ss) * m;
    V(theta_ss) = ((V(S_fp_ss) + V(S_lp_ss)) / (V(z_soil) - V(z_root)) + V(theta_pwp)) * m;
    V(ks_ss) = V(ks); /* subsoil.py:6-20 (no soil compaction) */
    V(k_ss) = (V(ks) / (1 + opow(V(theta_ss) / V(theta_sat), -V(m_bc)))) * m;
    V(h_ss) = (V(ha) / opow(V(theta_ss) / V(theta_sat), 1 / V(lambda_bc))) * m;
    /* soil */
    V(S_fp_s) = (V(S_fp_rz) + V(S_fp_ss)) * m;
    V(S_lp_s) = (V(S_lp_rz) + V(S_lp_ss)) * m;
    V(S_s) = (V(S_pwp_s) + V(S_fp_s) + V(S_lp_s)) * m;
    V(dS_s) = (V(S_s) - p->S_s) * m;
    V(theta) = ((V(S_fp_s) + V(S_lp_s)) / V(z_soil) + V(theta_pwp)) * m;
    V(k) = (V(ks) / (1 + opow(V(theta) / V(theta_sat), -V(m_bc)))) * m;
    V(h) = (V(ha) / opow(V(theta) / V(theta_sat), 1 / V(lambda_bc))) * m;
    /* numerics.calc_storage_kernel */
    V(S) = V(S_sur) + V(S_s) * m;       /* numerics.py:134 */
    V(dS) = V(S) - p->S * m;            /* numerics.py:147 */
}

static int calc_num_error_and_check(double *f, const step_ctx *c, const cell_m1 *p)
{
    const double atol = c->st->atol, rtol = c->st->rtol;
    double lhs = V(S) - p->S, rhs;
    if (c->st->enable_groundwater_boundary) {
        /* the balance error (numerics.py:271-290 with lateral flow, :396-420 without) and the closure the sanity check tests
         * (:822-841, :902-918) are different expressions when lateral flow is on: q_re there, q_ss here */
        if (c->st->enable_lateral_flow) {
            V(dS_num_error) = fabs(lhs - (V(prec) - V(q_sur) - V(aet) - V(q_re) - V(q_sub) + V(cpr_ss)));
            rhs = V(prec) - V(q_sur) - V(aet) - V(q_ss) - V(q_sub) + V(cpr_ss);
        } else {
            rhs = V(prec) - V(q_sur) - V(aet) - V(q_ss) + V(cpr_ss);
            V(dS_num_error) = fabs(lhs - rhs);
        }
    } else if (c->st->enable_lateral_flow) { /* numerics.py:224-246 */
        rhs = V(prec) - V(q_sur) - V(aet) - V(q_ss) - V(q_sub);
        V(dS_num_error) = fabs(lhs - rhs);
    } else { /* numerics.py:303-346 */
        rhs = V(prec) - V(q_sur) - V(aet) - V(q_ss);
        V(dS_num_error) = fabs(lhs - rhs);
        V(dS_rz_num_error) = fabs((V(S_rz) - p->S_rz) -
                                  (V(inf_mat_rz) + V(inf_mp_rz) + V(inf_sc_rz) + V(cpr_rz) - V(transp) - V(evap_soil) - V(q_rz)));
        V(dS_ss_num_error) = fabs((V(S_ss) - p->S_ss) - (V(inf_mp_ss) + V(q_rz) - V(q_ss) - V(cpr_rz)));
    }
    /* numerics.sanity_check :742-800 / :978-1016 */
    int ok = 1;
    if (V(maskCatch) != 0) {
        int close = fabs(lhs - rhs) <= (atol + rtol * fabs(rhs)); /* npx.isclose */
        if (!(isfinite(lhs) && isfinite(rhs))) close = (lhs == rhs);
        ok = ok && close;
    }
#define NZ(x) (isnan(x) ? 0.0 : (x))
    ok = ok && (NZ(V(S_fp_rz)) > -atol) && (NZ(V(S_lp_rz)) > -atol) && (NZ(V(S_fp_ss)) > -atol) && (NZ(V(S_lp_ss)) > -atol);
    ok = ok && (NZ(V(S_fp_rz)) - atol <= NZ(V(S_ufc_rz))) && (NZ(V(S_lp_rz)) - atol <= NZ(V(S_ac_rz))) &&
         (NZ(V(S_fp_ss)) - atol <= NZ(V(S_ufc_ss))) && (NZ(V(S_lp_ss)) - atol <= NZ(V(S_ac_ss)));
#undef NZ
    return ok;
}

/* ================================================================================== */
/* adaptive time stepping (grid-wide part)  adaptive_time_stepping.py:34-381          */
/* ================================================================================== */
typedef struct {
    int daily, hourly, tenmin;          /* which aggregate prec/ta[tau] take (first block, :130-189) */
    int c6, c7, c8, c9, c10, c11;       /* second block (:207-360) */
} ats_plan;

static const double *row_of(const double *day, int64_t i, int uniform) { return uniform ? day : day + (size_t)i * NSLOT; }

/* Multi-process runs (cells partitioned over ranks, roger_b200/parallel.py): the grid-wide any()/all() below must see
 * every rank's cells.  Same contract as roger_svat_set_flag_reducer (include/roger_b200.h): the callback ORs the flag
 * words in place across all ranks.  all(x) travels as !any(!x). */
static roger_flag_reduce_fn g_reduce_fn = NULL;
static void *g_reduce_user = NULL;
void svat_oracle_set_flag_reducer(roger_flag_reduce_fn fn, void *user) { g_reduce_fn = fn; g_reduce_user = user; }
static uint32_t reduce_or(uint32_t w)
{
    if (g_reduce_fn) g_reduce_fn(&w, 1, g_reduce_user);
    return w;
}

void svat_oracle_step(int64_t C, double *planes, const double *prec_day, const double *ta_day,
                      const double *pet_day, roger_svat_scalars *sc, const roger_svat_settings *st)
{
    const int uni = st->forcing_uniform;
    const double ta_fm = st->ta_fm, hpi = st->hpi;
#define PL(name) (planes + (size_t)RF_##name * (size_t)C)
    double *prec = PL(prec), *ta = PL(ta), *pet = PL(pet), *swe = PL(swe), *swe_top = PL(swe_top);
    double *prec_m1 = (double *)malloc(sizeof(double) * (size_t)C);
    memcpy(prec_m1, prec, sizeof(double) * (size_t)C);

    /* ---- grid-wide conditions :38-81 ---- */
    int all_prec_le0 = 1, all_ta_gt = 1, any_snowfall = 0, all_dry_cold = 1, any_hpi = 0, any_prec = 0, any_ta_gt = 0,
        all_le_hpi = 1, all_swe_le0 = 1, all_swetop_le0 = 1, any_swe = 0, any_swetop = 0;
    int64_t nrows = uni ? 1 : C;
    for (int64_t i = 0; i < nrows; i++) {
        const double *pr = row_of(prec_day, i, uni), *tr = row_of(ta_day, i, uni);
        for (int k = 0; k < NSLOT; k++) {
            double p = pr[k], t = tr[k];
            all_prec_le0 &= (p <= 0);
            all_ta_gt &= (t > ta_fm);
            any_snowfall |= ((p > 0) && (t <= ta_fm));
            all_dry_cold &= ((p <= 0) && (t <= ta_fm));
            any_hpi |= (p > hpi);
            any_prec |= (p > 0);
            any_ta_gt |= (t > ta_fm);
            all_le_hpi &= (p <= hpi);
        }
    }
    for (int64_t i = 0; i < C; i++) {
        all_swe_le0 &= (swe[i] <= 0);
        all_swetop_le0 &= (swe_top[i] <= 0);
        any_swe |= (swe[i] > 0);
        any_swetop |= (swe_top[i] > 0);
    }
    {
        uint32_t w = (uint32_t)(!all_prec_le0) | (uint32_t)(!all_ta_gt) << 1 | (uint32_t)any_snowfall << 2 | (uint32_t)(!all_dry_cold) << 3 |
                     (uint32_t)any_hpi << 4 | (uint32_t)any_prec << 5 | (uint32_t)any_ta_gt << 6 | (uint32_t)(!all_le_hpi) << 7 |
                     (uint32_t)(!all_swe_le0) << 8 | (uint32_t)(!all_swetop_le0) << 9 | (uint32_t)any_swe << 10 | (uint32_t)any_swetop << 11;
        w = reduce_or(w);
        all_prec_le0 = !(w & 1); all_ta_gt = !(w >> 1 & 1); any_snowfall = w >> 2 & 1; all_dry_cold = !(w >> 3 & 1);
        any_hpi = w >> 4 & 1; any_prec = w >> 5 & 1; any_ta_gt = w >> 6 & 1; all_le_hpi = !(w >> 7 & 1);
        all_swe_le0 = !(w >> 8 & 1); all_swetop_le0 = !(w >> 9 & 1); any_swe = w >> 10 & 1; any_swetop = w >> 11 & 1;
    }
    int cond0 = all_prec_le0 && all_swe_le0 && all_swetop_le0 && all_ta_gt;
    int cond00 = any_snowfall || all_dry_cold;
    int cond1 = any_hpi && any_prec && any_ta_gt;
    int cond2 = all_le_hpi && any_prec && any_ta_gt;
    int snow_warm = (any_swe || any_swetop) && any_ta_gt;
    int cond3 = any_hpi && any_prec && snow_warm;
    int cond4 = all_le_hpi && any_prec && snow_warm;
    int cond5 = all_prec_le0 && snow_warm;
    int cond_time = (sc->time % (24 * 60 * 60)) == 0;

    ats_plan pl;
    pl.daily = cond0 || cond00;
    pl.hourly = (cond2 || cond4 || cond5) && !cond1 && !cond3;
    pl.tenmin = (cond1 || cond3) && !cond2 && !cond4 && !cond5;
    const int64_t itt_day = sc->itt_day; /* all aggregates use the itt_day at entry (:126-128) */

    /* first block: prec / ta [tau] :130-189 */
    for (int64_t i = 0; i < C; i++) {
        const double *pr = row_of(prec_day, i, uni), *tr = row_of(ta_day, i, uni);
        if (pl.daily) { prec[i] = np_sum144(pr); ta[i] = day_nanmean(tr); }
        if (pl.hourly) { prec[i] = window_sum(pr, itt_day); ta[i] = window_nanmean(tr, itt_day); }
        if (pl.tenmin) { prec[i] = pr[itt_day]; ta[i] = tr[itt_day]; }
    }
    sc->dt_secs = cond_time ? 24 * 60 * 60 : 60 * 60; /* :143-144 */
    if (pl.hourly) sc->dt_secs = 60 * 60;
    if (pl.tenmin) sc->dt_secs = 10 * 60;

    /* event start / end :191-203 (swe[taum1] == swe[tau] here: after_timestep copied it) */
    int ev1a = 0, ev1b = 0, all_p_le0 = 1, all_snowing = 1, any_swe_m1 = 0, all_swe_le0_now = 1;
    for (int64_t i = 0; i < C; i++) {
        ev1a |= ((prec[i] > 0) && (ta[i] > ta_fm));
        ev1b |= (((swe[i] > 0) || (swe_top[i] > 0)) && (ta[i] > ta_fm));
        all_p_le0 &= (prec[i] <= 0);
        all_snowing &= ((prec[i] > 0) && (ta[i] <= ta_fm));
        any_swe_m1 |= (swe[i] > 0);
        all_swe_le0_now &= (swe[i] <= 0);
    }
    {
        uint32_t w = (uint32_t)ev1a | (uint32_t)ev1b << 1 | (uint32_t)(!all_p_le0) << 2 | (uint32_t)(!all_snowing) << 3 |
                     (uint32_t)any_swe_m1 << 4 | (uint32_t)(!all_swe_le0_now) << 5;
        w = reduce_or(w);
        ev1a = w & 1; ev1b = w >> 1 & 1; all_p_le0 = !(w >> 2 & 1); all_snowing = !(w >> 3 & 1);
        any_swe_m1 = w >> 4 & 1; all_swe_le0_now = !(w >> 5 & 1);
    }
    int cond_event1 = ev1a || ev1b;
    int cond_event2 = all_p_le0 || all_snowing || (any_swe_m1 && all_swe_le0_now);
    if (cond_event1) sc->time_event0 = 0;
    if (cond_event2) sc->time_event0 = sc->time_event0 + sc->dt_secs;

    /* second block :207-360 */
    const int64_t end_event = (int64_t)st->end_event;
    pl.c6 = (sc->time_event0 <= end_event) && (sc->dt_secs == 10 * 60);
    pl.c7 = (sc->time_event0 <= end_event) && (sc->dt_secs == 60 * 60);
    pl.c8 = (sc->time_event0 <= end_event) && (sc->dt_secs == 24 * 60 * 60);
    pl.c9 = (sc->time_event0 > end_event) && (sc->time % (60 * 60) != 0) && (sc->dt_secs == 10 * 60);
    pl.c10 = (sc->time_event0 > end_event) && (sc->time % (60 * 60) == 0) && ((sc->dt_secs == 10 * 60) || (sc->dt_secs == 60 * 60));
    pl.c11 = (sc->time_event0 > end_event) && (sc->time % (24 * 60 * 60) == 0) && (sc->dt_secs == 24 * 60 * 60);
    for (int64_t i = 0; i < C; i++) {
        const double *tr = row_of(ta_day, i, uni), *er = row_of(pet_day, i, uni);
        if (pl.c6) { pet[i] = er[itt_day]; ta[i] = tr[itt_day]; }
        if (pl.c7) { pet[i] = window_sum(er, itt_day); ta[i] = window_nanmean(tr, itt_day); }
        if (pl.c8) { pet[i] = np_sum144(er); ta[i] = day_nanmean(tr); }
        if (pl.c9) { pet[i] = er[itt_day]; ta[i] = tr[itt_day]; }
        if (pl.c10) { pet[i] = window_sum(er, itt_day); ta[i] = window_nanmean(tr, itt_day); }
        if (pl.c11) { pet[i] = np_sum144(er); ta[i] = day_nanmean(tr); }
    }
    if (pl.c6) { sc->event_id[1] = sc->event_id_counter; sc->dt = 1.0 / 6; sc->itt_day += 1; }
    if (pl.c7) { sc->event_id[1] = sc->event_id_counter; sc->dt = 1; sc->itt_day += 6; }
    if (pl.c8) { sc->dt = 24; sc->itt_day = 0; }
    if (pl.c9) { sc->event_id[1] = 0; sc->dt = 1.0 / 6; sc->dt_secs = 10 * 60; sc->itt_day += 1; }
    if (pl.c10) { sc->event_id[1] = 0; sc->dt = 1; sc->dt_secs = 60 * 60; sc->itt_day += 6; }
    if (pl.c11) { sc->event_id[1] = 0; sc->dt = 24; sc->dt_secs = 24 * 60 * 60; sc->itt_day = 0; }
    if ((sc->event_id[0] > 0) && (sc->event_id[1] == 0)) sc->event_id_counter += 1;

    /* ---- per-cell physics ---- */
    step_ctx ctx;
    ctx.st = st;
    ctx.dt = sc->dt;
    ctx.cond1 = (sc->event_id[0] == 0) && (sc->event_id[1] >= 1);
    ctx.cond2_scalar = (sc->event_id[0] >= 1);
    ctx.cond3_scalar = (sc->event_id[0] == sc->event_id[1]);
    ctx.cond4 = (sc->event_id[0] >= 1) && (sc->event_id[1] == 0);
    ctx.cond5 = (sc->event_id[1] >= 1);
    ctx.substeps = (int)nearbyint(sc->dt / (1.0 / 5)); /* npx.round: half to even, infiltration.py:513 */
    ctx.irrigation_month = (sc->month >= 4 && sc->month <= 9);

    int all_ok = 1;
#pragma omp parallel for schedule(static) reduction(&& : all_ok)
    for (int64_t i = 0; i < C; i++) {
        double f[RF_COUNT];
        for (int v = 0; v < RF_COUNT; v++) f[v] = planes[(size_t)v * (size_t)C + (size_t)i];
        cell_m1 p;
        p.z_wf = V(z_wf); p.z_wf_t0 = V(z_wf_t0); p.z_wf_t1 = V(z_wf_t1); p.y_mp = V(y_mp); p.y_sc = V(y_sc);
        p.S = V(S); p.S_rz = V(S_rz); p.S_ss = V(S_ss); p.S_s = V(S_s); p.z_root = V(z_root);
        p.prec = prec_m1[i]; /* prec[taum1]: the plane value before this step's aggregation */

        V(pet_res) = V(pet); /* :367-368 */
        calc_rain_int_top(f, &ctx);
        calc_rain_int_ground(f, &ctx);
        calc_snow_int_top(f, &ctx);
        calc_snow_int_ground(f, &ctx);
        calc_int(f);
        calc_evap_int_top(f);
        calc_evap_int_ground(f);
        calc_evap_dep(f);
        calc_evap_sur(f);
        calc_evap_soil(f);
        calc_transp(f, &ctx);
        calc_et_tail(f);
        calc_snow(f, &ctx);
        calculate_infiltration(f, &ctx, &p);
        calc_rise_of_saturation_water_table(f);
        calc_S_zsat(f);
        if (st->enable_lateral_flow) { /* subsurface_runoff.py:1455-1466 */
            calc_z_sat_layer(f);
            calc_potential_percolation_rz(f, &ctx, &p);
            calc_percolation_rz(f);
            calc_potential_lateral_subsurface_runoff(f, &ctx);
            calc_lateral_subsurface_runoff_rz(f);
            calc_potential_lateral_subsurface_runoff_ss(f);
            calc_potential_percolation_ss(f, &ctx);
            calc_lateral_subsurface_runoff_ss(f);
        } else { /* :1472-1478 */
            calc_potential_percolation_rz(f, &ctx, &p);
            calc_percolation_rz(f);
            calc_potential_percolation_ss(f, &ctx);
            calc_percolation_ss(f);
        }
        calc_cpr_rz(f, &ctx);
        if (st->enable_groundwater_boundary) calc_cpr_ss(f, &ctx); /* capillary_rise.py:346-358 */
        calc_storages(f, &ctx, &p);
        if (st->enable_groundwater_boundary) calc_q_re(f, &ctx);   /* roger.py:436-438 */
        int ok = calc_num_error_and_check(f, &ctx, &p);
        if (st->enable_groundwater_boundary) V(z_gw_m1) = V(z_gw); /* after_timestep of the setups: vs.z_gw[taum1] = vs.z_gw[tau] */
        all_ok = all_ok && ok;
        /* after_timestep: clamp tiny negatives (models/svat/svat.py:333-352) */
        if (st->clamp_small_negative_storage) {
            V(S_fp_rz) = ((V(S_fp_rz) > -1e-6) && (V(S_fp_rz) < 0)) ? 0 : V(S_fp_rz);
            V(S_lp_rz) = ((V(S_lp_rz) > -1e-6) && (V(S_lp_rz) < 0)) ? 0 : V(S_lp_rz);
            V(S_fp_ss) = ((V(S_fp_ss) > -1e-6) && (V(S_fp_ss) < 0)) ? 0 : V(S_fp_ss);
            V(S_lp_ss) = ((V(S_lp_ss) > -1e-6) && (V(S_lp_ss) < 0)) ? 0 : V(S_lp_ss);
        }
        for (int v = 0; v < RF_COUNT; v++) planes[(size_t)v * (size_t)C + (size_t)i] = f[v];
    }
    free(prec_m1);
    sc->itt += 1;
    sc->time += sc->dt_secs;
    sc->sanity_ok = !reduce_or((uint32_t)!all_ok);
    sc->event_id[0] = sc->event_id[1]; /* after_timestep */
#undef PL
}
