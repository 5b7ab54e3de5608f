/* roger_b200_sas.h -- C ABI of the offline StorAge-selection (SAS) oxygen-18 transport step.
 *
 * Replaces, for the deterministic solver, the body of transport.svat_transport_model_deterministic
 * (roger/core/transport.py:948-991): infiltration into the root zone, soil evaporation, transpiration,
 * root-zone percolation, infiltration into the subsoil, subsoil percolation, capillary rise
 * (each: cumulative StorAge -> SAS function -> backward travel-time distribution -> isotope mixing),
 * storage concentrations, ageing.  One call = one daily time step for every cell.
 *
 * Age-ranked arrays are (cell, age) with the age axis contiguous, exactly the interior of the reference's
 * `vs.sa_rz[2:-2, 2:-2, tau, :]` (roger/variables.py:2500-2525).
 */
#ifndef ROGER_B200_SAS_H
#define ROGER_B200_SAS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* per-cell daily forcing planes (roger/models/svat_oxygen18 set_forcing; benchmarks/SVATOXYGEN18_benchmark.py:385-398) */
enum roger_sas_forcing {
    RSF_inf_mat_rz = 0, RSF_inf_pf_rz, RSF_inf_pf_ss, RSF_evap_soil, RSF_transp, RSF_q_rz, RSF_q_ss, RSF_cpr_rz,
    RSF_C_in, /* isotope concentration of the input (after the snow-pack mixing of set_forcing) */
    RSF_COUNT
};

/* the five SAS-controlled fluxes, index into the sas_params block: (process, cell, 8) */
enum roger_sas_process { RSP_evap_soil = 0, RSP_transp, RSP_q_rz, RSP_q_ss, RSP_cpr_rz, RSP_COUNT };

/* per-cell scalar results of a step */
enum roger_sas_output {
    RSO_C_inf_mat_rz = 0, RSO_C_inf_pf_rz, RSO_C_inf_pf_ss, RSO_C_evap_soil, RSO_C_transp, RSO_C_q_rz, RSO_C_q_ss, RSO_C_cpr_rz,
    RSO_C_iso_inf_mat_rz, RSO_C_iso_inf_pf_rz, RSO_C_iso_inf_pf_ss, RSO_C_iso_evap_soil, RSO_C_iso_transp, RSO_C_iso_q_rz,
    RSO_C_iso_q_ss, RSO_C_iso_cpr_rz,
    RSO_C_rz, RSO_C_ss, RSO_C_s, RSO_C_iso_rz, RSO_C_iso_ss, RSO_C_iso_s,
    RSO_COUNT
};

/* per-cell age statistics and balance errors of a step (settings.age_statistics != 0):
 * transport.calculate_age_statistics_* (roger/core/transport.py:59-312) and numerics.calc_dS_num_error / calc_dC_num_error
 * (roger/core/numerics.py:424-445, :476-516).  Percentile ages follow numpy.interp(p, CDF[1:], 1..ages). */
enum roger_sas_statistic {
    RSS_tt10_transp = 0, RSS_tt25_transp, RSS_tt50_transp, RSS_tt75_transp, RSS_tt90_transp, RSS_ttavg_transp,
    RSS_tt10_q_ss, RSS_tt25_q_ss, RSS_tt50_q_ss, RSS_tt75_q_ss, RSS_tt90_q_ss, RSS_ttavg_q_ss,
    RSS_rt25_rz, RSS_rt50_rz, RSS_rt75_rz, RSS_rtavg_rz,
    RSS_rt25_ss, RSS_rt50_ss, RSS_rt75_ss, RSS_rtavg_ss,
    RSS_rt10_s, RSS_rt25_s, RSS_rt50_s, RSS_rt75_s, RSS_rt90_s, RSS_rtavg_s,
    RSS_dS_num_error, RSS_dC_num_error,
    RSS_sanity_ok, /* 1 / 0 per cell: water balance, solute balance and storage bounds of numerics.sanity_check (numerics.py:1051-1122);
                      the reference's scalar result is the AND over all cells */
    RSS_COUNT
};

/* ---- nitrate (settings.solute = 2): the anion transport plus the soil nitrogen cycle on age-ranked mineral nitrogen
 *      (roger/core/nitrate.py:68-696 calc_nitrogen_cycle_kernel and its process kernels; transport.py:948-991 with
 *      settings.enable_nitrate, :809-824 ageing of Nmin, :3623-3657 after_substep_nitrate).  The soil temperature of
 *      nitrate.py:6-66 is a per-cell scalar with a sine and an exponential of forcing statistics: the host side computes it
 *      (roger_b200/nitrate.py, NumPy as the reference) and hands it over as a forcing plane. */
enum roger_nitrate_param { /* per-cell parameter planes, roger/variables.py (km_denit_rz ... damp_soil_temp) */
    RNP_kfix_rz = 0, RNP_kmin_rz, RNP_kmin_ss, RNP_km_nit_rz, RNP_dmax_nit_rz, RNP_kngl_rz, RNP_km_denit_rz, RNP_dmax_denit_rz,
    RNP_km_denit_ss, RNP_dmax_denit_ss, RNP_z_soil, RNP_S_sat_rz, RNP_S_sat_ss, RNP_COUNT
};
enum roger_nitrate_forcing { /* per-cell daily planes besides the RSF_* ones */
    RNF_temp_soil = 0, /* vs.temp_soil[tau] of the day (nitrate.py:6-66)                 */
    RNF_z_root,        /* vs.z_root[tau]                                                  */
    RNF_lu_id,         /* vs.LU_ID[:, :, itt]: nitrogen-fixing crops (nitrate.py:339-343) */
    RNF_nup,           /* vs.nup: potential ammonium uptake                               */
    RNF_COUNT
};
enum roger_nitrate_output { /* per-cell results of the nitrogen cycle of the last step */
    RNO_nit_s = 0, RNO_denit_s, RNO_min_s, RNO_ngas_s, RNO_nfix_s, RNO_nh4_up, RNO_Nmin_s, RNO_COUNT
};

enum roger_crop_output { /* crop variant: the redistribution fluxes as applied (clamped to the source layer's water), their solute
                            concentration and mass (vs.re_rg, vs.C_re_rg, vs.M_re_rg, ... of roger/core/crop.py:2461-2828) */
    RCO_re_rg = 0, RCO_re_rl, RCO_C_re_rg, RCO_C_re_rl, RCO_M_re_rg, RCO_M_re_rl, RCO_COUNT
};

typedef struct roger_sas_settings {
    int32_t ages;          /* settings.ages (number of age bins)                       */
    int32_t substeps;      /* settings.sas_solver_substeps                              */
    double VSMOW_conc18O;  /* settings.VSMOW_conc18O                                    */
    double d18O_min;       /* settings.d18O_min                                         */
    double d18O_max;       /* settings.d18O_max                                         */
    int32_t store_tt;      /* keep the five travel-time distributions tt_* (diagnostics) */
    int32_t age_statistics; /* compute the RSS_* planes every step (settings.enable_age_statistics + numerical errors) */
    int32_t solute;         /* 0: stable isotope (oxygen-18 / deuterium: msa holds bin concentrations);
                               1: conservative anion (bromide / chloride: msa holds bin solute MASS, the *_transport_anion_kernel
                                  family of the reference; the RSO_C_iso_* result planes then hold the solute masses M_*);
                               2: nitrate = the anion transport plus the soil nitrogen cycle (two more age-ranked arrays Nmin_rz,
                                  Nmin_ss: state 4 and 5 of roger_sas_upload_state); deterministic solver;
                               3: virtual tracer (settings.enable_virtualtracer): the anion transport whose soil evaporation carries the
                                  solute with partition coefficient 1 (evapotranspiration.py:723-791: RSO_C_evap_soil, and M_evap_soil in
                                  RSO_C_iso_evap_soil) and whose solute balance counts it (numerics.py:517-549); deterministic solver, and the Euler
                                  scheme, whose anion branch it shares unchanged (transport.py:2308-2331) */
    int32_t solver;         /* 0: deterministic sub-step solver (transport.py:948-991); 1: explicit Euler with `substeps` steps per day
                               (svat_transport_model_euler, transport.py:2063-2413; statistics / balance errors are those of the day's
                               last Euler step, as the reference's output shows them -- isotope and anion branch alike);
                               2: RK4 (svat_transport_model_rk4, transport.py:1138-2047), isotope branch only: four stages of the Euler
                               building blocks per step on stage copies of the StorAge (nine more scratch arrays of cells x ages) */
} roger_sas_settings;

typedef struct roger_sas roger_sas; /* opaque handle */

#define ROGER_ENOTSUP 95 /* a combination this backend does not evaluate: an unknown SAS function id, gamma parameters outside
                            0 < a <= 170, 0 <= b <= 50, RK4 with a solute other than the isotope, ... (the message says which) */

int roger_sas_create(roger_sas **out, int64_t n_cells, const roger_sas_settings *settings, int device);
void roger_sas_destroy(roger_sas *h);
/* age-ranked state: which = 0 sa_rz, 1 sa_ss, 2 msa_rz, 3 msa_ss (nitrate: 4 Nmin_rz, 5 Nmin_ss); host buffers of n_cells*ages doubles */
int roger_sas_upload_state(roger_sas *h, int which, const double *host);
int roger_sas_download_state(roger_sas *h, int which, double *host);
/* raw device address of the age-major array (A planes of n_cells doubles; age bin j is plane (base + j) mod A,
 * base = roger_sas_age_base()) */
uint64_t roger_sas_state_ptr(roger_sas *h, int which);
int roger_sas_age_base(roger_sas *h);
/* initial condition of the benchmark / model setups computed on the device (benchmarks/SVATOXYGEN18_benchmark.py:172-218,
 * roger/models/svat_oxygen18 set_initial_conditions): sa[1:] = diff(linspace(0, S, ages)), sa[0] = 0, msa = conc0, msa[0] = 0.
 * S_rz, S_ss: n_cells doubles (host). */
int roger_sas_init_linear(roger_sas *h, const double *S_rz, const double *S_ss, double conc0);
/* maskCatch (n_cells) and SAS parameters (RSP_COUNT * n_cells * 8), roger/variables.py:4275.  Every family of roger/core/sas.py:
 * uniform 1, dirac 2, Kumaraswamy 3 / 31-37, gamma 4 (0 < a <= 170, 0 <= b <= 50), exponential 51 / 52, power 6 / 61 / 62. */
int roger_sas_set_parameters(roger_sas *h, const double *mask_catch, const double *sas_params);
/* anion transport only: partition coefficients of transpiration and percolation (vs.alpha_transp, vs.alpha_q), land use
 * and root-zone saturation storage (crop nutrient uptake stops above 80 % saturation, evapotranspiration.py:931-938);
 * n_cells doubles each */
int roger_sas_set_anion_parameters(roger_sas *h, const double *alpha_transp, const double *alpha_q, const double *lu_id,
                                   const double *S_sat_rz);
/* nitrate only: RNP_COUNT parameter planes of n_cells doubles, grid spacing (settings.dx, settings.dy), settings.zroot_to_zsoil_max
 * and the time step in hours (vs.dt = 24) */
int roger_sas_set_nitrate_parameters(roger_sas *h, const double *params, double dx, double dy, double zroot_to_zsoil_max, double dt_hours);
/* nitrate only: the day's RNF_COUNT planes of n_cells doubles, before roger_sas_step */
int roger_sas_set_nitrate_forcing(roger_sas *h, const double *planes);
/* nitrate only: RNO_COUNT planes of n_cells doubles from the last step */
int roger_sas_download_nitrate_outputs(roger_sas *h, double *host);
/* Crop variant of the offline model (settings.enable_crop_phenology: svat_crop_transport_model_deterministic,
 * roger/core/transport.py:1042-1087): every day then opens with crop.calculate_redistribution_transport (roger/core/crop.py:2829-2867,
 * anion kernels :2461-2558 and :2729-2828) -- water and solute move from the subsoil to the root zone after root growth (re_rg) and
 * back after root loss (re_rl), each a SAS-controlled withdrawal with its own parameter row (vs.sas_params_re_rg / _re_rl,
 * n_cells * 8 doubles each, same row layout as roger_sas_set_parameters).  Anion, nitrate and virtual-tracer solutes, deterministic solver, with or without age
 * statistics (ROGER_ENOTSUP otherwise).  The day's fluxes vs.re_rg, vs.re_rl (n_cells doubles each, host; an input of the offline
 * model, read from the SVAT-CROP output by the reference's setups) must be set before every roger_sas_step*. */
int roger_sas_set_crop_parameters(roger_sas *h, const double *sas_params_re_rg, const double *sas_params_re_rl);
int roger_sas_set_redistribution(roger_sas *h, const double *re_rg, const double *re_rl);
/* RCO_COUNT planes of n_cells doubles from the last step */
int roger_sas_download_redistribution(roger_sas *h, double *host);
/* one daily step.  forcing: RSF_COUNT planes of n_cells doubles (host memory). */
int roger_sas_step(roger_sas *h, const double *forcing);
/* same with forcing already on the device (raw address of RSF_COUNT * n_cells doubles) */
int roger_sas_step_device(roger_sas *h, uint64_t forcing);
/* per-cell scalar results of the last step: RSO_COUNT planes of n_cells doubles */
int roger_sas_download_outputs(roger_sas *h, double *host);
/* RSS_COUNT planes of n_cells doubles from the last step (needs settings.age_statistics) */
int roger_sas_download_statistics(roger_sas *h, double *host);
/* the previous step's storage concentration C_s per cell, which the solute balance error differences against
 * (vs.C_s[taum1]); set once after the initial state, then maintained by the steps.  n_cells doubles. */
int roger_sas_set_balance_reference(roger_sas *h, const double *C_s_prev);
/* storage bounds of the sanity check (vs.S_sat_rz, vs.S_sat_ss, vs.S_pwp_rz, vs.S_pwp_ss; n_cells doubles each) and its
 * tolerances (settings.atol, settings.rtol).  Default: no upper bound, atol = rtol = 1e-2. */
int roger_sas_set_sanity_bounds(roger_sas *h, const double *S_sat_rz, const double *S_sat_ss, const double *S_pwp_rz,
                                const double *S_pwp_ss, double atol, double rtol);
/* travel-time distribution of process p from the last step (needs settings.store_tt): n_cells*ages doubles */
int roger_sas_download_tt(roger_sas *h, int process, double *host);
/* out[i] = roger_pow(x[i], y[i]) -- or roger_exp(x[i]) when y is NULL -- evaluated by the device build of
 * roger_b200/csrc/roger_pow.h (host pointers, n doubles each): the parity tests use it to prove that the device and
 * the host evaluate the shared power function to identical bits. */
int roger_pow_device(const double *x, const double *y, double *out, int64_t n, int device);
/* out[i] = P(a[i], x[i]), the regularised lower incomplete gamma function of the gamma SAS family (sas.py:153,
 * scipy.special.gammainc), evaluated by the device build of roger_b200/csrc/roger_gamma.h (host pointers, n doubles each) */
int roger_gammainc_device(const double *a, const double *x, double *out, int64_t n, int device);
/* self-test of the kernel's hoisted-reciprocal division (sas_device.cuh div_by) against the division instruction on n
 * pseudo-random operand pairs; *mismatches must come back 0 */
int roger_selftest_division(int64_t n, uint64_t seed, int device, int64_t *mismatches);
int roger_sas_synchronize(roger_sas *h);
int roger_sas_last_timing(roger_sas *h, double *kernel_ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* ROGER_B200_SAS_H */
