/* roger_b200_fields.h -- per-cell field ids of the SVAT hot path (public ABI).
 *
 * Every per-cell quantity lives in one structure-of-arrays "plane" of C doubles
 * (C = number of interior cells, cell index fastest).  The ids below are what
 * roger_svat_upload()/roger_svat_download() take; the names are the reference's
 * variable names (roger/variables.py:198-6403) so a host-side mirror can map
 * `vs.<name>[2:-2, 2:-2(, tau)]` <-> plane by name.
 *
 * Kinds:
 *   P  parameter  -- read-only inside a time step (set at setup / monthly by the host side)
 *   S  state      -- read and written by every step (the [tau] level; [taum1] is not stored:
 *                    the reference copies tau->taum1 after every step, models/svat/svat.py:187-384)
 *   F  flux       -- written by a step for diagnostics/parity only (never read back by the next step)
 *
 * X(kind, name)
 */
#ifndef ROGER_B200_FIELDS_H
#define ROGER_B200_FIELDS_H

#define ROGER_SVAT_FIELDS(X) \
    /* ---- parameters ---------------------------------------------------- */ \
    X(P, maskCatch) X(P, maskRiverLake) X(P, lu_id) \
    X(P, z_soil) X(P, z_root) X(P, dmpv) X(P, lmpv) \
    X(P, theta_ac) X(P, theta_ufc) X(P, theta_pwp) X(P, theta_sat) X(P, theta_fc) \
    X(P, ks) X(P, kf) X(P, ks_ss) X(P, sealing) X(P, z_gw) \
    X(P, S_int_top_tot) X(P, S_int_ground_tot) \
    X(P, throughfall_coeff_top) X(P, throughfall_coeff_ground) \
    X(P, basal_evap_coeff) X(P, basal_transp_coeff) \
    X(P, rew) X(P, tew) X(P, z_evap) \
    X(P, S_ac_rz) X(P, S_ufc_rz) X(P, S_pwp_rz) \
    X(P, S_ac_ss) X(P, S_ufc_ss) X(P, S_pwp_ss) X(P, S_pwp_s) \
    X(P, wfs) X(P, ha) X(P, n_salv) X(P, m_bc) X(P, lambda_bc) \
    X(P, mp_drain_area) X(P, theta_4) X(P, theta_27) X(P, z_sc_max) \
    /* oneD (lateral flow) parameters */ \
    X(P, dmph) X(P, slope) \
    X(P, v_mp_layer_1) X(P, v_mp_layer_2) X(P, v_mp_layer_3) X(P, v_mp_layer_4) \
    X(P, v_mp_layer_5) X(P, v_mp_layer_6) X(P, v_mp_layer_7) X(P, v_mp_layer_8) \
    /* groundwater boundary (settings.enable_groundwater_boundary): drainable porosity of the vadose zone below the soil */ \
    X(P, n0) \
    /* ---- state ---------------------------------------------------------- */ \
    X(S, prec) X(S, ta) X(S, pet) \
    X(S, S_int_top) X(S, S_int_ground) X(S, swe_top) X(S, swe_ground) X(S, swe_top_tot) \
    X(S, S_dep) X(S, S_snow) X(S, swe) X(S, z0) \
    X(S, S_fp_rz) X(S, S_lp_rz) X(S, S_fp_ss) X(S, S_lp_ss) X(S, z_sat) \
    X(S, de) X(S, k_stress_evap) X(S, k_stress_transp) \
    X(S, z_wf) X(S, z_wf_t0) X(S, z_wf_t1) X(S, z_wf_fc) X(S, y_mp) X(S, y_sc) X(S, no_wf) \
    X(S, theta_d) X(S, theta_d_rel) X(S, theta_d_t0) X(S, theta_d_rel_t0) X(S, theta_d_fp) \
    X(S, z_sc) \
    X(S, prec_event_csum) X(S, t_event_csum) X(S, inf_mat_event_csum) X(S, inf_mat_pot_event_csum) \
    X(S, inf_mp_event_csum) X(S, inf_sc_event_csum) X(S, t_sat) X(S, inf_mat_rz) \
    X(S, theta_rz) X(S, theta_ss) X(S, theta) \
    X(S, k_rz) X(S, k_ss) X(S, k) X(S, h_rz) X(S, h_ss) X(S, h) \
    X(S, S_sur) X(S, S_rz) X(S, S_ss) X(S, S_s) X(S, S) \
    /* groundwater boundary: vadose-zone reservoir below the soil and its capacity (groundwater_flow.py:99-160); z_gw_m1 is \
       vs.z_gw[taum1], the depth of the water table of the previous step (capillary_rise.py:262-272) */ \
    X(S, S_vad) X(S, S_vad_tot) X(S, z_gw_m1) \
    /* ---- fluxes / diagnostics ------------------------------------------ */ \
    X(F, pet_res) \
    X(F, rain_top) X(F, int_rain_top) X(F, rain_ground) X(F, int_rain_ground) \
    X(F, snow_top) X(F, int_snow_top) X(F, snow_ground) X(F, int_snow_ground) \
    X(F, int_top) X(F, int_ground) X(F, int_prec) \
    X(F, evap_int_top) X(F, evap_int_ground) X(F, evap_int) X(F, evap_dep) X(F, evap_sur) \
    X(F, evap_coeff) X(F, pevap_soil) X(F, evap_soil) \
    X(F, transp_coeff) X(F, pt) X(F, ptransp) X(F, ptransp_res) X(F, transp) \
    X(F, aet_soil) X(F, aet) \
    X(F, snow_melt_top) X(F, snow_melt_drip) X(F, snow_melt_ground) X(F, snow_melt) X(F, q_snow) \
    X(F, pi_gr) X(F, pi_m) X(F, Fs) \
    X(F, inf_mat_pot) X(F, inf_mat) X(F, lmpv_non_sat) X(F, z_sc_non_sat) \
    X(F, inf_mp) X(F, inf_mp_rz) X(F, inf_mp_ss) X(F, inf_ss) X(F, inf_sc) \
    X(F, inf_sc_rz) X(F, inf_rz) X(F, inf) \
    X(F, q_hof) X(F, q_sof) X(F, q_sur) \
    X(F, S_zsat) X(F, S_zsat_rz) X(F, S_zsat_ss) \
    X(F, q_pot_rz) X(F, q_rz) X(F, q_pot_ss) X(F, q_ss) X(F, cpr_rz) \
    X(F, S_fp_s) X(F, S_lp_s) \
    X(F, dS_rz) X(F, dS_ss) X(F, dS_s) X(F, dS) X(F, irr_demand) \
    X(F, dS_num_error) X(F, dS_rz_num_error) X(F, dS_ss_num_error) \
    /* oneD (lateral flow) diagnostics */ \
    X(F, q_sub_mat_pot) X(F, q_sub_mp_pot) X(F, q_sub_pot) X(F, q_sub_mat_share) X(F, q_sub_mp_share) \
    X(F, q_sub_rz) X(F, q_sub_mat_rz) X(F, q_sub_mp_rz) X(F, q_sub_mp_pot_rz) \
    X(F, q_sub_mat_pot_ss) X(F, q_sub_mp_pot_ss) X(F, q_sub_pot_ss) \
    X(F, q_sub_ss) X(F, q_sub_mat_ss) X(F, q_sub_mp_ss) \
    X(F, q_sub_mat) X(F, q_sub_mp) X(F, q_sub) \
    X(F, z_sat_layer_1) X(F, z_sat_layer_2) X(F, z_sat_layer_3) X(F, z_sat_layer_4) \
    X(F, z_sat_layer_5) X(F, z_sat_layer_6) X(F, z_sat_layer_7) X(F, z_sat_layer_8) \
    /* groundwater boundary: capillary rise from the water table into the subsoil, groundwater recharge */ \
    X(F, cpr_ss) X(F, q_re)

#define ROGER_FIELD_ENUM(kind, name) RF_##name,
enum roger_svat_field { ROGER_SVAT_FIELDS(ROGER_FIELD_ENUM) RF_COUNT };
#undef ROGER_FIELD_ENUM

enum roger_field_kind { ROGER_KIND_P = 0, ROGER_KIND_S = 1, ROGER_KIND_F = 2 };

#endif /* ROGER_B200_FIELDS_H */
