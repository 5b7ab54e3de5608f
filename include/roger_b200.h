/* roger_b200.h -- C ABI of the B200-native RoGeR SVAT / SAS compute backend.
 *
 * The reference (Hydrology-IFH/roger) has no FFI: its plug-in boundary is Python-level
 * (roger/backend.py:3 BACKENDS, roger/core/operators.py:148-186, roger/routines.py:239-383,
 * roger/roger.py:355-489 RogerSetup.step).  This header is what a `cuda` backend branch of
 * those files binds through ctypes.  Plain pointers and sizes only,
 * no torch / CUDA types in any signature; all device memory is owned by the handle or by
 * the caller (raw device addresses passed as uint64_t).
 *
 * Return value of every int function: 0 on success, negative errno-style code on failure
 * (-ROGER_EINVAL bad argument, -ROGER_ECUDA CUDA runtime error, -ROGER_ENODEV no CUDA device).
 * roger_last_error() returns a human readable message for the calling thread.
 */
#ifndef ROGER_B200_H
#define ROGER_B200_H

#include <stdint.h>
#include <stddef.h>
#include "roger_b200_fields.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ROGER_EINVAL 22
#define ROGER_ENODEV 19
#define ROGER_ECUDA 5
#define ROGER_SLOTS_PER_DAY 144 /* 10-minute forcing slots, roger/models/svat/svat.py:150-174 */

/* Model switches and physical constants read by the step (roger/settings.py:18-136). */
typedef struct roger_svat_settings {
    double ta_fm;               /* settings.ta_fm   freeze/melt threshold (degC)            */
    double hpi;                 /* settings.hpi     heavy-rain threshold (mm/10 min)        */
    double end_event;           /* settings.end_event (s)                                   */
    double sf;                  /* settings.sf      degree-day factor                       */
    double rmax;                /* settings.rmax    liquid water retention of snow (%)      */
    double r_mp;                /* settings.r_mp    macropore radius (mm)                   */
    double l_sc;                /* settings.l_sc    shrinkage crack length (mm/m2)          */
    double pi;                  /* settings.pi                                              */
    double transp_water_stress; /* settings.transp_water_stress                             */
    double dx;                  /* settings.dx      (lateral flow only)                     */
    double atol, rtol;          /* settings.atol/rtol used by numerics.sanity_check         */
    int32_t enable_lateral_flow;                       /* oneD model                        */
    int32_t enable_macropore_lower_boundary_condition; /* infiltration.py:1013              */
    int32_t clamp_small_negative_storage; /* after_timestep clamp of S_fp/S_lp in (-1e-6,0),
                                             models/svat/svat.py:333-352 (SVAT yes, oneD no) */
    int32_t forcing_uniform;    /* 1: prec/ta/pet day blocks are 144 values shared by all cells
                                   (benchmarks/SVAT_benchmark.py:162-170); 0: (C,144) per cell   */
    int32_t enable_groundwater_boundary; /* settings.enable_groundwater_boundary (roger/roger.py:425-438): the water table
                                   z_gw is a boundary condition the host sets (plane z_gw, m below the surface); a step then
                                   adds the capillary rise into the subsoil (calc_cpr_ss, capillary_rise.py:176-324), the
                                   groundwater recharge through the vadose-zone reservoir (calc_q_re,
                                   groundwater_flow.py:99-160) and the balance terms of numerics.py:271-290, :396-420;
                                   single steps only (no day-blocked stepping) */
    int32_t reserved_;
    double kf_max;              /* settings.kf_max (mm/h), groundwater boundary only          */
} roger_svat_settings;

/* Global (grid-wide) scalars of the adaptive time stepping, roger/variables.py:198-460.
 * event_id[0] = [taum1], event_id[1] = [tau]. */
typedef struct roger_svat_scalars {
    int64_t time;             /* vs.time (s)            */
    int64_t itt;              /* vs.itt                 */
    int64_t dt_secs;          /* vs.dt_secs             */
    int64_t itt_day;          /* vs.itt_day             */
    int64_t event_id[2];      /* vs.event_id            */
    int64_t event_id_counter; /* vs.event_id_counter    */
    int64_t time_event0;      /* vs.time_event0         */
    int64_t month;            /* vs.month[tau] (1..12)  */
    int64_t sanity_ok;        /* result of numerics.sanity_check for the last step (1/0) */
    double dt;                /* vs.dt (h)              */
} roger_svat_scalars;

typedef struct roger_svat roger_svat; /* opaque handle */

/* ---- lifetime ------------------------------------------------------------------ */
int roger_svat_create(roger_svat **out, int64_t n_cells, const roger_svat_settings *settings, int device);
void roger_svat_destroy(roger_svat *h);
const char *roger_last_error(void);
int roger_svat_field_count(void);
const char *roger_svat_field_name(int field);
int roger_svat_field_kind(int field);

/* ---- host <-> device mirrors of named planes (replaces RogerVariables.__setattr__/__getattr__
 *      for the fields the step owns, roger/state.py:213-251).  `host` holds n_cells doubles. */
int roger_svat_upload(roger_svat *h, int field, const double *host);
int roger_svat_download(roger_svat *h, int field, double *host);
/* Asynchronous diagnostic read-back (replaces the reference's threaded netCDF hand-over, roger/io_tools/netcdf.py:145-210):
 * snapshots the plane as of the steps issued so far and copies it to `host_pinned` (page-locked, n_cells doubles) on a
 * second stream while later steps run.  The buffer is valid after roger_svat_download_wait(). */
int roger_svat_download_async(roger_svat *h, int field, double *host_pinned);
int roger_svat_download_wait(roger_svat *h);
/* how a snapshot leaves the device: 1 (default) an SM copy kernel writing into the caller's page-locked buffer through its mapped
 * device pointer, on a high-priority stream beside the step kernels (the copy engine's transfer is starved by a DRAM-saturating
 * kernel: 3 GB/s measured against 33 GB/s idle); 0 the copy engine (also the automatic fallback for buffers that are not mapped) */
int roger_svat_set_d2h_mode(roger_svat *h, int mode);
/* Makes sure `n` device snapshot planes exist so that roger_svat_download_async / _diagnostics_async never allocate inside
 * a stepping loop; a driver calls it once when it configures its output (n = planes handed over per output time).  The
 * copy stream, its events and the fence ring are created by roger_svat_create. */
int roger_svat_reserve_snapshots(roger_svat *h, int n);
/* Device layout of the planes, for zero-copy use from other kernels: tile-major, [tile][plane][cell] with `tile_cells`
 * consecutive cells per tile; plane `f` (a P or S field id) of cell i is the double at
 *   ps_base + (((i / tile_cells) * n_ps_planes + roger_svat_field_slot(f)) * tile_cells + i % tile_cells) * 8,
 * flux field f likewise with flux_base, n_flux_planes and f - n_ps_planes (flux_base is 0 until a flux plane was asked for).
 * i is the DEVICE position of a cell: after roger_svat_optimize_cell_order it is no longer the caller's cell index. */
typedef struct roger_svat_layout {
    uint64_t ps_base, flux_base;
    int64_t n_tiles;
    int32_t tile_cells, n_ps_planes, n_flux_planes, reserved;
} roger_svat_layout;
int roger_svat_tile_layout(roger_svat *h, roger_svat_layout *out);

/* physical slot of a P / S field inside a tile (see roger_svat_layout); -1 for flux fields */
int roger_svat_field_slot(int field);
/* Re-orders the cells on the device by infiltration capacity (sort key ks * (1 - sealing), cells outside the catchment last)
 * so that the threads of a warp hold similar cells: the macropore / shrinkage-crack sub-step loops of
 * roger/core/infiltration.py:430-1318 run only for ponded cells, and with spatially uncorrelated parameters every warp
 * would pay for a few of them.  Columns are independent (no lateral routing on this path), so the results are identical;
 * every entry point keeps speaking the caller's cell numbering.  Call after the parameters are uploaded and before the
 * first forcing day is handed over.  (The reference has no counterpart: its arrays are (x, y) grids.) */
int roger_svat_optimize_cell_order(roger_svat *h);

int roger_svat_set_scalars(roger_svat *h, const roger_svat_scalars *s);
int roger_svat_get_scalars(roger_svat *h, roger_svat_scalars *s);

/* ---- forcing: the day block that RogerSetup.set_forcing writes into vs.prec_day/ta_day/pet_day
 *      at midnight (roger/models/svat/svat.py:150-174).  Host pointers; 144 doubles each when
 *      settings.forcing_uniform, else n_cells*144 in the reference's (cell, slot) order; per-cell blocks are re-tiled on
 *      the device to [tile][slot][cell] so that the 144-slot aggregations of adaptive_time_stepping.py:385-437 read
 *      coalesced.  The copy is asynchronous on the handle's stream from the caller's (ideally pinned) buffer. */
int roger_svat_set_forcing_day(roger_svat *h, const double *prec_day, const double *ta_day,
                               const double *pet_day);
/* same, from buffers already resident on the device (raw addresses) */
int roger_svat_set_forcing_day_device(roger_svat *h, uint64_t prec_day, uint64_t ta_day, uint64_t pet_day);

/* ---- multi-GPU: the grid-wide any()/all() of adaptive_time_stepping.py:38-81/:191-201 and the global_and of
 *      numerics.sanity_check are evaluated on per-rank flag words; with the grid partitioned over several
 *      processes the caller supplies a collective that ORs `n` flag words in place across all ranks
 *      (replaces roger/distributed.py global_and/global_or; NCCL or gloo from torch.distributed).
 *      Every rank calls it the same number of times per step. */
typedef void (*roger_flag_reduce_fn)(uint32_t *flags, int n, void *user);
int roger_svat_set_flag_reducer(roger_svat *h, roger_flag_reduce_fn fn, void *user);
/* The same exchange without leaving the device queue: one ncclAllGather of the flag words on the handle's compute stream
 * (NCCL over NVLink), OR-ed on the host behind the read-back every step performs anyway -- no host callback, no extra
 * synchronisation.  Rank 0 obtains an id with roger_nccl_unique_id (128 bytes) and distributes it by any means; every rank
 * then calls roger_svat_enable_nccl_flags with it.  NCCL is bound at run time (the libnccl.so.2 the process already carries);
 * -ROGER_ENODEV if none is loadable.  Takes precedence over a flag reducer callback. */
int roger_nccl_unique_id(uint8_t *id128);
int roger_svat_enable_nccl_flags(roger_svat *h, const uint8_t *id128, int rank, int world);

/* ---- the hot path: ONE adaptive time step = adaptive_time_stepping + interception + ET + snow +
 *      infiltration + percolation (+ lateral) + capillary rise + storages + numerical error +
 *      after_timestep roll (roger/roger.py:389-485).  Updates scalars (time, itt, dt, ...).
 *      write_fluxes != 0 additionally stores every F-kind plane (parity / diagnostics mode). */
int roger_svat_step(roger_svat *h, int write_fluxes);
/* run adaptive steps until vs.time reaches the next midnight (at most 144 steps); returns the
 * number of steps taken through *n_steps. */
int roger_svat_run_day(roger_svat *h, int write_fluxes, int *n_steps);
/* up to max_steps adaptive steps, never past the next midnight; *n_steps = steps taken.  With spatially uniform forcing and
 * no snow anywhere the scheduler lays out the coming steps in advance and runs them in ONE kernel launch (state re-read
 * from L2 instead of HBM); results are identical to single steps. */
int roger_svat_run_steps(roger_svat *h, int write_fluxes, int max_steps, int *n_steps);
/* vs.dt_secs of every step the last roger_svat_step / roger_svat_run_steps / roger_svat_run_day call took (at most `capacity`
 * values are written); returns their number */
int roger_svat_last_step_dts(roger_svat *h, int *dt_secs, int capacity);
/* how such pre-scheduled steps are executed: 0 = never batch (one launch + flag read-back per step), 1 = ONE launch of the
 * multi-step kernel per batch (default up to 262144 cells), 2 = the ordinary step kernel launched once per step, back to
 * back, with a single flag read-back (and, on several ranks, a single flag exchange) per batch (default for larger grids;
 * also with write_fluxes and diagnostics accumulators); max_block bounds the steps per batch (2..144) */
int roger_svat_set_block_stepping(roger_svat *h, int enabled, int max_block);
/* ---- diagnostics accumulators on the device (roger/diagnostics/{rate,average,maximum,minimum}.py): the named planes are
 *      folded into accumulator planes after every step (e.g. the daily sums that drive the offline transport model,
 *      test/make_data_for_svat_transport.py:279-300).  Accumulating a flux plane switches the steps to the
 *      flux-storing kernels. */
#define ROGER_DIAG_RATE 0    /* sum over the interval,                      roger/diagnostics/rate.py:72-81     */
#define ROGER_DIAG_AVERAGE 1 /* mean over the steps of the interval,        roger/diagnostics/average.py:71-97  */
#define ROGER_DIAG_MAXIMUM 2 /* running maximum of a 2-D variable,          roger/diagnostics/maximum.py:76-92  */
#define ROGER_DIAG_MINIMUM 3 /* running minimum of a 2-D variable,          roger/diagnostics/minimum.py        */
#define ROGER_DIAG_MAXIMUM_T 4 /* variables with a time-step dimension ((x, y, timesteps) in the reference): the maximum of the
                                  last two time levels, overwritten every step,  roger/diagnostics/maximum.py:60-75 */
#define ROGER_DIAG_MINIMUM_T 5 /* likewise the minimum,                          roger/diagnostics/minimum.py        */
int roger_svat_set_diagnostics(roger_svat *h, const int *fields, const int *modes, int n);
int roger_svat_download_diagnostics(roger_svat *h, double *host, int reset);
/* The same hand-over off the critical path (the reference's writer thread, io_tools/netcdf.py:145-210): the accumulators are
 * snapshotted on the compute stream, optionally reset, and copied to `host_pinned` on the copy stream while later steps run.
 * `*count` receives the number of accumulated steps (AVERAGE accumulators are delivered as SUMS: divide by it).
 * roger_svat_download_fence() marks the copy stream after the transfers issued so far and returns a token (>= 0);
 * roger_svat_fence_wait(token) blocks until that point is reached and may be called from another thread (it touches no
 * handle state besides the token's event).  roger_svat_download_wait() still recycles the snapshot buffers. */
int roger_svat_download_diagnostics_async(roger_svat *h, double *host_pinned, int reset, int64_t *count);
int roger_svat_download_fence(roger_svat *h);
int roger_svat_fence_wait(roger_svat *h, int token);
/* RATE only (modes = NULL) */
int roger_svat_set_rate_fields(roger_svat *h, const int *fields, int n);
int roger_svat_download_rates(roger_svat *h, double *host, int reset);
uint64_t roger_svat_rates_ptr(roger_svat *h);
/* In-memory checkpoint / rollback of the whole run state (planes, scheduler scalars and flags, the uniform forcing day); the
 * device-resident counterpart of the reference's restart file (roger/restart.py:32-174).  A per-cell forcing day block has to
 * be handed over again after a rollback. */
int roger_svat_checkpoint(roger_svat *h);
int roger_svat_rollback(roger_svat *h);
int roger_svat_synchronize(roger_svat *h);
/* self-test hook (no device needed): hourly window sum / nanmean of a 144-slot forcing row (adaptive_time_stepping.py:410-437)
 * by the six-load form the kernels use and by the 144-slot form; out4 = {sum, sum_full, nanmean, nanmean_full} */
int roger_selftest_window_sums(const double *row144, int itt_day, double *out4);
/* CUDA-event time (ms) spent in kernels of the most recent roger_svat_step / run_day call and the
 * number of kernel launches it made (bench.py roofline / gpu_launches). */
int roger_svat_last_timing(roger_svat *h, double *kernel_ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* ROGER_B200_H */
