// svat_api.cu -- kernels, host-side adaptive-time-stepping scheduler and the C ABI of the SVAT / oneD
// hot path (include/roger_b200.h).  sm_100a only; there is no CPU fallback: every entry point
// fails with -ROGER_ENODEV / -ROGER_ECUDA if no usable CUDA device is present.
//
// Data layout in HBM: tile-major structure of arrays, [tile of 256 cells][plane][cell] (svat_device.cuh: Io): a warp's
// 32 threads read / write 256 contiguous bytes per field, a block's whole working set is one contiguous region and a
// plane access is base + immediate.  Parameter (P) and state (S) planes live in one allocation; flux (F) planes are
// allocated lazily, only when a caller asks for diagnostics.  Per-cell forcing day blocks arrive in the reference's
// (cell, slot) order and are re-tiled to [tile][slot][cell] on upload, so the 144-slot aggregations read coalesced.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "svat_device.cuh"

namespace roger {

// ------------------------------------------------------------------------------------------------
// field tables
#define ROGER_NAME(kind, name) #name,
static const char* const kFieldName[RF_COUNT] = {ROGER_SVAT_FIELDS(ROGER_NAME)};
#undef ROGER_NAME
#define ROGER_KIND(kind, name) ROGER_KIND_##kind,
static const int kFieldKind[RF_COUNT] = {ROGER_SVAT_FIELDS(ROGER_KIND)};
#undef ROGER_KIND

__host__ __device__ constexpr bool kind_is_F(int id) { return id >= RF_pet_res; }  // enum order is P.., S.., F..
constexpr int kNumPS = RF_pet_res;
static_assert(RF_prec > RF_v_mp_layer_8 && RF_pet_res > RF_S, "field enum must be ordered P.., S.., F..");
constexpr int kNumF = RF_COUNT - RF_pet_res;

__host__ __device__ constexpr bool lat_only(int id) {
    return (id >= RF_dmph && id <= RF_v_mp_layer_8) || (id >= RF_q_sub_mat_pot && id <= RF_z_sat_layer_8);
}

enum : unsigned {
    FLAG_SWE_GT0 = 1u, FLAG_SWE_NOT_LE0 = 2u, FLAG_SWETOP_GT0 = 4u, FLAG_SWETOP_NOT_LE0 = 8u, FLAG_INSANE = 16u,
    // forcing day flags
    DAY_NOT_PREC_LE0 = 1u << 8, DAY_NOT_TA_GT = 1u << 9, DAY_SNOWFALL = 1u << 10, DAY_NOT_DRY_COLD = 1u << 11,
    DAY_HPI = 1u << 12, DAY_PREC = 1u << 13, DAY_TA_GT = 1u << 14, DAY_NOT_LE_HPI = 1u << 15,
    // per-step event flags (adaptive_time_stepping.py:191-201)
    EV_RAIN_WARM = 1u << 16, EV_SNOW_WARM = 1u << 17, EV_NOT_PREC_LE0 = 1u << 18, EV_NOT_SNOWING = 1u << 19,
};

// ------------------------------------------------------------------------------------------------
// numpy-order sums of a 144-slot forcing row (np.add.reduce pairwise summation, 8 lanes, halves of 72)
template <typename F>
__host__ __device__ inline double pairwise144(F at) {
    double half[2];
    for (int h = 0; h < 2; h++) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = at(h * 72 + j);
        for (int i = 8; i < 72; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) r[j] += at(h * 72 + i + j);
        }
        half[h] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    }
    return 0.0 + (half[0] + half[1]);
}

// one cell's 144 forcing slots inside a re-tiled day block ([tile][slot][cell]: consecutive slots are kTile doubles apart)
struct TiledRow {
    const double* p;
    __host__ __device__ double operator[](int k) const { return p[(size_t)k * kTile]; }
};
__host__ __device__ inline TiledRow tiled_row(const double* day, int64_t i) {
    return TiledRow{day + (size_t)(i >> kTileShift) * (144 * kTile) + (size_t)(i & kTileMask)};
}

struct RowAgg {  // aggregates of one forcing row for the three time-step classes
    template <class R>
    __host__ __device__ static double day_sum(R row) {
        return pairwise144([&](int k) { return row[k]; });
    }
    template <class R>
    __host__ __device__ static double day_nanmean(R row) {
        int cnt = 0;
        for (int k = 0; k < 144; k++) cnt += (row[k] == row[k]) ? 1 : 0;
        double tot = pairwise144([&](int k) { double v = row[k]; return (v == v) ? v : 0.0; });
        return tot / (double)cnt;
    }
    // Sum of the six slots itt_day .. itt_day + 5 as numpy's pairwise reduction of the 144-slot row with every other slot
    // zero computes it (adaptive_time_stepping.py:410-437: the hourly aggregates are sums over a masked copy of the whole
    // day).  The reduction keeps 8 running lanes per half of 72 slots, slot k feeding lane k mod 8 of half k / 72; six
    // consecutive slots hit six different lanes, and a lane that sees the value v among zeros ends as v + 0.0 (exact for every
    // v; -0.0 becomes +0.0 exactly as in the full sum).  So six loads and the two lane trees give the same bits as walking all
    // 144 slots.  NAN0: nanmean's zero-filled copy; *cnt receives the number of non-NaN slots.
    template <bool NAN0, class R>
    __host__ __device__ static double window6(R row, int itt_day, int* cnt) {
        double r0[8], r1[8];
        int n_ok = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = itt_day + ((j - itt_day) & 7);  // the one slot of lane j inside an 8-slot span starting at itt_day
            double v = 0.0;
            const bool in = (k < itt_day + 6) && (k < 144);
            if (in) {
                v = row[k];
                if (NAN0) { const bool ok = (v == v); n_ok += ok ? 1 : 0; v = ok ? v : 0.0; }
                v = v + 0.0;
            }
            const bool upper = k >= 72;
            r0[j] = upper ? 0.0 : v;
            r1[j] = upper ? v : 0.0;
        }
        const double h0 = ((r0[0] + r0[1]) + (r0[2] + r0[3])) + ((r0[4] + r0[5]) + (r0[6] + r0[7]));
        const double h1 = ((r1[0] + r1[1]) + (r1[2] + r1[3])) + ((r1[4] + r1[5]) + (r1[6] + r1[7]));
        if (cnt) *cnt = n_ok;
        return 0.0 + (h0 + h1);
    }
    template <class R>
    __host__ __device__ static double win_sum(R row, int itt_day) { return window6<false>(row, itt_day, nullptr); }
    template <class R>
    __host__ __device__ static double win_nanmean(R row, int itt_day) {
        int cnt = 0;
        const double tot = window6<true>(row, itt_day, &cnt);
        return tot / (double)cnt;
    }
    // the straightforward forms (all 144 slots), kept for the self-test below
    template <class R>
    __host__ __device__ static double win_sum_full(R row, int itt_day) {
        return pairwise144([&](int k) { return (k >= itt_day && k < itt_day + 6) ? row[k] : 0.0; });
    }
    template <class R>
    __host__ __device__ static double win_nanmean_full(R row, int itt_day) {
        int cnt = 0;
        for (int k = itt_day; k < itt_day + 6 && k < 144; k++) cnt += (row[k] == row[k]) ? 1 : 0;
        double tot = pairwise144([&](int k) {
            double v = (k >= itt_day && k < itt_day + 6) ? row[k] : 0.0;
            return (v == v) ? v : 0.0;
        });
        return tot / (double)cnt;
    }
};

// ------------------------------------------------------------------------------------------------
struct Planes {
    double* ps;  // n_tiles x kNumPS x kTile
    double* fl;  // n_tiles x kNumF x kTile (may be null)
    int64_t C;
    __host__ __device__ double* at(int id, int64_t i) const {  // plane `id` of cell i
        const size_t t = (size_t)(i >> kTileShift), l = (size_t)(i & kTileMask);
        return kind_is_F(id) ? fl + ((t * kNumF + (size_t)(id - kNumPS)) << kTileShift) + l
                             : ps + ((t * kNumPS + (size_t)ps_slot(id)) << kTileShift) + l;
    }
};

__device__ __forceinline__ void warp_or(unsigned* flags, unsigned bits) {
    unsigned all = __reduce_or_sync(0xffffffffu, bits);
    if ((threadIdx.x & 31) == 0 && all) atomicOr(flags, all);
}

// snow flags of the current state (used before the first step and after a host upload of swe / swe_top)
__global__ void k_snow_flags(Planes pl, unsigned* flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bits = 0;
    if (i < pl.C) {
        double swe = *pl.at(RF_swe, i), st = *pl.at(RF_swe_top, i);
        bits |= (swe > 0) ? FLAG_SWE_GT0 : 0;
        bits |= !(swe <= 0) ? FLAG_SWE_NOT_LE0 : 0;
        bits |= (st > 0) ? FLAG_SWETOP_GT0 : 0;
        bits |= !(st <= 0) ? FLAG_SWETOP_NOT_LE0 : 0;
    }
    warp_or(flags, bits);
}

// grid-wide conditions on the day's forcing block (adaptive_time_stepping.py:38-80), distributed forcing; any order of
// the elements will do, so the re-tiled block is walked linearly (the padding cells of the last tile hold copies of cell 0)
__global__ void k_day_flags(const double* __restrict__ prec_day, const double* __restrict__ ta_day, int64_t n,
                            double ta_fm, double hpi, unsigned* flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bits = 0;
    for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double p = prec_day[i], t = ta_day[i];
        bits |= !(p <= 0) ? DAY_NOT_PREC_LE0 : 0;
        bits |= !(t > ta_fm) ? DAY_NOT_TA_GT : 0;
        bits |= ((p > 0) && (t <= ta_fm)) ? DAY_SNOWFALL : 0;
        bits |= !((p <= 0) && (t <= ta_fm)) ? DAY_NOT_DRY_COLD : 0;
        bits |= (p > hpi) ? DAY_HPI : 0;
        bits |= (p > 0) ? DAY_PREC : 0;
        bits |= (t > ta_fm) ? DAY_TA_GT : 0;
        bits |= !(p <= hpi) ? DAY_NOT_LE_HPI : 0;
    }
    warp_or(flags, bits);
}

// first block of adaptive_time_stepping (:130-189) for per-cell forcing: prec/ta[tau] for the chosen class,
// old prec kept as prec[taum1]; plus the event start/end conditions (:191-201)
__global__ void k_aggregate_a(Planes pl, double* __restrict__ prec_m1, const double* __restrict__ prec_day,
                              const double* __restrict__ ta_day, int class_a, int itt_day, double ta_fm, unsigned* flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bits = 0;
    if (i < pl.C) {
        double* prec = pl.at(RF_prec, i);
        double* ta = pl.at(RF_ta, i);
        double p = *prec, t = *ta;
        prec_m1[i] = p;
        const TiledRow pr = tiled_row(prec_day, i), tr = tiled_row(ta_day, i);
        if (class_a == 1) { p = RowAgg::day_sum(pr); t = RowAgg::day_nanmean(tr); }
        else if (class_a == 2) { p = RowAgg::win_sum(pr, itt_day); t = RowAgg::win_nanmean(tr, itt_day); }
        else if (class_a == 3) { p = pr[itt_day]; t = tr[itt_day]; }
        if (class_a) { *prec = p; *ta = t; }
        double swe = *pl.at(RF_swe, i), st = *pl.at(RF_swe_top, i);
        bits |= ((p > 0) && (t > ta_fm)) ? EV_RAIN_WARM : 0;
        bits |= (((swe > 0) || (st > 0)) && (t > ta_fm)) ? EV_SNOW_WARM : 0;
        bits |= !(p <= 0) ? EV_NOT_PREC_LE0 : 0;
        bits |= !((p > 0) && (t <= ta_fm)) ? EV_NOT_SNOWING : 0;
    }
    warp_or(flags, bits);
}

// The fused step.  One thread = one soil column; all parameters and state in registers.
#ifndef ROGER_MIN_BLOCKS
#define ROGER_MIN_BLOCKS 2
#endif
#ifndef ROGER_BLOCK
#define ROGER_BLOCK 256
#endif
#ifndef ROGER_BULK_PREFETCH
#define ROGER_BULK_PREFETCH 1
#endif
#ifndef ROGER_BULK_PREFETCH_DIST
#define ROGER_BULK_PREFETCH_DIST 0  // tiles ahead of the block's own
#endif
constexpr unsigned kPrefetchChunk = 32 * kTile * sizeof(double);  // 64 KB per instruction
constexpr unsigned kPrefetchChunks = (unsigned)((kTileBytesPS + kPrefetchChunk - 1) / kPrefetchChunk);
template <bool LAT>
__host__ __device__ constexpr unsigned kPrefetchEnd() {  // the head of the tile: the planes a state-only step reads
    return (unsigned)((LAT ? kHotOned : kHotSvat) * kTile * sizeof(double));
}
static_assert(ROGER_BLOCK == kTile, "one k_svat_step block per tile");
template <bool LAT, bool FLUX, bool GWB = false>
__global__ void __launch_bounds__(ROGER_BLOCK, ROGER_MIN_BLOCKS)
k_svat_step(Planes pl, const double* __restrict__ prec_m1_plane, const double* __restrict__ ta_day,
            const double* __restrict__ pet_day, StepPlan p, unsigned* flags) {
    const int64_t C = pl.C;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned bits = 0;
#if ROGER_BULK_PREFETCH
    // The block's tile is one contiguous region: ask the memory system for all of it (or for the tile of the block that
    // will follow on this SM) with a handful of bulk L2 prefetches, instead of discovering the ~100 planes one dependent
    // load group after the other.  No registers, no shared memory, no completion to wait for.
    if (threadIdx.x * kPrefetchChunk < kPrefetchEnd<LAT>()) {
        const int64_t tile = (int64_t)blockIdx.x + ROGER_BULK_PREFETCH_DIST;
        if (tile * kTile < C) {
            const char* src = reinterpret_cast<const char*>(pl.ps) + (size_t)tile * kTileBytesPS + (size_t)threadIdx.x * kPrefetchChunk;
            const unsigned left = (unsigned)(kPrefetchEnd<LAT>() - threadIdx.x * kPrefetchChunk);
            const unsigned bytes = left < kPrefetchChunk ? left : kPrefetchChunk;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
        }
    }
#endif
    if (i < C) {
        Cell c;
#ifdef ROGER_DEBUG_NAN_INIT  // use-before-load detector for the lazy loads (tests only)
        {
            double* raw = reinterpret_cast<double*>(&c);
            for (unsigned k = 0; k < sizeof(Cell) / sizeof(double); k++) raw[k] = __longlong_as_double(0x7ff8000000000000ULL);
        }
#endif
        const Io io{pl.ps, pl.fl, i, p.flux_mask[0], p.flux_mask[1]};
        c.maskCatch = io.ldp(RF_maskCatch);
        c.m = c.maskCatch;
        // ---- adaptive_time_stepping.py:130-189 / :256-360: forcing for this cell ----
        c.prec = io.lds(RF_prec); c.ta = io.lds(RF_ta); c.pet = io.lds(RF_pet);
        if (p.uniform) {
            c.prec_m1 = c.prec;
            if (p.class_a) { c.prec = p.u_prec; c.ta = p.u_ta_a; }
            if (p.class_b) { c.pet = p.u_pet; c.ta = p.u_ta_b; }
        } else {
            c.prec_m1 = prec_m1_plane[i];  // prec / ta already hold the class-A aggregates (k_aggregate_a)
            const TiledRow tr = tiled_row(ta_day, i), er = tiled_row(pet_day, i);
            if (p.class_b == 1) { c.pet = er[p.itt_day]; c.ta = tr[p.itt_day]; }
            else if (p.class_b == 2) { c.pet = RowAgg::win_sum(er, p.itt_day); c.ta = RowAgg::win_nanmean(tr, p.itt_day); }
            else if (p.class_b == 3) { c.pet = RowAgg::day_sum(er); c.ta = RowAgg::day_nanmean(tr); }
        }
        io.sts(RF_prec, c.prec); io.sts(RF_ta, c.ta); io.sts(RF_pet, c.pet);

        const bool ok = (c.m == 1.0) ? cell_step<LAT, FLUX, true, GWB>(c, p, io) : cell_step<LAT, FLUX, false, GWB>(c, p, io);

        bits |= (c.swe > 0) ? FLAG_SWE_GT0 : 0;
        bits |= !(c.swe <= 0) ? FLAG_SWE_NOT_LE0 : 0;
        bits |= (c.swe_top > 0) ? FLAG_SWETOP_GT0 : 0;
        bits |= !(c.swe_top <= 0) ? FLAG_SWETOP_NOT_LE0 : 0;
        bits |= ok ? 0 : FLAG_INSANE;
    }
    warp_or(flags, bits);
}

// Diagnostics accumulators on the device, updated after every step (the reference's `diagnose`):
//   RATE    acc += v                                         roger/diagnostics/rate.py:72-81
//   AVERAGE acc = (isnan(acc) ? 0 : acc) + v, / count later  average.py:71-84, :94-97
//   MAXIMUM acc = nanmax(acc', v), acc' = NaN if acc == 0 and v < 0     maximum.py:76-92 (2-D variables)
//   MINIMUM acc = nanmin(acc', v), acc' = NaN if acc == 0 and v > 0     minimum.py
// k, h of the soil column for the state-only kernels (svat_device.cuh: soil_kh_refresh)
__global__ void k_soil_kh_refresh(Planes pl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.C) return;
    const Io io{pl.ps, pl.fl, i};
    soil_kh_refresh(io);
}

// ---- host-order <-> tile-order movers.  The C ABI speaks plain planes of C doubles (cell index fastest, the caller's
// cell numbering); on the device cell position d holds the caller's cell order[d] (identity while `order` is null, see
// roger_svat_optimize_cell_order).
__global__ void k_scatter_plane(Planes pl, int field, const double* __restrict__ src, const int* __restrict__ order) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < pl.C) *pl.at(field, d) = src[order ? order[d] : d];
}
__global__ void k_gather_plane(Planes pl, int field, double* __restrict__ dst, const int* __restrict__ order) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < pl.C) dst[order ? order[d] : d] = *pl.at(field, d);
}
__global__ void k_gather_linear(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ order, int64_t C) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < C) dst[order ? order[d] : d] = src[d];
}
// Snapshot plane -> page-locked host memory, written by the SMs (16-byte stores over PCIe through the mapped host pointer).
// The copy engine's device-to-host transfer of a snapshot is starved by the DRAM-saturating step kernel running beside it
// (measured on B200: 33 GB/s with the GPU idle, 3 GB/s beside k_svat_step); SM-issued loads take part in the same
// arbitration as the step kernel's own.  Runs on the high-priority copy stream: its blocks slip in as step blocks retire.
constexpr int kD2HBlocks = 16;
__global__ void __launch_bounds__(256) k_copy_to_host(const double2* __restrict__ src, double2* __restrict__ dst, int64_t n2,
                                                       const double* __restrict__ src_tail, double* __restrict__ dst_tail, int tail) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}
__global__ void k_copy_plane(Planes pl, int dst_field, int src_field) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < pl.C) *pl.at(dst_field, i) = *pl.at(src_field, i);
}
// sort key of the cell order: the matrix infiltration capacity scale ks * (1 - sealing) decides, together with the rain
// intensity, whether a cell ponds -- i.e. whether its macropore / shrinkage-crack sub-step loops run or take the
// no-ponding shortcut (svat_device.cuh); masked cells (a different code path altogether) sort to the end
__global__ void k_order_keys(Planes pl, double* __restrict__ keys, int* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.C) return;
    const double m = *pl.at(RF_maskCatch, i);
    double k = *pl.at(RF_ks, i) * (1.0 - *pl.at(RF_sealing, i));
    if (!(k == k)) k = 0.0;
    keys[i] = (m == 1.0) ? k : 1e300;
    ids[i] = (int)i;
}
__global__ void k_compose_order(const int* __restrict__ old_order, const int* __restrict__ sorted_pos, int* __restrict__ out, int64_t C) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < C) out[d] = old_order ? old_order[sorted_pos[d]] : sorted_pos[d];
}
__global__ void k_invert_order(const int* __restrict__ order, int* __restrict__ inv, int64_t C) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < C) inv[order[d]] = (int)d;
}
// per-cell forcing under a cell order: rows [row0, row0 + n) of the caller's block (staged at `src`) go to the device
// positions inv[row]; one warp per row, lanes along the slots (the writes are scattered by construction)
__global__ void __launch_bounds__(256) k_forcing_scatter_rows(const double* __restrict__ src, int64_t row0, int64_t n, double* __restrict__ dst,
                                                               const int* __restrict__ inv) {
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    const int lane = threadIdx.x & 31;
    const int64_t d = inv[row0 + r];
    double* out = dst + (size_t)(d >> kTileShift) * (144 * kTile) + (size_t)(d & kTileMask);
    const double* row = src + (size_t)r * 144;
    for (int k = lane; k < 144; k += 32) out[(size_t)k * kTile] = row[k];
}
__global__ void k_forcing_pad(double* __restrict__ dst, int64_t C, int64_t Cpad) {  // padding cells of the last tile repeat position C - 1
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t npad = Cpad - C;
    if (j >= npad * 144) return;
    const int64_t d = C + j / 144;
    const int k = (int)(j % 144);
    const int64_t s = C - 1;
    dst[(size_t)(d >> kTileShift) * (144 * kTile) + (size_t)k * kTile + (size_t)(d & kTileMask)] =
        dst[(size_t)(s >> kTileShift) * (144 * kTile) + (size_t)k * kTile + (size_t)(s & kTileMask)];
}
// A chunk of a per-cell forcing day block, (cell, slot) order -> [tile][slot][cell].  One block per 32 cells: rows are read
// with the lanes along the slots (contiguous 256 B per request), turned in shared memory and written with the lanes along
// the cells.  `cell0` is a multiple of 32; cells at and beyond C (padding of the last tile) receive cell `C - 1`'s row so
// that linear walks over the re-tiled block (k_day_flags) see no foreign values.
__global__ void __launch_bounds__(256) k_forcing_to_tiles(const double* __restrict__ src, int64_t src_cell0, double* __restrict__ dst,
                                                           int64_t cell0, int64_t C) {
    __shared__ double t[32][145];
    const int64_t base = cell0 + (int64_t)blockIdx.x * 32;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = w; r < 32; r += 8) {
        int64_t cell = base + r;
        if (cell >= C) cell = C - 1;
        const double* row = src + (size_t)(cell - src_cell0) * 144;
        for (int k = lane; k < 144; k += 32) t[r][k] = row[k];
    }
    __syncthreads();
    const int64_t cell = base + lane;
    double* out = dst + (size_t)(cell >> kTileShift) * (144 * kTile) + (size_t)(cell & kTileMask);
    for (int k = w; k < 144; k += 8) out[(size_t)k * kTile] = t[lane][k];
}

__global__ void k_accumulate_rates(Planes pl, double* __restrict__ acc, const int* __restrict__ fields, const int* __restrict__ modes, int n,
                                   double* __restrict__ prev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pl.C) return;
    for (int k = 0; k < n; k++) {
        const double v = *pl.at(fields[k], i);
        double a = acc[(size_t)k * pl.C + i];
        const int mode = modes[k];
        if (mode == ROGER_DIAG_RATE) a = a + v;
        else if (mode == ROGER_DIAG_AVERAGE) a = ((a != a) ? 0.0 : a) + v;
        else if (mode == ROGER_DIAG_MAXIMUM_T || mode == ROGER_DIAG_MINIMUM_T) {
            // variables with a time-step dimension (maximum.py:60-75): the extreme of the last two time levels, not a running one
            const bool mx = (mode == ROGER_DIAG_MAXIMUM_T);
            double p = prev[(size_t)k * pl.C + i];  // [taum1]: the value the previous step left
            prev[(size_t)k * pl.C + i] = v;
            if ((p == 0) && (mx ? (v < 0) : (v > 0))) p = __longlong_as_double(0x7ff8000000000000ULL);
            if (p != p) a = v;
            else if (v == v) a = mx ? ((v > p) ? v : p) : ((v < p) ? v : p);
            else a = p;
        } else {
            const bool mx = (mode == ROGER_DIAG_MAXIMUM);
            if ((a == 0) && (mx ? (v < 0) : (v > 0))) a = __longlong_as_double(0x7ff8000000000000ULL);
            if (a != a) a = v;                    // nanmax / nanmin of (NaN, v) = v (NaN if both are)
            else if (v == v) a = mx ? ((v > a) ? v : a) : ((v < a) ? v : a);
        }
        acc[(size_t)k * pl.C + i] = a;
    }
}

// Several consecutive adaptive steps of every cell in ONE launch (uniform forcing, snow-free stretch of a day: the
// host scheduler knows every step's dt / aggregates in advance, svat_api.cu plan_day_block).  A thread re-reads the
// state it stored in the previous step: at 4 blocks x 128 cells x ~120 planes per SM the whole grid's resident working
// set is ~70 MB and stays in the 126 MB L2, so HBM sees the state once per launch instead of once per step.
#ifndef ROGER_STEPS_MIN_BLOCKS
#define ROGER_STEPS_MIN_BLOCKS 2
#endif
#ifndef ROGER_STEPS_BLOCK
#define ROGER_STEPS_BLOCK 128
#endif
template <bool LAT>
__global__ void __launch_bounds__(ROGER_STEPS_BLOCK, ROGER_STEPS_MIN_BLOCKS)
k_svat_steps(Planes pl, const StepPlan* __restrict__ plans, int n_plans, unsigned* flags) {
    __shared__ StepPlan sp;
    const int64_t C = pl.C;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < n_plans; s++) {
        __syncthreads();
        {
            const unsigned* src = reinterpret_cast<const unsigned*>(plans + s);
            unsigned* dst = reinterpret_cast<unsigned*>(&sp);
            for (int k = threadIdx.x; k < (int)(sizeof(StepPlan) / sizeof(unsigned)); k += blockDim.x) dst[k] = src[k];
        }
        __syncthreads();
        const StepPlan& p = sp;
        unsigned bits = 0;
        if (i < C) {
            Cell c;
            const Io io{pl.ps, pl.fl, i};
            c.maskCatch = io.ldp(RF_maskCatch);
            c.m = c.maskCatch;
            c.prec = io.lds(RF_prec); c.ta = io.lds(RF_ta); c.pet = io.lds(RF_pet);
            c.prec_m1 = c.prec;
            if (p.class_a) { c.prec = p.u_prec; c.ta = p.u_ta_a; }
            if (p.class_b) { c.pet = p.u_pet; c.ta = p.u_ta_b; }
            io.sts(RF_prec, c.prec); io.sts(RF_ta, c.ta); io.sts(RF_pet, c.pet);
            const bool ok = (c.m == 1.0) ? cell_step<LAT, false, true>(c, p, io) : cell_step<LAT, false, false>(c, p, io);
            bits |= (c.swe > 0) ? FLAG_SWE_GT0 : 0;
            bits |= !(c.swe <= 0) ? FLAG_SWE_NOT_LE0 : 0;
            bits |= (c.swe_top > 0) ? FLAG_SWETOP_GT0 : 0;
            bits |= !(c.swe_top <= 0) ? FLAG_SWETOP_NOT_LE0 : 0;
            bits |= ok ? 0 : FLAG_INSANE;
        }
        warp_or(flags + s, bits);
    }
}

}  // namespace roger

// ================================================================================================
// host side
// ================================================================================================
using namespace roger;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return -code;
}
namespace roger {
int set_last_error(int code, const std::string& msg) { return fail(code, msg); }  // shared with sas_api.cu
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ROGER_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    } while (0)

struct roger_svat {
    int64_t C = 0;
    int device = 0;
    roger_svat_settings st{};
    roger_svat_scalars sc{};
    double* d_ps = nullptr;
    double* d_fl = nullptr;
    double* d_prec_m1 = nullptr;
    double* d_forc[3] = {nullptr, nullptr, nullptr};  // prec_day, ta_day, pet_day re-tiled to [tile][slot][cell] (per-cell forcing)
    const double* forc[3] = {nullptr, nullptr, nullptr};
    double* d_stage = nullptr;        // one host-order plane (upload / download) == kForcingChunk cells of a forcing day block
    int* d_order = nullptr;           // device position -> caller's cell index (null: identity), roger_svat_optimize_cell_order
    int* d_inv = nullptr;             // caller's cell index -> device position
    int64_t n_tiles = 0, Cpad = 0;
    double h_forc[3][144];  // uniform forcing (host copy for the scheduler)
    unsigned* d_flags = nullptr;  // [0] snow/sanity flags, [1] day flags, [2] event flags
    unsigned* h_flags = nullptr;  // pinned
    unsigned snow_flags = 0;
    unsigned day_flags = 0;
    bool snow_flags_valid = false;
    bool ks_ss_synced = false;
    bool soil_kh_stale = false;   // the state-only kernels do not evaluate k, h of the soil column: refreshed on access
    bool have_forcing = false;
    // uniform forcing: the (uniform) value the prec / ta planes currently hold, tracked by the scheduler
    double u_cur_prec = 0, u_cur_ta = 0;
    bool u_cur_valid = false;
    // RATE diagnostics accumulated on the device (roger_svat_set_rate_fields)
    std::vector<int> rate_fields, rate_modes;
    bool rates_need_flux = false;
    double* d_rates = nullptr;
    double* d_rates_prev = nullptr;  // [taum1] planes of the MAXIMUM_T / MINIMUM_T accumulators
    int* d_rate_fields = nullptr;  // n field ids followed by n modes
    int64_t diag_count = 0;        // steps accumulated since the last reset (average_nitts)
    // day-blocked stepping (run_block)
    int max_block = 144;
    std::vector<int> last_dts;   // dt_secs of every step of the last roger_svat_step / roger_svat_run_steps call
    int step_batching = 1;   // 0: one launch + flag read-back per step; 1: several steps in ONE launch (k_svat_steps);
                             // 2: one launch per step, queued back to back with a single read-back per batch
    void* d_plans = nullptr;
    void* h_plans = nullptr;
    unsigned* d_block_flags = nullptr;
    unsigned* h_block_flags = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;          // asynchronous diagnostic read-back (roger_svat_download_async), high priority
    int d2h_mode = 1;                            // >= 1: SM copy kernel into mapped host memory where possible (1: default grid,
                                                 // n > 1: n blocks), 0: copy engine
    std::vector<double*> snap_free, snap_busy;   // device snapshot planes
    cudaEvent_t ev_snap = nullptr;
    static constexpr int kFences = 64;           // ring of copy-stream fences (roger_svat_download_fence)
    cudaEvent_t fences[kFences] = {};
    int fence_next = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_kernel_ms = 0;
    int64_t last_launches = 0;
    roger_flag_reduce_fn reduce_fn = nullptr;
    void* reduce_user = nullptr;
    // flag exchange across ranks on the compute stream (roger_svat_enable_nccl_flags)
    void* nccl_comm = nullptr;
    int nccl_world = 1;
    unsigned* d_gather = nullptr;   // world x kMaxFlagWords
    unsigned* h_gather = nullptr;   // pinned, same size
    static constexpr int kMaxFlagWords = 144;
    // in-memory checkpoint (roger_svat_checkpoint / roger_svat_rollback)
    double* d_ckpt_ps = nullptr;
    double* d_ckpt_fl = nullptr;
    struct Saved {
        roger_svat_scalars sc;
        unsigned snow_flags, day_flags;
        bool snow_flags_valid, ks_ss_synced, soil_kh_stale, have_forcing, u_cur_valid;
        double u_cur_prec, u_cur_ta;
        double h_forc[3][144];
        int64_t diag_count;
    } saved{};
    bool have_ckpt = false;
    Planes planes() const { return Planes{d_ps, d_fl, C}; }
    int grid256() const { return (int)((C + 255) / 256); }
};

static int ensure_flux_planes(roger_svat* h) {
    if (h->d_fl) return 0;
    CU(cudaSetDevice(h->device));
    CU(cudaMalloc(&h->d_fl, kTileBytesF * (size_t)h->n_tiles));
    CU(cudaMemsetAsync(h->d_fl, 0, kTileBytesF * (size_t)h->n_tiles, h->stream));
    return 0;
}

// ---- NCCL, bound at run time.  The library links against the CUDA runtime only; the all-gather of the scheduler's flag
// words uses the NCCL the process already carries (torch's bundled libnccl.so.2, else the system one), looked up with dlopen.
namespace {
struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
    static constexpr int kUint32 = 3;  // ncclUint32 (nccl.h: ncclInt8 0, ncclUint8 1, ncclInt32 2, ncclUint32 3)
};
NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (torch), if any
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.why = std::string("libnccl.so.2 not loadable: ") + dlerror(); return api; }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
    if (!api.ok) api.why = "libnccl.so.2 lacks an expected symbol";
    return api;
}
}  // namespace
#define NCCL(call)                                                                                          \
    do {                                                                                                    \
        int r__ = (call);                                                                                   \
        if (r__ != 0)                                                                                       \
            return fail(ROGER_ECUDA, std::string(#call) + ": " + (nccl_api().GetErrorString ? nccl_api().GetErrorString(r__) : "NCCL error")); \
    } while (0)

// n flag words of this rank (device address, final on the compute stream) -> the bitwise OR over all ranks, on the host.
// With NCCL enabled the exchange is one ncclAllGather of n words on the compute stream followed by the same read-back a
// single rank does: no host callback, no second synchronisation.  `done` (optional) is recorded behind the transfers.
static int fetch_flags(roger_svat* h, const unsigned* d_words, int n, uint32_t* out, cudaEvent_t done = nullptr) {
    if (n > roger_svat::kMaxFlagWords) return fail(ROGER_EINVAL, "fetch_flags: too many words");
    if (h->nccl_comm) {
        NCCL(nccl_api().AllGather(d_words, h->d_gather, (size_t)n, NcclApi::kUint32, h->nccl_comm, h->stream));
        CU(cudaMemcpyAsync(h->h_gather, h->d_gather, sizeof(unsigned) * n * h->nccl_world, cudaMemcpyDeviceToHost, h->stream));
        if (done) CU(cudaEventRecord(done, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int k = 0; k < n; k++) {
            uint32_t w = 0;
            for (int r = 0; r < h->nccl_world; r++) w |= h->h_gather[(size_t)r * n + k];
            out[k] = w;
        }
        return 0;
    }
    CU(cudaMemcpyAsync(h->h_gather, d_words, sizeof(unsigned) * n, cudaMemcpyDeviceToHost, h->stream));
    if (done) CU(cudaEventRecord(done, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < n; k++) out[k] = h->h_gather[k];
    if (h->reduce_fn) h->reduce_fn(out, n, h->reduce_user);
    return 0;
}

// k, h of the soil column are evaluated lazily for the state-only kernels: bring the two planes up to date
static int refresh_soil_kh(roger_svat* h) {
    if (!h->soil_kh_stale) return 0;
    k_soil_kh_refresh<<<(int)((h->C + 255) / 256), 256, 0, h->stream>>>(h->planes());
    CU(cudaGetLastError());
    h->soil_kh_stale = false;
    return 0;
}
static bool is_soil_kh(int field) { return field == RF_k || field == RF_h; }

extern "C" {

const char* roger_last_error(void) { return g_err.c_str(); }
int roger_svat_field_count(void) { return RF_COUNT; }
const char* roger_svat_field_name(int f) { return (f >= 0 && f < RF_COUNT) ? kFieldName[f] : nullptr; }
int roger_svat_field_kind(int f) { return (f >= 0 && f < RF_COUNT) ? kFieldKind[f] : -1; }

static int ensure_copy_stream(roger_svat* h);
static int copy_plane_to_host(roger_svat* h, const double* snap, double* host_pinned);

int roger_svat_create(roger_svat** out, int64_t n_cells, const roger_svat_settings* settings, int device) {
    if (!out || !settings || n_cells <= 0) return fail(ROGER_EINVAL, "roger_svat_create: bad argument");
    if (n_cells >= (int64_t(1) << 31)) return fail(ROGER_EINVAL, "roger_svat_create: at most 2^31 - 1 cells per handle; shard the grid");
    if (settings->enable_groundwater_boundary && !(settings->kf_max > 0))
        return fail(ROGER_EINVAL, "roger_svat_create: settings.kf_max must be positive with enable_groundwater_boundary");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ROGER_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                      " (this backend has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ROGER_EINVAL, "roger_svat_create: bad device index");
    CU(cudaSetDevice(device));
    roger_svat* h = new roger_svat();
    h->C = n_cells;
    h->n_tiles = (n_cells + kTile - 1) / kTile;
    h->Cpad = h->n_tiles * kTile;
    h->device = device;
    h->st = *settings;
    h->sc.dt = 1.0;
    h->sc.dt_secs = 3600;
    h->sc.event_id_counter = 1;
    h->sc.sanity_ok = 1;
    h->sc.month = 1;
    // Day-blocked launches pay off where the per-step launch + flag read-back latency (~35 us) is comparable to the kernel
    // time; on large grids the per-step kernel is faster per step than the blocked one.
    h->step_batching = (n_cells <= 262144) ? 1 : 2;
    const size_t plane = sizeof(double) * (size_t)n_cells;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&h->d_ps, kTileBytesPS * (size_t)h->n_tiles));
    CU(cudaMemsetAsync(h->d_ps, 0, kTileBytesPS * (size_t)h->n_tiles, h->stream));
    CU(cudaMalloc(&h->d_prec_m1, plane));
    CU(cudaMalloc(&h->d_stage, plane));
    CU(cudaMalloc(&h->d_flags, 4 * sizeof(unsigned)));
    CU(cudaMemsetAsync(h->d_flags, 0, 4 * sizeof(unsigned), h->stream));
    CU(cudaHostAlloc(&h->h_flags, 4 * sizeof(unsigned), cudaHostAllocDefault));
    CU(cudaMalloc(&h->d_gather, sizeof(unsigned) * roger_svat::kMaxFlagWords));
    CU(cudaHostAlloc(&h->h_gather, sizeof(unsigned) * roger_svat::kMaxFlagWords, cudaHostAllocDefault));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    // everything the asynchronous read-back needs exists from the start: no allocation, stream or event creation ever
    // happens inside a stepping loop (snapshot planes: roger_svat_reserve_snapshots)
    { int r = ensure_copy_stream(h); if (r) { roger_svat_destroy(h); return r; } }
    for (auto& f : h->fences) CU(cudaEventCreateWithFlags(&f, cudaEventDisableTiming | cudaEventBlockingSync));
    *out = h;
    return 0;
}

void roger_svat_destroy(roger_svat* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_ps);
    cudaFree(h->d_fl);
    cudaFree(h->d_prec_m1);
    cudaFree(h->d_stage);
    cudaFree(h->d_order); cudaFree(h->d_inv);
    cudaFree(h->d_ckpt_ps); cudaFree(h->d_ckpt_fl);
    for (auto p : h->d_forc) cudaFree(p);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    for (auto p : h->snap_free) cudaFree(p);
    for (auto p : h->snap_busy) cudaFree(p);
    if (h->ev_snap) cudaEventDestroy(h->ev_snap);
    for (auto& e : h->fences) if (e) cudaEventDestroy(e);
    cudaFree(h->d_rates); cudaFree(h->d_rates_prev); cudaFree(h->d_rate_fields);
    cudaFree(h->d_plans); cudaFree(h->d_block_flags);
    if (h->h_plans) cudaFreeHost(h->h_plans);
    if (h->h_block_flags) cudaFreeHost(h->h_block_flags);
    cudaFree(h->d_flags);
    cudaFreeHost(h->h_flags);
    if (h->nccl_comm) nccl_api().CommDestroy(h->nccl_comm);
    cudaFree(h->d_gather);
    if (h->h_gather) cudaFreeHost(h->h_gather);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
}

int roger_svat_upload(roger_svat* h, int field, const double* host) {
    if (!h || !host || field < 0 || field >= RF_COUNT) return fail(ROGER_EINVAL, "roger_svat_upload: bad argument");
    if (field == RF_k || field == RF_h) h->soil_kh_stale = false;  // the caller's values stand until the next state-only step
    CU(cudaSetDevice(h->device));
    if (kind_is_F(field)) { int r = ensure_flux_planes(h); if (r) return r; }
    CU(cudaMemcpyAsync(h->d_stage, host, sizeof(double) * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
    k_scatter_plane<<<h->grid256(), 256, 0, h->stream>>>(h->planes(), field, h->d_stage, h->d_order);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    if (field == RF_swe || field == RF_swe_top) h->snow_flags_valid = false;
    if (field == RF_prec || field == RF_ta) h->u_cur_valid = false;
    if (field == RF_ks_ss) h->ks_ss_synced = (h->sc.itt > 0) ? false : h->ks_ss_synced;
    return 0;
}

int roger_svat_download(roger_svat* h, int field, double* host) {
    if (!h || !host || field < 0 || field >= RF_COUNT) return fail(ROGER_EINVAL, "roger_svat_download: bad argument");
    CU(cudaSetDevice(h->device));
    if (kind_is_F(field)) { int r = ensure_flux_planes(h); if (r) return r; }
    if (is_soil_kh(field)) { int r = refresh_soil_kh(h); if (r) return r; }
    k_gather_plane<<<h->grid256(), 256, 0, h->stream>>>(h->planes(), field, h->d_stage, h->d_order);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host, h->d_stage, sizeof(double) * (size_t)h->C, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// a snapshot plane from the pool (roger_svat_reserve_snapshots fills it ahead of time; grows on demand otherwise)
static int take_snapshot_plane(roger_svat* h, double** snap) {
    if (!h->snap_free.empty()) { *snap = h->snap_free.back(); h->snap_free.pop_back(); }
    else CU(cudaMalloc(snap, sizeof(double) * (size_t)h->C));
    h->snap_busy.push_back(*snap);
    return 0;
}

// Makes sure `n` snapshot planes exist, so that asynchronous read-backs inside a stepping loop never allocate: a driver
// calls this once when it configures its output (n = planes handed over per output time).
int roger_svat_reserve_snapshots(roger_svat* h, int n) {
    if (!h || n < 0 || n > 4096) return fail(ROGER_EINVAL, "roger_svat_reserve_snapshots: bad argument");
    CU(cudaSetDevice(h->device));
    while ((int)(h->snap_free.size() + h->snap_busy.size()) < n) {
        double* snap = nullptr;
        CU(cudaMalloc(&snap, sizeof(double) * (size_t)h->C));
        h->snap_free.push_back(snap);
    }
    return 0;
}

// Snapshot of a plane as of the steps issued so far, copied to the host off the critical path: a gather into a host-order
// snapshot plane on the compute stream (microseconds), then the PCIe transfer on a second stream while later steps run.
int roger_svat_download_async(roger_svat* h, int field, double* host_pinned) {
    if (!h || !host_pinned || field < 0 || field >= RF_COUNT) return fail(ROGER_EINVAL, "roger_svat_download_async: bad argument");
    CU(cudaSetDevice(h->device));
    if (kind_is_F(field)) { int r = ensure_flux_planes(h); if (r) return r; }
    if (is_soil_kh(field)) { int r = refresh_soil_kh(h); if (r) return r; }
    const size_t plane = sizeof(double) * (size_t)h->C;
    double* snap = nullptr;
    { int r = take_snapshot_plane(h, &snap); if (r) return r; }
    k_gather_plane<<<h->grid256(), 256, 0, h->stream>>>(h->planes(), field, snap, h->d_order);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev_snap, h->stream));
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_snap, 0));
    return copy_plane_to_host(h, snap, host_pinned);
}

// snapshot plane (device) -> caller's page-locked buffer, on the copy stream
static int copy_plane_to_host(roger_svat* h, const double* snap, double* host_pinned) {
    const size_t plane = sizeof(double) * (size_t)h->C;
    void* mapped = nullptr;
    if (h->d2h_mode >= 1 && ((uintptr_t)host_pinned % 16 == 0) && cudaHostGetDevicePointer(&mapped, host_pinned, 0) == cudaSuccess && mapped) {
        const int64_t n2 = h->C / 2;
        const int tail = (int)(h->C - 2 * n2);
        // A FEW blocks only: two step blocks fill an SM's register file, so every SM that hosts a copy block runs the step at
        // half rate for as long as the PCIe writes take; a dozen SMs issue more stores than the link absorbs.
        const int blocks = (h->d2h_mode == 1) ? kD2HBlocks : h->d2h_mode;
        k_copy_to_host<<<blocks, 256, 0, h->copy_stream>>>(reinterpret_cast<const double2*>(snap), reinterpret_cast<double2*>(mapped), n2,
                                                             snap + 2 * n2, reinterpret_cast<double*>(mapped) + 2 * n2, tail);
        CU(cudaGetLastError());
        return 0;
    }
    cudaGetLastError();  // (a buffer that is not mapped: clear the sticky-free error of the query and use the copy engine)
    CU(cudaMemcpyAsync(host_pinned, snap, plane, cudaMemcpyDeviceToHost, h->copy_stream));
    return 0;
}

static int ensure_copy_stream(roger_svat* h) {
    if (!h->copy_stream) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = greatest priority
        CU(cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
    }
    return 0;
}

int roger_svat_download_diagnostics_async(roger_svat* h, double* host_pinned, int reset, int64_t* count) {
    if (!h || !host_pinned) return fail(ROGER_EINVAL, "roger_svat_download_diagnostics_async: bad argument");
    if (h->rate_fields.empty()) return fail(ROGER_EINVAL, "roger_svat_download_diagnostics_async: no diagnostics set");
    CU(cudaSetDevice(h->device));
    const size_t plane = sizeof(double) * (size_t)h->C;
    const size_t n = h->rate_fields.size();
    std::vector<double*> snaps(n);
    for (size_t k = 0; k < n; k++) {
        double* snap = nullptr;
        { int r = take_snapshot_plane(h, &snap); if (r) return r; }
        snaps[k] = snap;
        k_gather_linear<<<h->grid256(), 256, 0, h->stream>>>(h->d_rates + k * (size_t)h->C, snap, h->d_order, h->C);
        CU(cudaGetLastError());
    }
    if (reset) CU(cudaMemsetAsync(h->d_rates, 0, plane * n, h->stream));
    CU(cudaEventRecord(h->ev_snap, h->stream));
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_snap, 0));
    for (size_t k = 0; k < n; k++) { int r = copy_plane_to_host(h, snaps[k], host_pinned + k * (size_t)h->C); if (r) return r; }
    if (count) *count = h->diag_count;
    if (reset) h->diag_count = 0;
    return 0;
}

int roger_svat_download_fence(roger_svat* h) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_download_fence: null handle");
    CU(cudaSetDevice(h->device));
    const int t = h->fence_next;
    h->fence_next = (h->fence_next + 1) % roger_svat::kFences;
    CU(cudaEventRecord(h->fences[t], h->copy_stream));
    return t;
}

int roger_svat_fence_wait(roger_svat* h, int token) {
    if (!h || token < 0 || token >= roger_svat::kFences || !h->fences[token]) return fail(ROGER_EINVAL, "roger_svat_fence_wait: bad token");
    CU(cudaEventSynchronize(h->fences[token]));
    return 0;
}

int roger_svat_set_d2h_mode(roger_svat* h, int mode) {
    if (!h || mode < 0 || mode > 65535) return fail(ROGER_EINVAL, "roger_svat_set_d2h_mode: 0 (copy engine), 1 (SM copy kernel, default grid) or the number of copy blocks");
    h->d2h_mode = mode;
    return 0;
}

int roger_svat_download_wait(roger_svat* h) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_download_wait: null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->copy_stream));
    for (auto p : h->snap_busy) h->snap_free.push_back(p);
    h->snap_busy.clear();
    return 0;
}

int roger_svat_field_slot(int field) { return (field >= 0 && field < kNumPS) ? ps_slot(field) : -1; }

int roger_svat_tile_layout(roger_svat* h, roger_svat_layout* out) {
    if (!h || !out) return fail(ROGER_EINVAL, "roger_svat_tile_layout: bad argument");
    { int r = refresh_soil_kh(h); if (r) return r; }  // k, h current as of this call; later state-only steps do not update them
    out->ps_base = (uint64_t)(uintptr_t)h->d_ps;
    out->flux_base = (uint64_t)(uintptr_t)h->d_fl;
    out->tile_cells = kTile;
    out->n_tiles = h->n_tiles;
    out->n_ps_planes = kNumPS;
    out->n_flux_planes = kNumF;
    return 0;
}

int roger_svat_set_scalars(roger_svat* h, const roger_svat_scalars* s) {
    if (!h || !s) return fail(ROGER_EINVAL, "roger_svat_set_scalars: bad argument");
    h->sc = *s;
    return 0;
}
int roger_svat_get_scalars(roger_svat* h, roger_svat_scalars* s) {
    if (!h || !s) return fail(ROGER_EINVAL, "roger_svat_get_scalars: bad argument");
    *s = h->sc;
    return 0;
}

static unsigned day_flags_host(const double* pr, const double* tr, int n, double ta_fm, double hpi) {
    unsigned bits = 0;
    for (int k = 0; k < n; k++) {
        double p = pr[k], t = tr[k];
        bits |= !(p <= 0) ? DAY_NOT_PREC_LE0 : 0;
        bits |= !(t > ta_fm) ? DAY_NOT_TA_GT : 0;
        bits |= ((p > 0) && (t <= ta_fm)) ? DAY_SNOWFALL : 0;
        bits |= !((p <= 0) && (t <= ta_fm)) ? DAY_NOT_DRY_COLD : 0;
        bits |= (p > hpi) ? DAY_HPI : 0;
        bits |= (p > 0) ? DAY_PREC : 0;
        bits |= (t > ta_fm) ? DAY_TA_GT : 0;
        bits |= !(p <= hpi) ? DAY_NOT_LE_HPI : 0;
    }
    return bits;
}

static int after_forcing(roger_svat* h) {
    if (h->st.forcing_uniform) {
        h->day_flags = day_flags_host(h->h_forc[0], h->h_forc[1], 144, h->st.ta_fm, h->st.hpi);
    } else {
        CU(cudaMemsetAsync(h->d_flags + 1, 0, sizeof(unsigned), h->stream));
        int64_t n = h->Cpad * 144;  // the re-tiled block; its padding cells repeat the last cell
        int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
        k_day_flags<<<blocks, 256, 0, h->stream>>>(h->forc[0], h->forc[1], n, h->st.ta_fm, h->st.hpi, h->d_flags + 1);
        CU(cudaGetLastError());
        uint32_t w = 0;
        { int r = fetch_flags(h, h->d_flags + 1, 1, &w); if (r) return r; }
        h->day_flags = w;
    }
    if (h->st.forcing_uniform && (h->nccl_comm || h->reduce_fn)) {  // every rank may hold its own "uniform" series
        h->h_flags[1] = h->day_flags;
        CU(cudaMemcpyAsync(h->d_flags + 1, h->h_flags + 1, sizeof(unsigned), cudaMemcpyHostToDevice, h->stream));
        uint32_t w = 0;
        { int r = fetch_flags(h, h->d_flags + 1, 1, &w); if (r) return r; }
        h->day_flags = w;
    }
    h->have_forcing = true;
    return 0;
}

// Per-cell forcing: (cell, slot) day blocks are re-tiled to [tile][slot][cell] in chunks of kForcingChunk cells through the
// staging plane (host source) or straight from the caller's device buffer.
static int retile_forcing(roger_svat* h, int v, const double* src, bool src_on_host) {
    if (!h->d_forc[v]) CU(cudaMalloc(&h->d_forc[v], sizeof(double) * 144 * (size_t)h->Cpad));
    const int64_t chunk = std::max<int64_t>(32, (h->C / 144) & ~int64_t(31));  // cells per staging pass (the plane holds C doubles)
    auto place = [&](const double* dev_rows, int64_t row0, int64_t n, bool last) -> int {  // rows [row0, row0 + n) -> tiles
        if (h->d_order) {
            k_forcing_scatter_rows<<<(int)((n + 7) / 8), 256, 0, h->stream>>>(dev_rows, row0, n, h->d_forc[v], h->d_inv);
            CU(cudaGetLastError());
            if (last && h->Cpad > h->C) {
                k_forcing_pad<<<(int)(((h->Cpad - h->C) * 144 + 255) / 256), 256, 0, h->stream>>>(h->d_forc[v], h->C, h->Cpad);
                CU(cudaGetLastError());
            }
        } else {
            const int64_t upto = last ? h->Cpad : row0 + n;  // the last pass also fills the padding cells of the last tile
            k_forcing_to_tiles<<<(int)((upto - row0 + 31) / 32), 256, 0, h->stream>>>(dev_rows, row0, h->d_forc[v], row0, h->C);
            CU(cudaGetLastError());
        }
        return 0;
    };
    if (!src_on_host) {
        { int r = place(src, 0, h->C, true); if (r) return r; }
    } else if (h->C < 32 * 144) {  // tiny grids: the staging plane is smaller than one 32-cell group of rows
        double* tmp = nullptr;
        CU(cudaMalloc(&tmp, sizeof(double) * 144 * (size_t)h->C));
        CU(cudaMemcpyAsync(tmp, src, sizeof(double) * 144 * (size_t)h->C, cudaMemcpyHostToDevice, h->stream));
        { int r = place(tmp, 0, h->C, true); if (r) { cudaFree(tmp); return r; } }
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaFree(tmp));
    } else {
        for (int64_t c0 = 0; c0 < h->C; c0 += chunk) {
            const int64_t n = std::min(chunk, h->C - c0);
            CU(cudaMemcpyAsync(h->d_stage, src + (size_t)c0 * 144, sizeof(double) * 144 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
            { int r = place(h->d_stage, c0, n, c0 + n >= h->C); if (r) return r; }
        }
    }
    h->forc[v] = h->d_forc[v];
    return 0;
}

int roger_svat_set_forcing_day(roger_svat* h, const double* prec_day, const double* ta_day, const double* pet_day) {
    if (!h || !prec_day || !ta_day || !pet_day) return fail(ROGER_EINVAL, "roger_svat_set_forcing_day: bad argument");
    CU(cudaSetDevice(h->device));
    const double* src[3] = {prec_day, ta_day, pet_day};
    if (h->st.forcing_uniform) {
        for (int v = 0; v < 3; v++) memcpy(h->h_forc[v], src[v], sizeof(double) * 144);
    } else {
        for (int v = 0; v < 3; v++) { int r = retile_forcing(h, v, src[v], true); if (r) return r; }
    }
    return after_forcing(h);
}

int roger_svat_set_forcing_day_device(roger_svat* h, uint64_t prec_day, uint64_t ta_day, uint64_t pet_day) {
    if (!h || !prec_day || !ta_day || !pet_day) return fail(ROGER_EINVAL, "roger_svat_set_forcing_day_device: bad argument");
    CU(cudaSetDevice(h->device));
    const double* src[3] = {(const double*)(uintptr_t)prec_day, (const double*)(uintptr_t)ta_day, (const double*)(uintptr_t)pet_day};
    if (h->st.forcing_uniform) {
        for (int v = 0; v < 3; v++)
            CU(cudaMemcpyAsync(h->h_forc[v], src[v], sizeof(double) * 144, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    } else {
        for (int v = 0; v < 3; v++) { int r = retile_forcing(h, v, src[v], false); if (r) return r; }
    }
    return after_forcing(h);
}

// ---- host scheduler: the grid-wide part of adaptive_time_stepping_kernel (adaptive_time_stepping.py:34-381) evaluated on
// reduction flags.  Split in two halves around the event conditions (:191-201), which need prec / ta [tau] of every cell.
struct SchedA {  // what the second half needs from the first
    bool any_swe, any_swetop, all_swe_le0;
    int itt_day;
};

static int schedule_first_half(roger_svat* h, unsigned sf, roger_svat_scalars& sc, StepPlan& p, SchedA& a) {
    const unsigned df = h->day_flags;
    const bool all_prec_le0 = !(df & DAY_NOT_PREC_LE0), all_ta_gt = !(df & DAY_NOT_TA_GT);
    const bool any_snowfall = df & DAY_SNOWFALL, all_dry_cold = !(df & DAY_NOT_DRY_COLD);
    const bool any_hpi = df & DAY_HPI, any_prec = df & DAY_PREC, any_ta_gt = df & DAY_TA_GT, all_le_hpi = !(df & DAY_NOT_LE_HPI);
    a.any_swe = sf & FLAG_SWE_GT0;
    a.any_swetop = sf & FLAG_SWETOP_GT0;
    a.all_swe_le0 = !(sf & FLAG_SWE_NOT_LE0);
    const bool all_swetop_le0 = !(sf & FLAG_SWETOP_NOT_LE0);
    const bool cond0 = all_prec_le0 && a.all_swe_le0 && all_swetop_le0 && all_ta_gt;
    const bool cond00 = any_snowfall || all_dry_cold;
    const bool cond1 = any_hpi && any_prec && any_ta_gt;
    const bool cond2 = all_le_hpi && any_prec && any_ta_gt;
    const bool snow_warm = (a.any_swe || a.any_swetop) && any_ta_gt;
    const bool cond3 = any_hpi && any_prec && snow_warm;
    const bool cond4 = all_le_hpi && any_prec && snow_warm;
    const bool cond5 = all_prec_le0 && snow_warm;
    const bool daily = cond0 || cond00;
    const bool hourly = (cond2 || cond4 || cond5) && !cond1 && !cond3;
    const bool tenmin = (cond1 || cond3) && !cond2 && !cond4 && !cond5;
    a.itt_day = (int)sc.itt_day;
    if (a.itt_day < 0 || a.itt_day > 143)
        return fail(ROGER_EINVAL, "roger_svat_step: itt_day outside the forcing day (set_forcing_day not called at midnight?)");
    p = StepPlan{};
    p.class_a = tenmin ? 3 : (hourly ? 2 : (daily ? 1 : 0));
    sc.dt_secs = (sc.time % 86400 == 0) ? 86400 : 3600;
    if (hourly) sc.dt_secs = 3600;
    if (tenmin) sc.dt_secs = 600;
    return 0;
}

// uniform forcing: prec / ta [tau] after the first aggregation block are known on the host
static void uniform_first_aggregates(roger_svat* h, StepPlan& p, int itt_day) {
    const double* pr = h->h_forc[0];
    const double* tr = h->h_forc[1];
    double pa = h->u_cur_prec, ta_a = h->u_cur_ta;
    if (p.class_a == 1) { pa = RowAgg::day_sum(pr); ta_a = RowAgg::day_nanmean(tr); }
    else if (p.class_a == 2) { pa = RowAgg::win_sum(pr, itt_day); ta_a = RowAgg::win_nanmean(tr, itt_day); }
    else if (p.class_a == 3) { pa = pr[itt_day]; ta_a = tr[itt_day]; }
    p.u_prec = pa;
    p.u_ta_a = ta_a;
}

static unsigned uniform_event_flags(const roger_svat_settings& st, const StepPlan& p, const SchedA& a) {
    unsigned ev = 0;
    ev |= ((p.u_prec > 0) && (p.u_ta_a > st.ta_fm)) ? EV_RAIN_WARM : 0;
    ev |= ((a.any_swe || a.any_swetop) && (p.u_ta_a > st.ta_fm)) ? EV_SNOW_WARM : 0;
    ev |= !(p.u_prec <= 0) ? EV_NOT_PREC_LE0 : 0;
    ev |= !((p.u_prec > 0) && (p.u_ta_a <= st.ta_fm)) ? EV_NOT_SNOWING : 0;
    return ev;
}

static void schedule_second_half(roger_svat* h, unsigned ev, const SchedA& a, roger_svat_scalars& sc, StepPlan& p) {
    const roger_svat_settings& st = h->st;
    const int itt_day = a.itt_day;
    const bool cond_event1 = (ev & EV_RAIN_WARM) || (ev & EV_SNOW_WARM);
    // swe[taum1] == swe[tau] at this point (after_timestep copied it): any(swe_m1 > 0) & all(swe <= 0)
    const bool cond_event2 = !(ev & EV_NOT_PREC_LE0) || !(ev & EV_NOT_SNOWING) || (a.any_swe && a.all_swe_le0);
    if (cond_event1) sc.time_event0 = 0;
    if (cond_event2) sc.time_event0 = sc.time_event0 + sc.dt_secs;

    const int64_t end_event = (int64_t)st.end_event;
    const bool c6 = (sc.time_event0 <= end_event) && (sc.dt_secs == 600);
    const bool c7 = (sc.time_event0 <= end_event) && (sc.dt_secs == 3600);
    const bool c8 = (sc.time_event0 <= end_event) && (sc.dt_secs == 86400);
    const bool c9 = (sc.time_event0 > end_event) && (sc.time % 3600 != 0) && (sc.dt_secs == 600);
    const bool c10 = (sc.time_event0 > end_event) && (sc.time % 3600 == 0) && ((sc.dt_secs == 600) || (sc.dt_secs == 3600));
    const bool c11 = (sc.time_event0 > end_event) && (sc.time % 86400 == 0) && (sc.dt_secs == 86400);
    if (c6) { sc.event_id[1] = sc.event_id_counter; sc.dt = 1.0 / 6; sc.itt_day += 1; p.class_b = 1; }
    if (c7) { sc.event_id[1] = sc.event_id_counter; sc.dt = 1; sc.itt_day += 6; p.class_b = 2; }
    if (c8) { sc.dt = 24; sc.itt_day = 0; p.class_b = 3; }
    if (c9) { sc.event_id[1] = 0; sc.dt = 1.0 / 6; sc.dt_secs = 600; sc.itt_day += 1; p.class_b = 1; }
    if (c10) { sc.event_id[1] = 0; sc.dt = 1; sc.dt_secs = 3600; sc.itt_day += 6; p.class_b = 2; }
    if (c11) { sc.event_id[1] = 0; sc.dt = 24; sc.dt_secs = 86400; sc.itt_day = 0; p.class_b = 3; }
    if ((sc.event_id[0] > 0) && (sc.event_id[1] == 0)) sc.event_id_counter += 1;

    if (st.forcing_uniform) {
        const double* tr = h->h_forc[1];
        const double* er = h->h_forc[2];
        if (p.class_b == 1) { p.u_pet = er[itt_day]; p.u_ta_b = tr[itt_day]; }
        else if (p.class_b == 2) { p.u_pet = RowAgg::win_sum(er, itt_day); p.u_ta_b = RowAgg::win_nanmean(tr, itt_day); }
        else if (p.class_b == 3) { p.u_pet = RowAgg::day_sum(er); p.u_ta_b = RowAgg::day_nanmean(tr); }
    }
    p.dt = sc.dt;
    p.ta_fm = st.ta_fm; p.sf = st.sf; p.retention = (10000. / (100 - st.rmax) / 100.);
    p.r_mp = st.r_mp; p.r_mp2 = st.r_mp * st.r_mp; p.r_mp3_6 = std::pow(st.r_mp, 3.0) / 6.0;
    p.l_sc = st.l_sc; p.pi = st.pi; p.transp_water_stress = st.transp_water_stress; p.dx = st.dx;
    p.atol = st.atol; p.rtol = st.rtol;
    p.sqrt6x2 = std::pow(6.0, 0.5) * 2;
    p.itt_day = itt_day;
    p.substeps = (int)std::nearbyint(sc.dt / (1.0 / 5));
    p.cond1 = (sc.event_id[0] == 0) && (sc.event_id[1] >= 1);
    p.cond2s = (sc.event_id[0] >= 1);
    p.cond3s = (sc.event_id[0] == sc.event_id[1]);
    p.cond4 = (sc.event_id[0] >= 1) && (sc.event_id[1] == 0);
    p.cond5 = (sc.event_id[1] >= 1);
    p.irrigation_month = (sc.month >= 4 && sc.month <= 9);
    p.uniform = st.forcing_uniform;
    p.lbc = st.enable_macropore_lower_boundary_condition;
    p.clamp = st.clamp_small_negative_storage;
    p.kf_max = st.kf_max;
}

// bookkeeping at the end of a step (roger.py:447-448 and the after_timestep roll)
static void finish_step_scalars(roger_svat* h, const StepPlan& p, roger_svat_scalars& sc) {
    sc.itt += 1;
    sc.time += sc.dt_secs;
    sc.event_id[0] = sc.event_id[1];
    if (h->st.forcing_uniform) {  // the (uniform) values the prec / ta planes hold now
        if (p.class_a) { h->u_cur_prec = p.u_prec; h->u_cur_ta = p.u_ta_a; }
        if (p.class_b) h->u_cur_ta = p.u_ta_b;
    }
}

static int ensure_snow_flags(roger_svat* h, int64_t* launches) {
    if (h->snow_flags_valid) return 0;
    CU(cudaMemsetAsync(h->d_flags, 0, sizeof(unsigned), h->stream));
    k_snow_flags<<<(int)((h->C + 255) / 256), 256, 0, h->stream>>>(h->planes(), h->d_flags);
    CU(cudaGetLastError());
    (*launches)++;
    uint32_t w = 0;
    { int r = fetch_flags(h, h->d_flags, 1, &w); if (r) return r; }
    h->snow_flags = w;
    h->snow_flags_valid = true;
    return 0;
}

static int ensure_uniform_current(roger_svat* h) {  // the "keep" class reads prec / ta [tau] as they are
    if (!h->st.forcing_uniform || h->u_cur_valid) return 0;
    CU(cudaMemcpyAsync(&h->u_cur_prec, h->planes().at(RF_prec, 0), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&h->u_cur_ta, h->planes().at(RF_ta, 0), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->u_cur_valid = true;
    return 0;
}

// One adaptive time step: host scheduler + one launch of the fused per-cell kernel.
// which flux planes a flux-storing step writes: all of them when the caller asked for fluxes (parity / diagnostics mode), only the
// accumulated ones when a diagnostics accumulator is the reason the flux kernel runs
static void set_flux_mask(const roger_svat* h, bool caller_wants_all, StepPlan& p) {
    p.flux_mask[0] = p.flux_mask[1] = ~0ULL;
    if (caller_wants_all) return;
    p.flux_mask[0] = p.flux_mask[1] = 0;
    for (int f : h->rate_fields)
        if (kind_is_F(f)) p.flux_mask[(f - kNumPS) >> 6] |= 1ULL << ((f - kNumPS) & 63);
}

static int step_impl(roger_svat* h, int write_fluxes) {
    if (!h->have_forcing) return fail(ROGER_EINVAL, "roger_svat_step: no forcing day set");
    CU(cudaSetDevice(h->device));
    const bool caller_wants_fluxes = write_fluxes != 0;
    if (h->rates_need_flux) write_fluxes = 1;  // flux rates are accumulated from the flux planes
    if (write_fluxes) { int r = ensure_flux_planes(h); if (r) return r; }
    const roger_svat_settings& st = h->st;
    roger_svat_scalars& sc = h->sc;
    const int threads = ROGER_BLOCK;
    const int blocks = (int)((h->C + threads - 1) / threads);
    int64_t launches = 0;
    CU(cudaEventRecord(h->ev0, h->stream));
    { int r = ensure_snow_flags(h, &launches); if (r) return r; }
    { int r = ensure_uniform_current(h); if (r) return r; }

    StepPlan p;
    SchedA a;
    { int r = schedule_first_half(h, h->snow_flags, sc, p, a); if (r) return r; }
    unsigned ev = 0;
    if (st.forcing_uniform) {
        uniform_first_aggregates(h, p, a.itt_day);
        ev = uniform_event_flags(st, p, a);
    } else {
        CU(cudaMemsetAsync(h->d_flags + 2, 0, sizeof(unsigned), h->stream));
        k_aggregate_a<<<blocks, threads, 0, h->stream>>>(h->planes(), h->d_prec_m1, h->forc[0], h->forc[1], p.class_a, a.itt_day,
                                                          st.ta_fm, h->d_flags + 2);
        CU(cudaGetLastError());
        launches++;
        uint32_t w = 0;
        { int r = fetch_flags(h, h->d_flags + 2, 1, &w); if (r) return r; }
        ev = w;
    }
    schedule_second_half(h, ev, a, sc, p);
    set_flux_mask(h, caller_wants_fluxes, p);

    CU(cudaMemsetAsync(h->d_flags, 0, sizeof(unsigned), h->stream));
    const Planes pl = h->planes();
#define LAUNCH(LAT, FLUX, GWB) \
    k_svat_step<LAT, FLUX, GWB><<<blocks, threads, 0, h->stream>>>(pl, h->d_prec_m1, h->forc[1], h->forc[2], p, h->d_flags)
    if (st.enable_groundwater_boundary) {  // the water table as a boundary condition: capillary rise into the subsoil, recharge
        if (st.enable_lateral_flow) { if (write_fluxes) LAUNCH(true, true, true); else LAUNCH(true, false, true); }
        else { if (write_fluxes) LAUNCH(false, true, true); else LAUNCH(false, false, true); }
    } else if (st.enable_lateral_flow) { if (write_fluxes) LAUNCH(true, true, false); else LAUNCH(true, false, false); }
    else { if (write_fluxes) LAUNCH(false, true, false); else LAUNCH(false, false, false); }
#undef LAUNCH
    h->soil_kh_stale = !write_fluxes;
    CU(cudaGetLastError());
    launches++;
    if (!h->rate_fields.empty()) {
        const int n = (int)h->rate_fields.size();
        k_accumulate_rates<<<(int)((h->C + 255) / 256), 256, 0, h->stream>>>(pl, h->d_rates, h->d_rate_fields, h->d_rate_fields + n, n, h->d_rates_prev);
        CU(cudaGetLastError());
        launches++;
        h->diag_count++;
    }
    if (!h->ks_ss_synced) {  // subsoil.calc_ks (subsoil.py:6-20): ks_ss = ks from the end of the first step on
        k_copy_plane<<<h->grid256(), 256, 0, h->stream>>>(pl, RF_ks_ss, RF_ks);
        CU(cudaGetLastError());
        launches++;
        h->ks_ss_synced = true;
    }
    {
        uint32_t w = 0;
        int r = fetch_flags(h, h->d_flags, 1, &w, h->ev1);
        if (r) return r;
        h->snow_flags = w;
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_kernel_ms += ms;
    h->last_launches += launches;

    finish_step_scalars(h, p, sc);
    sc.sanity_ok = (h->snow_flags & FLAG_INSANE) ? 0 : 1;
    h->last_dts.push_back((int)sc.dt_secs);
    return 0;
}

// Day-blocked stepping.  With spatially uniform forcing every input of the scheduler except the four snow flags is known
// on the host.  While no cell holds snow and every step's air temperature is above the freeze / melt threshold, no snow
// can accrue (interception.py:163, snow.py:14), the flags stay clear, and the schedule of the coming steps -- dt,
// aggregates, event bookkeeping -- can be laid out in advance and executed by ONE launch (k_svat_steps).  Returns the
// number of steps executed through *n_done (0: the stretch does not qualify, the caller steps one at a time).
// Steps whose scheduler inputs are all known on the host (spatially uniform forcing, no snow anywhere, every step warmer than
// ta_fm so that none can accrue) are laid out in advance and executed without a host round trip in between: `queued` = false
// runs them in ONE launch of k_svat_steps (the state comes back from L2: right for catchment-scale grids), `queued` = true
// launches the ordinary step kernel once per step, back to back, and reads all flag words back once -- no launch gap, no
// read-back latency and, on several GPUs, one flag exchange per batch instead of one per step.
static int run_block(roger_svat* h, int max_steps, int* n_done, bool queued, int write_fluxes, bool caller_wants_fluxes) {
    *n_done = 0;
    if (h->st.enable_groundwater_boundary) return 0;  // single steps only (the batched kernels are built without that variant)
    const roger_svat_settings& st = h->st;
    if (!st.forcing_uniform || !h->ks_ss_synced || !h->have_forcing || max_steps < 2) return 0;
    CU(cudaSetDevice(h->device));
    int64_t launches = 0;
    { int r = ensure_snow_flags(h, &launches); if (r) return r; }
    { int r = ensure_uniform_current(h); if (r) return r; }
    const unsigned snow_bits = FLAG_SWE_GT0 | FLAG_SWE_NOT_LE0 | FLAG_SWETOP_GT0 | FLAG_SWETOP_NOT_LE0;
    if (h->snow_flags & snow_bits) return 0;
    const unsigned sf = 0;  // no snow anywhere, and it stays that way while every step is warm

    std::vector<StepPlan> plans;
    std::vector<int> dts;
    roger_svat_scalars sc = h->sc;  // tentative: committed once the block has run
    // the tentative planning below advances h->u_cur_*: put them back on every path that does not commit the block
    struct Restore {
        roger_svat* h; double prec, ta; bool commit = false;
        ~Restore() { if (!commit) { h->u_cur_prec = prec; h->u_cur_ta = ta; } }
    } restore{h, h->u_cur_prec, h->u_cur_ta};
    while ((int)plans.size() < max_steps && (int)plans.size() < (int)h->max_block) {
        roger_svat_scalars sc_try = sc;
        StepPlan p;
        SchedA a;
        { int r = schedule_first_half(h, sf, sc_try, p, a); if (r) return r; }
        uniform_first_aggregates(h, p, a.itt_day);
        schedule_second_half(h, uniform_event_flags(st, p, a), a, sc_try, p);
        const double ta_seen_a = p.u_ta_a;                                  // ta [tau] during interception .. (class a / kept)
        const double ta_final = p.class_b ? p.u_ta_b : ta_seen_a;          // ta [tau] the processes see
        if (!(ta_seen_a > st.ta_fm) || !(ta_final > st.ta_fm)) break;       // a cold step: snow may accrue, stop before it
        finish_step_scalars(h, p, sc_try);
        sc = sc_try;
        set_flux_mask(h, caller_wants_fluxes, p);
        plans.push_back(p);
        dts.push_back((int)sc_try.dt_secs);
        if (sc.time % 86400 == 0) break;  // the next day needs a new forcing block
    }
    if (plans.size() < 2) return 0;  // not worth a block (the tentative bookkeeping is undone by `restore`)
    const int n = (int)plans.size();
    if (write_fluxes) { int r = ensure_flux_planes(h); if (r) return r; }
    if (!h->d_plans) {
        CU(cudaMalloc(&h->d_plans, sizeof(StepPlan) * h->max_block));
        CU(cudaMalloc(&h->d_block_flags, sizeof(unsigned) * h->max_block));
        CU(cudaHostAlloc(&h->h_plans, sizeof(StepPlan) * h->max_block, cudaHostAllocDefault));
        CU(cudaHostAlloc(&h->h_block_flags, sizeof(unsigned) * h->max_block, cudaHostAllocDefault));
    }
    CU(cudaEventRecord(h->ev0, h->stream));
    CU(cudaMemsetAsync(h->d_block_flags, 0, sizeof(unsigned) * n, h->stream));
    const Planes pl = h->planes();
    if (queued) {
        const int threads = ROGER_BLOCK;
        const int blocks = (int)((h->C + threads - 1) / threads);
        for (int k = 0; k < n; k++) {
            const StepPlan& p = plans[k];
#define LAUNCH(LAT, FLUX) \
    k_svat_step<LAT, FLUX><<<blocks, threads, 0, h->stream>>>(pl, h->d_prec_m1, h->forc[1], h->forc[2], p, h->d_block_flags + k)
            if (st.enable_lateral_flow) { if (write_fluxes) LAUNCH(true, true); else LAUNCH(true, false); }
            else { if (write_fluxes) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
            CU(cudaGetLastError());
            launches++;
            if (!h->rate_fields.empty()) {
                const int nr = (int)h->rate_fields.size();
                k_accumulate_rates<<<(int)((h->C + 255) / 256), 256, 0, h->stream>>>(pl, h->d_rates, h->d_rate_fields, h->d_rate_fields + nr, nr, h->d_rates_prev);
                CU(cudaGetLastError());
                launches++;
                h->diag_count++;
            }
        }
        h->soil_kh_stale = !write_fluxes;
    } else {
        const int threads = ROGER_STEPS_BLOCK;
        const int blocks = (int)((h->C + threads - 1) / threads);
        memcpy(h->h_plans, plans.data(), sizeof(StepPlan) * n);
        CU(cudaMemcpyAsync(h->d_plans, h->h_plans, sizeof(StepPlan) * n, cudaMemcpyHostToDevice, h->stream));
        if (st.enable_lateral_flow) k_svat_steps<true><<<blocks, threads, 0, h->stream>>>(pl, (const StepPlan*)h->d_plans, n, h->d_block_flags);
        else k_svat_steps<false><<<blocks, threads, 0, h->stream>>>(pl, (const StepPlan*)h->d_plans, n, h->d_block_flags);
        h->soil_kh_stale = true;
        CU(cudaGetLastError());
        launches++;
    }
    unsigned all = 0;
    std::vector<uint32_t> words(n);
    { int r = fetch_flags(h, h->d_block_flags, n, words.data(), h->ev1); if (r) return r; }
    for (int k = 0; k < n; k++) all |= words[k];
    if (all & snow_bits) return fail(ROGER_EINVAL, "internal error: snow appeared inside a block scheduled as snow-free");
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_kernel_ms += ms;
    h->last_launches += launches;
    h->snow_flags = words[n - 1] & ~FLAG_INSANE;
    h->last_dts.insert(h->last_dts.end(), dts.begin(), dts.end());
    h->sc = sc;
    h->sc.sanity_ok = (all & FLAG_INSANE) ? 0 : 1;
    restore.commit = true;
    *n_done = n;
    return 0;
}

int roger_svat_step(roger_svat* h, int write_fluxes) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_step: null handle");
    h->last_kernel_ms = 0;
    h->last_launches = 0;
    h->last_dts.clear();
    return step_impl(h, write_fluxes);
}

// Up to max_steps adaptive steps, never past the next midnight (a new forcing block is due there).  Uses the day-blocked
// launch where the stretch qualifies (run_block) and single steps otherwise; the results are identical either way.
int roger_svat_run_steps(roger_svat* h, int write_fluxes, int max_steps, int* n_steps) {
    if (!h || max_steps < 1) return fail(ROGER_EINVAL, "roger_svat_run_steps: bad argument");
    h->last_kernel_ms = 0;
    h->last_launches = 0;
    h->last_dts.clear();
    int n = 0;
    do {
        int nb = 0;
        if (h->step_batching == 1 && !write_fluxes && h->rate_fields.empty() && !h->rates_need_flux) {
            int r = run_block(h, max_steps - n, &nb, false, 0, false);
            if (r) return r;
        } else if (h->step_batching == 2) {
            int r = run_block(h, max_steps - n, &nb, true, (write_fluxes || h->rates_need_flux) ? 1 : 0, write_fluxes != 0);
            if (r) return r;
        }
        if (nb == 0) {
            int r = step_impl(h, write_fluxes);
            if (r) return r;
            nb = 1;
        }
        n += nb;
    } while (h->sc.time % 86400 != 0 && n < max_steps);
    if (n_steps) *n_steps = n;
    return 0;
}

int roger_svat_run_day(roger_svat* h, int write_fluxes, int* n_steps) {
    return roger_svat_run_steps(h, write_fluxes, 144, n_steps);
}

// Diagnostics accumulators on the device: after every step the named planes are folded into accumulator planes.
int roger_svat_set_diagnostics(roger_svat* h, const int* fields, const int* modes, int n) {
    if (!h || n < 0 || n > 64 || (n > 0 && !fields)) return fail(ROGER_EINVAL, "roger_svat_set_diagnostics: bad argument");
    CU(cudaSetDevice(h->device));
    for (int k = 0; k < n; k++) {
        if (fields[k] < 0 || fields[k] >= RF_COUNT) return fail(ROGER_EINVAL, "roger_svat_set_diagnostics: bad field id");
        const int m = modes ? modes[k] : ROGER_DIAG_RATE;
        if (m < ROGER_DIAG_RATE || m > ROGER_DIAG_MINIMUM_T) return fail(ROGER_EINVAL, "roger_svat_set_diagnostics: bad mode");
    }
    cudaFree(h->d_rates); h->d_rates = nullptr;
    cudaFree(h->d_rates_prev); h->d_rates_prev = nullptr;
    cudaFree(h->d_rate_fields); h->d_rate_fields = nullptr;
    h->rate_fields.assign(fields, fields + n);
    h->rate_modes.assign(n, ROGER_DIAG_RATE);
    if (modes) h->rate_modes.assign(modes, modes + n);
    h->rates_need_flux = false;
    h->diag_count = 0;
    for (int f : h->rate_fields) h->rates_need_flux = h->rates_need_flux || kind_is_F(f) || is_soil_kh(f);  // k, h: full kernel only
    if (n == 0) return 0;
    if (h->rates_need_flux) { int r = ensure_flux_planes(h); if (r) return r; }
    const size_t bytes = sizeof(double) * (size_t)n * (size_t)h->C;
    CU(cudaMalloc(&h->d_rates, bytes));
    CU(cudaMemsetAsync(h->d_rates, 0, bytes, h->stream));
    bool two_level = false;
    for (int m : h->rate_modes) two_level = two_level || m == ROGER_DIAG_MAXIMUM_T || m == ROGER_DIAG_MINIMUM_T;
    if (two_level) {  // [taum1] of the first accumulated step: the planes as they are now
        CU(cudaMalloc(&h->d_rates_prev, bytes));
        CU(cudaMemsetAsync(h->d_rates_prev, 0, bytes, h->stream));
        for (int k = 0; k < n; k++)
            if (h->rate_modes[k] == ROGER_DIAG_MAXIMUM_T || h->rate_modes[k] == ROGER_DIAG_MINIMUM_T) {
                if (is_soil_kh(h->rate_fields[k])) { int r = refresh_soil_kh(h); if (r) return r; }
                k_gather_plane<<<h->grid256(), 256, 0, h->stream>>>(h->planes(), h->rate_fields[k], h->d_rates_prev + (size_t)k * h->C, nullptr);
            }
        CU(cudaGetLastError());
    }
    CU(cudaMalloc(&h->d_rate_fields, sizeof(int) * 2 * n));
    CU(cudaMemcpyAsync(h->d_rate_fields, h->rate_fields.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_rate_fields + n, h->rate_modes.data(), sizeof(int) * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_svat_set_rate_fields(roger_svat* h, const int* fields, int n) { return roger_svat_set_diagnostics(h, fields, nullptr, n); }

// the accumulated planes (n_fields x n_cells doubles, in the order given to roger_svat_set_diagnostics); AVERAGE planes are
// divided by the number of accumulated steps (average.py:94-97); reset != 0 zeroes the accumulators afterwards (start of the
// next output interval)
int roger_svat_download_diagnostics(roger_svat* h, double* host, int reset) {
    if (!h || !host) return fail(ROGER_EINVAL, "roger_svat_download_diagnostics: bad argument");
    if (h->rate_fields.empty()) return fail(ROGER_EINVAL, "roger_svat_download_diagnostics: no diagnostics set");
    CU(cudaSetDevice(h->device));
    const size_t plane = sizeof(double) * (size_t)h->C;
    const size_t bytes = plane * h->rate_fields.size();
    if (!h->d_order) {
        CU(cudaMemcpyAsync(host, h->d_rates, bytes, cudaMemcpyDeviceToHost, h->stream));
    } else {
        for (size_t k = 0; k < h->rate_fields.size(); k++) {  // back to the caller's cell numbering, one plane at a time
            k_gather_linear<<<h->grid256(), 256, 0, h->stream>>>(h->d_rates + k * (size_t)h->C, h->d_stage, h->d_order, h->C);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(host + k * (size_t)h->C, h->d_stage, plane, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    if (reset) CU(cudaMemsetAsync(h->d_rates, 0, bytes, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (h->diag_count > 0)
        for (size_t k = 0; k < h->rate_fields.size(); k++)
            if (h->rate_modes[k] == ROGER_DIAG_AVERAGE) {
                double* p = host + k * (size_t)h->C;
                const double cnt = (double)h->diag_count;
                for (int64_t i = 0; i < h->C; i++) p[i] = p[i] / cnt;
            }
    if (reset) h->diag_count = 0;
    return 0;
}

int roger_svat_download_rates(roger_svat* h, double* host, int reset) { return roger_svat_download_diagnostics(h, host, reset); }

uint64_t roger_svat_rates_ptr(roger_svat* h) { return h ? (uint64_t)(uintptr_t)h->d_rates : 0; }

int roger_svat_last_step_dts(roger_svat* h, int* dt_secs, int capacity) {
    if (!h || (!dt_secs && capacity > 0)) return fail(ROGER_EINVAL, "roger_svat_last_step_dts: bad argument");
    const int n = (int)h->last_dts.size();
    for (int k = 0; k < n && k < capacity; k++) dt_secs[k] = h->last_dts[k];
    return n;
}

int roger_svat_set_block_stepping(roger_svat* h, int enabled, int max_block) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_set_block_stepping: null handle");
    if (enabled < 0 || enabled > 2) return fail(ROGER_EINVAL, "roger_svat_set_block_stepping: mode 0 (off), 1 (one launch per batch) or 2 (queued launches)");
    h->step_batching = enabled;
    if (max_block >= 2 && max_block <= 144 && !h->d_plans) h->max_block = max_block;
    return 0;
}

int roger_svat_set_flag_reducer(roger_svat* h, roger_flag_reduce_fn fn, void* user) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_set_flag_reducer: null handle");
    h->reduce_fn = fn;
    h->reduce_user = user;
    return 0;
}

// window sums of the hourly forcing aggregates: the six-load form against the 144-slot form, on a caller-supplied row
// (host code of the same templates the kernels instantiate; tests/test_abi.py feeds rows with NaN, -0.0 and every itt_day)
int roger_selftest_window_sums(const double* row144, int itt_day, double* out4) {
    if (!row144 || !out4 || itt_day < 0 || itt_day > 143) return fail(ROGER_EINVAL, "roger_selftest_window_sums: bad argument");
    out4[0] = RowAgg::win_sum(row144, itt_day);
    out4[1] = RowAgg::win_sum_full(row144, itt_day);
    out4[2] = RowAgg::win_nanmean(row144, itt_day);
    out4[3] = RowAgg::win_nanmean_full(row144, itt_day);
    return 0;
}

int roger_nccl_unique_id(uint8_t* id128) {
    if (!id128) return fail(ROGER_EINVAL, "roger_nccl_unique_id: null buffer");
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(ROGER_ENODEV, "NCCL unavailable: " + api.why);
    NcclApi::UniqueId id;
    NCCL(api.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}

int roger_svat_enable_nccl_flags(roger_svat* h, const uint8_t* id128, int rank, int world) {
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ROGER_EINVAL, "roger_svat_enable_nccl_flags: bad argument");
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(ROGER_ENODEV, "NCCL unavailable: " + api.why);
    CU(cudaSetDevice(h->device));
    NcclApi::UniqueId id;
    memcpy(id.internal, id128, 128);
    void* comm = nullptr;
    NCCL(api.CommInitRank(&comm, world, id, rank));
    if (h->nccl_comm) api.CommDestroy(h->nccl_comm);
    h->nccl_comm = comm;
    h->nccl_world = world;
    cudaFree(h->d_gather);
    cudaFreeHost(h->h_gather);
    h->d_gather = nullptr; h->h_gather = nullptr;
    CU(cudaMalloc(&h->d_gather, sizeof(unsigned) * roger_svat::kMaxFlagWords * world));
    CU(cudaHostAlloc(&h->h_gather, sizeof(unsigned) * roger_svat::kMaxFlagWords * world, cudaHostAllocDefault));
    h->snow_flags_valid = false;  // the flags known so far were this rank's own
    return 0;
}

// Re-orders the cells ON THE DEVICE so that neighbouring threads hold cells of similar infiltration capacity
// (sort key ks * (1 - sealing), masked cells last).  Columns are independent, so any order computes the same numbers; what
// changes is which cells share a warp: the macropore / shrinkage-crack sub-step loops run only for ponded cells, and with
// spatially uncorrelated parameters nearly every warp holds a few of them and pays for all (measured on a rain step at
// 4096^2 heterogeneous cells: 15 of 32 lanes active in the macropore loop).  Every entry point keeps speaking the caller's
// cell numbering: uploads, downloads, forcing blocks and diagnostics are permuted on the way.  Call it after the
// parameters are uploaded and before stepping; calling it again re-sorts from the current planes.
int roger_svat_optimize_cell_order(roger_svat* h) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_optimize_cell_order: null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->copy_stream));  // no read-back in flight while the planes move
    const int64_t C = h->C;
    const Planes pl = h->planes();
    double *keys = nullptr, *keys_out = nullptr;
    int *ids = nullptr, *pos = nullptr, *new_order = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    CU(cudaMalloc(&keys, sizeof(double) * C)); CU(cudaMalloc(&keys_out, sizeof(double) * C));
    CU(cudaMalloc(&ids, sizeof(int) * C)); CU(cudaMalloc(&pos, sizeof(int) * C)); CU(cudaMalloc(&new_order, sizeof(int) * C));
    k_order_keys<<<h->grid256(), 256, 0, h->stream>>>(pl, keys, ids);
    CU(cudaGetLastError());
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, ids, pos, (int)C, 0, 64, h->stream));
    CU(cudaMalloc(&tmp, tmp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, ids, pos, (int)C, 0, 64, h->stream));  // stable
    // pos[d] = current device position that moves to position d; in the caller's numbering: old_order[pos[d]]
    k_compose_order<<<h->grid256(), 256, 0, h->stream>>>(h->d_order, pos, new_order, C);
    CU(cudaGetLastError());
    // move every plane: gather with the old order into the host-order staging plane, scatter with the new one
    const int n_fields = h->d_fl ? RF_COUNT : kNumPS;
    for (int f = 0; f < n_fields; f++) {
        k_gather_plane<<<h->grid256(), 256, 0, h->stream>>>(pl, f, h->d_stage, h->d_order);
        k_scatter_plane<<<h->grid256(), 256, 0, h->stream>>>(pl, f, h->d_stage, new_order);
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));  // (d_prec_m1 is scratch of a single step: nothing to move)
    if (!h->rate_fields.empty()) {  // accumulators are positional: restart them
        CU(cudaMemsetAsync(h->d_rates, 0, sizeof(double) * h->rate_fields.size() * (size_t)C, h->stream));
        h->diag_count = 0;
    }
    cudaFree(h->d_order);
    h->d_order = new_order;
    if (!h->d_inv) CU(cudaMalloc(&h->d_inv, sizeof(int) * C));
    k_invert_order<<<h->grid256(), 256, 0, h->stream>>>(h->d_order, h->d_inv, C);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    cudaFree(keys); cudaFree(keys_out); cudaFree(ids); cudaFree(pos); cudaFree(tmp);
    h->have_forcing = h->st.forcing_uniform ? h->have_forcing : false;  // a per-cell day block is positional: hand it over again
    return 0;
}

// In-memory checkpoint of the run: every plane, the scheduler's scalars and flags, the uniform forcing day.  The reference's
// counterpart is its restart file (roger/restart.py:32-174: all variables with `write_to_restart` + scalars); here the copy
// stays in HBM (one device-to-device copy of the plane allocation) so that a stretch of the run can be repeated from exactly
// the same state -- bench.py times `value` and `e2e` over the same simulated window this way.  Per-cell forcing day blocks
// and the diagnostics accumulators' contents are not part of it: hand the day block over again after a rollback.
int roger_svat_checkpoint(roger_svat* h) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_checkpoint: null handle");
    CU(cudaSetDevice(h->device));
    const size_t ps_bytes = kTileBytesPS * (size_t)h->n_tiles, fl_bytes = kTileBytesF * (size_t)h->n_tiles;
    if (!h->d_ckpt_ps) CU(cudaMalloc(&h->d_ckpt_ps, ps_bytes));
    CU(cudaMemcpyAsync(h->d_ckpt_ps, h->d_ps, ps_bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (h->d_fl) {
        if (!h->d_ckpt_fl) CU(cudaMalloc(&h->d_ckpt_fl, fl_bytes));
        CU(cudaMemcpyAsync(h->d_ckpt_fl, h->d_fl, fl_bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    roger_svat::Saved& v = h->saved;
    v.sc = h->sc; v.snow_flags = h->snow_flags; v.day_flags = h->day_flags; v.snow_flags_valid = h->snow_flags_valid;
    v.ks_ss_synced = h->ks_ss_synced; v.soil_kh_stale = h->soil_kh_stale; v.have_forcing = h->have_forcing;
    v.u_cur_valid = h->u_cur_valid; v.u_cur_prec = h->u_cur_prec; v.u_cur_ta = h->u_cur_ta; v.diag_count = h->diag_count;
    memcpy(v.h_forc, h->h_forc, sizeof(v.h_forc));
    h->have_ckpt = true;
    return 0;
}

int roger_svat_rollback(roger_svat* h) {
    if (!h || !h->have_ckpt) return fail(ROGER_EINVAL, "roger_svat_rollback: no checkpoint");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->copy_stream));
    CU(cudaMemcpyAsync(h->d_ps, h->d_ckpt_ps, kTileBytesPS * (size_t)h->n_tiles, cudaMemcpyDeviceToDevice, h->stream));
    if (h->d_fl && h->d_ckpt_fl) CU(cudaMemcpyAsync(h->d_fl, h->d_ckpt_fl, kTileBytesF * (size_t)h->n_tiles, cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    const roger_svat::Saved& v = h->saved;
    h->sc = v.sc; h->snow_flags = v.snow_flags; h->day_flags = v.day_flags; h->snow_flags_valid = v.snow_flags_valid;
    h->ks_ss_synced = v.ks_ss_synced; h->soil_kh_stale = v.soil_kh_stale;
    h->have_forcing = h->st.forcing_uniform ? v.have_forcing : false;
    h->u_cur_valid = v.u_cur_valid; h->u_cur_prec = v.u_cur_prec; h->u_cur_ta = v.u_cur_ta; h->diag_count = v.diag_count;
    memcpy(h->h_forc, v.h_forc, sizeof(v.h_forc));
    return 0;
}

int roger_svat_synchronize(roger_svat* h) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_synchronize: null handle");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int roger_svat_last_timing(roger_svat* h, double* kernel_ms, int64_t* launches) {
    if (!h) return fail(ROGER_EINVAL, "roger_svat_last_timing: null handle");
    if (kernel_ms) *kernel_ms = h->last_kernel_ms;
    if (launches) *launches = h->last_launches;
    return 0;
}

}  // extern "C"
