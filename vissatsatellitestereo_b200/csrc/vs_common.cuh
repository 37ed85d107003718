// Shared declarations for libvissat_b200 (sm_100a).  Internal header; the public ABI is include/vissat_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/vissat_b200.h"

#define VS_MAX_DEGREE 5
#define VS_MAX_TERMS 56  // C(5+3,3)
#define VS_MAX_STREAMS 4

// ENU -> (fractional col, fractional row, altitude) polynomial in box-normalised coordinates
// p = ((x,y,z) - center) / half, |p_i| <= 1 inside the validated box.  Terms are ordered by
// (i, j, k) lexicographic with i + j + k <= degree (see poly_term_order()).
struct VsPoly {
    int degree;  // 0 = exact chain only
    int n_terms;
    double center[3];
    double inv_half[3];
    double coef[3][VS_MAX_TERMS];  // [0] = col_f, [1] = row_f, [2] = alt
    // mixed-precision split of the same polynomial (used when d64 > 0): terms of total degree <= d64 (1 or 2)
    // in float64 (degree-d64 index order), terms of degree d64+1..D in float32 (degree-D index order, low
    // terms zero).  d64 == 0: every term in float64.
    int d64;
    double coefLo[3][10];
    float coefR[3][VS_MAX_TERMS];
};

struct VsGeoParams {
    double lat0, lon0, alt0;
    // precomputed for the exact chain (pymap3d order of operations)
    double sin_lat0, cos_lat0, sin_lon0, cos_lon0;
    double x0, y0, z0;  // ECEF of the ENU origin
    double lam0;        // UTM central meridian (rad)
    double north_off;   // 0 or 1e7
    double ul_e, ul_n, row_res, col_res;
    int xsize, ysize;
};

// Kernel-side description of the peer stores of stage B (vs_set_exchange): passed by value to k_grid_finalize.
// plane[j] = rank j's (rows_j + halo rows) x W plane of the view being finalised, as mapped into this process.
struct VsPeerPlan {
    int n;                           // ranks; 0 = no peer stores
    int halo;
    unsigned inv;                    // floor(2^32 * n / H): y * inv >> 32 ~ band of row y
    int row0[VS_MAX_RANKS + 1];      // band boundaries; row0[n] = H
    float* plane[VS_MAX_RANKS];
    // sparse exchange: occupancy bitmaps of the ranks (nullptr = dense exchange: empty tiles are stored as NaN)
    uint32_t* occ[VS_MAX_RANKS];
    int occ_words, tiles_x, view;    // words per tile, tile columns, global view index of this plane
};

// Local occupancy marking of stage B (vs_set_occupancy), passed by value to k_grid_finalize.
struct VsOccPlan {
    uint32_t* occ;
    int occ_words, tiles_x, view;
};

struct vs_ctx {
    int device;
    bool aoi_set;
    vs_aoi aoi;
    VsGeoParams geo;
    VsPoly poly;
    vs_fit_info fit;
    double ambiguity_eps;
    uint64_t launches;
    int sm_count;
    // small device scratch for setup-time evaluations
    double* d_scratch;
    size_t scratch_doubles;
    void* d_exact;  // VsExactParams on the device (geo_chain.cuh)
    bool no_tma;    // true unless VISSAT_TMA=1: stage B uses the plain-load kernels (see api.cu)
    bool k1_warp_agg; // VISSAT_K1_WARPAGG=1 (read at context creation): warp-aggregated scatter in K1 (A/B variant)
    int k2_mode;    // stage-B kernel choice, read at context creation: 0 = auto (float-space kernel for dense calls, key-space
                    // kernel for the sparse mode), 1 = VISSAT_K2_LEGACY=1 (float-space everywhere), 2 = VISSAT_K2_KEYS=1
    // vs_views_to_dsm runs odd and even views on two internal streams (stage A of one view overlaps stage B of the
    // previous one: they are bound by different pipes); the odd views scatter into a second, library-owned key grid
    cudaStream_t side_stream[VS_MAX_STREAMS];
    cudaEvent_t fork_event, join_event[VS_MAX_STREAMS];
    uint32_t* d_keygrid_extra[VS_MAX_STREAMS];   // [0] unused: stream 0 scatters into the caller's key grid
    size_t keygrid_extra_cells;
    uint32_t* d_keygrid_alt[VS_MAX_STREAMS];     // second key grid of every internal stream (stage B zeroes the other one)
    size_t keygrid_alt_cells;
    // VISSAT_PRIO=1 / 2 (A/B): stage B (1) or stage A (2) of every internal stream runs on a high-priority twin stream
    int prio_mode;
    cudaStream_t prio_stream[VS_MAX_STREAMS];
    cudaEvent_t prio_ev_a[VS_MAX_STREAMS], prio_ev_b[VS_MAX_STREAMS], prio_join[VS_MAX_STREAMS];
    bool fold_clear;    // VISSAT_FOLD_CLEAR=0: keep the whole-grid memset per view (A/B)
    int n_streams;      // VISSAT_STREAMS=1..4 (default 4: measured 4.66 / 3.78 / 3.49 / 3.45 ms per C2 step for 1..4)
    // optional per-view kernel timing of vs_views_to_dsm
    // peer-store exchange of stage B (exchange.cu)
    bool xch_on;
    vs_exchange xch;
    // occupancy marking without an exchange (vs_set_occupancy)
    uint32_t* occ;
    int occ_words;
    const float* occ_stack_base;
    int64_t occ_view0;
    // sparse mode of vs_views_to_dsm (occupancy bitmap in use): per internal stream a "touched" byte per tile, written
    // by stage A; stage B and the key-grid clear visit only touched tiles (pipeline.cu)
    unsigned char* d_touched[VS_MAX_STREAMS];
    size_t touched_tiles;
    unsigned char* cur_touched;   // the map of the view being rasterised (nullptr: dense mode)
    // scratch of vs_fuse_views_sparse: per-bin tile lists (fuse.cu)
    int* d_fuse_plan;
    size_t fuse_plan_ints;
    // co-scheduled stage A+B path (stage_ab.cu): three rotating key grids per internal stream, two queue counters per launch
    uint32_t* d_keygrid_ab[VS_MAX_STREAMS][3];
    size_t keygrid_ab_cells;
    int* d_ab_counters;
    size_t ab_counters_ints;
    bool ab_on;         // vs_set_coschedule / VISSAT_AB (default on)
    int ab_streams;     // VISSAT_AB_STREAMS (default 2): internal streams of the co-scheduled path
    bool timing;
    std::vector<cudaEvent_t> ev_pool;   // 3 events per logged view: before A, between A and B, after B
    size_t ev_used;
};

void vs_set_error(const std::string& msg);
int vs_cuda_fail(cudaError_t e, const char* what);

#define VS_CUDA(call)                                          \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return vs_cuda_fail(e__, #call); \
    } while (0)

#define VS_CHECK_LAUNCH(ctx, name)                                 \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return vs_cuda_fail(e__, name);    \
        (ctx)->launches++;                                         \
    } while (0)

#define VS_REQUIRE(cond, msg)           \
    do {                                \
        if (!(cond)) {                  \
            vs_set_error(msg);          \
            return VS_ERR_INVALID;      \
        }                               \
    } while (0)

struct VsDeviceGuard {
    int prev;
    bool ok;
    explicit VsDeviceGuard(int dev) : prev(-1), ok(false) {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || (cudaSetDevice(dev) == cudaSuccess);
        if (prev == dev) prev = -1;
    }
    ~VsDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- order-preserving keys ------------------------------------------------------------------------------
// key(a) < key(b)  <=>  a < b for non-NaN floats; 0 is reserved for "empty" (it is the image of a negative
// NaN, which never occurs: altitudes entering the scatter are finite).
__host__ __device__ __forceinline__ uint32_t vs_key32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    // negative: invert all bits; non-negative: set the top bit.  Branch-free: b ^ (sign_mask | 0x80000000)
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float vs_unkey32(uint32_t k) {
    // top bit set: clear it; top bit clear: invert everything.  Branch-free: k ^ (0x80000000 | ~sign_mask).
    // k == 0 -> 0xffffffff = NaN (empty cell)
    const uint32_t m = (uint32_t)((int32_t)k >> 31);          // 0xffffffff if the top bit is set
    return __uint_as_float(k ^ (0x80000000u | ~m));
}
__device__ __forceinline__ unsigned long long vs_key64(double d) {
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double vs_unkey64(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);  // k==0 -> all ones = NaN
}

// host-side helpers implemented in api.cu
int vs_ensure_side_streams(vs_ctx* ctx, int n);   // side_stream[0..n), join_event[0..n), fork_event
int vs_ensure_scratch(vs_ctx* ctx, size_t doubles);
int vs_upload_exact_params(vs_ctx* ctx);  // aoi_fit.cu
