// Stage A kernels: depth map (or explicit points) -> per-cell max-height key grid.
//
// K1  k_unproject_scatter   aggregate_2p5d_util.py:75-98 + lib/proj_to_grid.py:42-61, fused.
//     Streams the depth map with 16-byte loads (4 pixels per thread per step, grid-stride over a persistent
//     grid sized to the SM count), unprojects in float64 (the matrix is pre-composed with the box
//     normalisation so the perspective divide yields the polynomial's arguments directly), evaluates the
//     per-AOI ENU->(col,row,alt) polynomial (aoi_fit.cu), and scatters with a no-return atomicMax (RED.MAX)
//     on order-preserving float32 keys.  Consecutive pixels of one thread that fall in the same cell are
//     merged in registers before the atomic.  Bound by the FP64 pipe and the L2 atomic units, not by HBM
//     (4 bytes read per pixel); see DESIGN.md.
//     Points outside the fitted altitude range take the exact chain through an out-of-line call (rare).
// K1x k_unproject_scatter_exact  exact chain for every pixel, used only when the polynomial is disabled.
// K1b k_points_scatter      lib/proj_to_grid.py:42-61 for explicit float64 (E,N,alt) rows, 64-bit keys.
#include "rasterize_common.cuh"

using namespace vsras;

namespace {

// K1.  Requires n_pix < 2^32 and W < 2^16 for the multiply-shift row index (checked on the host).
// LEAN: the throughput variant -- no counters, no height map, no tile marking, no warp aggregation; those options cost
// registers in the streaming loop even when they are switched off at run time (80-register cap, 3 CTAs per SM).
template <int D, int D64, bool LEAN>
__global__ void __launch_bounds__(kThreads, D64 == 1 ? 3 : 2)
k_unproject_scatter(RasterParams p_in, PolyCoefs pc, const VsExactParams* __restrict__ ex, const float* __restrict__ depth,
                    uint32_t* __restrict__ keygrid, float* __restrict__ height_map_in,
                    unsigned long long* __restrict__ stats_in) {
    RasterParams p = p_in;
    if (LEAN) {
        p.touched = nullptr;
        p.warp_agg = 0;
    }
    float* __restrict__ height_map = LEAN ? nullptr : height_map_in;
    unsigned long long* __restrict__ stats = LEAN ? nullptr : stats_in;
    const int64_t n_pix = (int64_t)p.H * p.W;
    const unsigned n_chunks = (unsigned)((n_pix + PX - 1) / PX);
    const bool audit = !LEAN && stats != nullptr;
    const bool want_hm = !LEAN && height_map != nullptr;
    unsigned cnt[4] = {0, 0, 0, 0};

    // Row pitch a multiple of PX (the usual case): no chunk straddles a row and chunk c is the aligned float4 c, so the
    // depth of the NEXT chunk of this thread is requested (cp.async into a per-thread shared-memory slot, no registers
    // held) before the current one is processed; the global-load latency is otherwise exposed once per iteration and
    // the kernel has only 24 warps per SM to hide it with.  Each thread reads only its own slot: no block barrier.
    __shared__ __align__(16) float4 s_pre[2][kThreads];
    const bool simple = (p.W % PX) == 0;
    const unsigned n_full = (unsigned)(n_pix / PX);
    const unsigned stride = gridDim.x * blockDim.x;
    const float4* __restrict__ depth4 = reinterpret_cast<const float4*>(depth);
    // shared-window addresses of this thread's two slots, computed once (32-bit; a slot is 16 bytes)
    const unsigned slot0 = (unsigned)__cvta_generic_to_shared(&s_pre[0][threadIdx.x]);
    const unsigned slot_sum = 2u * slot0 + (unsigned)(sizeof(float4) * kThreads);   // slot0 + slot1
    unsigned cur_slot = slot0;
    {
        const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
        if (simple && first < n_full) cp_async16_s(slot0, depth4 + first);
        cp_async_commit();
    }
    for (unsigned chunk = blockIdx.x * blockDim.x + threadIdx.x; chunk < n_chunks; chunk += stride) {
        const unsigned base = chunk * PX;
        const unsigned next_slot = slot_sum - cur_slot;   // the other slot
        if (simple && chunk + stride < n_full) cp_async16_s(next_slot, depth4 + (chunk + stride));
        cp_async_commit();
        cp_async_wait_but_one();
        const float4 cur = ld_shared_f4(cur_slot);
        cur_slot = next_slot;
        // row = base / W by multiply-shift (magic = ceil(2^48 / W), exact for base < 2^32, W < 2^16)
        const unsigned row0 = (unsigned)(((unsigned long long)base * p.w_magic) >> 48);
        const unsigned col0 = base - row0 * (unsigned)p.W;
        // straddles rows / ragged tail (rare).  With a pitch that is a multiple of PX only the last, partial chunk can.
        if (simple ? (chunk >= n_full) : (col0 + PX > (unsigned)p.W || (int64_t)base + PX > n_pix)) {
            scatter_chunk_generic<D, D64>(p, pc, ex, depth, base, n_pix, keygrid, height_map, audit, cnt);
            continue;
        }
        const float4 t4 = simple ? cur : ld_stream_f4(reinterpret_cast<const float4*>(depth + base));
        const float d4[PX] = {t4.x, t4.y, t4.z, t4.w};
        // aggregate_2p5d_util.py:76: depth <= 0 (and NaN) is invalid
        if (!(d4[0] > 0.0f || d4[1] > 0.0f || d4[2] > 0.0f || d4[3] > 0.0f)) {
            if (want_hm)
                *reinterpret_cast<float4*>(height_map + base) =
                    make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
            continue;
        }

        // ---- aggregate_2p5d_util.py:86-90: X = M . [col, row, 1, depth]; (u, v, w) = box-normalised X / X_3
        double u[PX], v[PX], w[PX];
        {
            const double fr = (double)row0;
            const double fc0 = (double)col0;
            const double r0 = fma(p.Mn[0][1], fr, p.Mn[0][2]), r1 = fma(p.Mn[1][1], fr, p.Mn[1][2]);
            const double r2 = fma(p.Mn[2][1], fr, p.Mn[2][2]), r3 = fma(p.M3[1], fr, p.M3[2]);
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                const double fc = fc0 + (double)i;   // exact
                const double d = (double)d4[i];
                const double rw = fast_rcp(fma(p.M3[0], fc, fma(p.M3[3], d, r3)));
                u[i] = fma(p.Mn[0][0], fc, fma(p.Mn[0][3], d, r0)) * rw;
                v[i] = fma(p.Mn[1][0], fc, fma(p.Mn[1][3], d, r1)) * rw;
                w[i] = fma(p.Mn[2][0], fc, fma(p.Mn[2][3], d, r2)) * rw;
            }
        }
        // ---- classification on float32 copies.  |x| <= 1 is false for NaN/inf, so a non-finite position is
        // never "in the box" (aggregate_2p5d_util.py:92 / lib/proj_to_grid.py:48); the box has a margin around
        // the grid, so float32 rounding at its faces is harmless.
        float uf[PX], vf[PX], wf[PX];
        bool fast[PX];
        bool any_slow = false;
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            uf[i] = (float)u[i];
            vf[i] = (float)v[i];
            wf[i] = (float)w[i];
            const bool in_box = d4[i] > 0.0f && fabsf(uf[i]) <= 1.0f && fabsf(vf[i]) <= 1.0f;
            const bool in_alt = fabsf(wf[i]) <= 1.0f;
            fast[i] = in_box && in_alt;
            any_slow |= in_box && !in_alt;
        }
        if (audit || want_hm) {   // block-uniform
            float hm[PX];
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                const bool finite = d4[i] > 0.0f && fabsf(uf[i]) < CUDART_INF_F && fabsf(vf[i]) < CUDART_INF_F &&
                                    fabsf(wf[i]) < CUDART_INF_F;
                cnt[VS_STAT_VALID] += finite;
                hm[i] = finite ? (float)fma(w[i], p.half_z, p.center_z) : CUDART_NAN_F;
            }
            if (want_hm) *reinterpret_cast<float4*>(height_map + base) = make_float4(hm[0], hm[1], hm[2], hm[3]);
        }

        // ---- ENU -> (fractional col, fractional row, altitude): the per-AOI polynomial, 4 pixels in lockstep
        int ci[PX], ri[PX];
        bool amb[PX];
        uint32_t key[PX];
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            double val[PX];
            if (D64 > 0) {
                vs_poly_eval_n<(D64 > 0 ? D64 : 1), PX, double>(pc.c64[o], u, v, w, val);
                float hi[PX];
                vs_poly_eval_n_f32x2<D, PX>(pc.c32[o], uf, vf, wf, hi);
#pragma unroll
                for (int i = 0; i < PX; ++i) val[i] += (double)hi[i];
            } else {
                vs_poly_eval_n<D, PX, double>(pc.c64[o], u, v, w, val);
            }
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                if (o == 2) {
                    key[i] = vs_key32((float)val[i]);
                } else {
                    const int q = LEAN ? floor_to_int_fast(val[i]) : __double2int_rd(val[i]);  // floor, lib/proj_to_grid.py:42-43
                    if (o == 0) ci[i] = q; else ri[i] = q;
                    if (audit) {
                        const bool a = fabs(val[i] - rint(val[i])) < p.eps;
                        amb[i] = (o == 0) ? a : (amb[i] || a);
                    }
                }
            }
        }

        // ---- lib/proj_to_grid.py:44-61: bounds mask + per-cell max.  Same-cell neighbours merge in registers
        // (forward chain), everything predicated.
        int cell[PX];
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            if (audit && fast[i] && amb[i]) {
                // Audited mode (stats requested: the drop-in API path): a point whose polynomial row or column is within
                // eps of a cell edge -- where the polynomial's ~1e-8 cell error could change the floor -- is re-evaluated
                // with the exact chain and scattered from there (about 4e-7 of the points; the call is out of line).
                ++cnt[VS_STAT_AMBIGUOUS];
                const int r = scatter_exact_point(ex, u[i], v[i], w[i], keygrid, p.touched, p.tiles_x);
                cnt[VS_STAT_INGRID] += (r != 0);
                fast[i] = false;
            }
            fast[i] = fast[i] && (unsigned)ci[i] < (unsigned)p.xsize && (unsigned)ri[i] < (unsigned)p.ysize;
            cell[i] = ri[i] * p.xsize + ci[i];
            if (audit) cnt[VS_STAT_INGRID] += fast[i];
        }
        if (p.touched != nullptr) {   // block-uniform: sparse mode
            int prev = -1;
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                const int t = (ri[i] >> 5) * p.tiles_x + (ci[i] >> 6);
                if (fast[i] && t != prev) {
                    p.touched[t] = 1;
                    prev = t;
                }
            }
        }
#pragma unroll
        for (int i = 0; i + 1 < PX; ++i) {
            const bool same = fast[i] && fast[i + 1] && cell[i] == cell[i + 1];
            key[i + 1] = same ? max(key[i + 1], key[i]) : key[i + 1];
            fast[i] = fast[i] && !same;
        }
        if (p.warp_agg) {
            // A/B variant (north_star item 3, "warp-aggregated atomics"): lanes of the warp that hit the same cell are
            // found with match.any, their keys reduced with redux.max, and one lane issues the atomic.  Measured on
            // C4 (4 pixels per cell) and C2: see profiles/README.md -- a warp covers 128 consecutive pixels of ONE
            // image row, so only neighbouring lanes ever share a cell; the in-thread merge above already removes most
            // duplicates and the match costs more than the atomics it saves.  Kept switchable, off by default.
            const unsigned lane = threadIdx.x & 31;
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                const unsigned act = __activemask();
                const int tag = fast[i] ? cell[i] : -1 - (int)lane;
                const unsigned grp = __match_any_sync(act, tag);
                const uint32_t best = __reduce_max_sync(grp, key[i]);
                if (fast[i] && lane == (unsigned)(__ffs(grp) - 1)) atomicMax(keygrid + cell[i], best);
            }
        } else {
#pragma unroll
            for (int i = 0; i < PX; ++i)
                if (fast[i]) atomicMax(keygrid + cell[i], key[i]);
        }

        if (any_slow) {  // points outside the fitted altitude range: exact chain, out of line (rare)
#pragma unroll
            for (int i = 0; i < PX; ++i) {
                if (d4[i] > 0.0f && fabsf(uf[i]) <= 1.0f && fabsf(vf[i]) <= 1.0f && !(fabsf(wf[i]) <= 1.0f) &&
                    fabsf(wf[i]) < CUDART_INF_F) {
                    ++cnt[VS_STAT_EXACT];
                    const int r = scatter_exact_point(ex, u[i], v[i], w[i], keygrid, p.touched, p.tiles_x);
                    cnt[VS_STAT_INGRID] += (r != 0);
                    cnt[VS_STAT_AMBIGUOUS] += (r == 2);
                }
            }
        }
    }
    flush_stats(stats, cnt[0], cnt[1], cnt[2], cnt[3]);
}

// Exact chain for every pixel (polynomial disabled: vs_set_aoi(max_degree = 0) or a fit that failed validation).
__global__ void __launch_bounds__(kThreads)
k_unproject_scatter_exact(RasterParams p, VsEllipsoidConsts c, VsGeoParams g, double center_x, double half_x,
                          double center_y, double half_y, const float* __restrict__ depth,
                          uint32_t* __restrict__ keygrid, float* __restrict__ height_map,
                          unsigned long long* __restrict__ stats) {
    const bool all_pixels = true;
    const int64_t n_pix = (int64_t)p.H * p.W;
    unsigned n_valid = 0, n_ingrid = 0, n_amb = 0, n_exact = 0;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n_pix;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const float df = depth[idx];
        if (all_pixels && height_map != nullptr) height_map[idx] = CUDART_NAN_F;
        if (!(df > 0.0f)) continue;
        const int row = (int)(idx / p.W);
        const int col = (int)(idx - (int64_t)row * p.W);
        const double d = (double)df, fc = (double)col, fr = (double)row;
        const double hw = fma(p.M3[0], fc, fma(p.M3[1], fr, fma(p.M3[3], d, p.M3[2])));
        const double rw = 1.0 / hw;
        const double u = fma(p.Mn[0][0], fc, fma(p.Mn[0][1], fr, fma(p.Mn[0][3], d, p.Mn[0][2]))) * rw;
        const double v = fma(p.Mn[1][0], fc, fma(p.Mn[1][1], fr, fma(p.Mn[1][3], d, p.Mn[1][2]))) * rw;
        const double w = fma(p.Mn[2][0], fc, fma(p.Mn[2][1], fr, fma(p.Mn[2][3], d, p.Mn[2][2]))) * rw;
        if (!(fabs(u) < CUDART_INF && fabs(v) < CUDART_INF && fabs(w) < CUDART_INF)) continue;
        if (all_pixels) {
            ++n_valid;
            if (height_map != nullptr) height_map[idx] = (float)fma(w, p.half_z, p.center_z);
        }
        if (!(fabs(u) <= 1.0 && fabs(v) <= 1.0)) continue;
        ++n_exact;
        double E, N, A;
        vs_enu_to_utm_exact(c, g, fma(u, half_x, center_x), fma(v, half_y, center_y), fma(w, p.half_z, p.center_z), E, N,
                            A);
        const double colf = (E - g.ul_e) / g.col_res, rowf = (g.ul_n - N) / g.row_res;
        const double cfl = floor(colf), rfl = floor(rowf);
        if (cfl >= 0.0 && rfl >= 0.0 && cfl < (double)p.xsize && rfl < (double)p.ysize && A == A) {
            ++n_ingrid;
            const double fcx = colf - cfl, frx = rowf - rfl;
            if (fcx < p.eps || fcx > 1.0 - p.eps || frx < p.eps || frx > 1.0 - p.eps) ++n_amb;
            atomicMax(keygrid + ((int)rfl * p.xsize + (int)cfl), vs_key32((float)A));
        }
    }
    flush_stats(stats, n_valid, n_ingrid, n_amb, n_exact);
}

__global__ void __launch_bounds__(kThreads)
k_points_scatter(const double* __restrict__ pts, int64_t n, double xoff, double yoff, double xres, double yres,
                 int xsize, int ysize, unsigned long long* __restrict__ keygrid, unsigned long long* __restrict__ stats) {
    unsigned n_valid = 0, n_ingrid = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double E = pts[3 * i], N = pts[3 * i + 1], A = pts[3 * i + 2];
        // lib/proj_to_grid.py:42-43 (row uses xresolution, col uses yresolution -- as in the reference)
        const double rfl = floor(__ddiv_rn(__dsub_rn(yoff, N), xres));
        const double cfl = floor(__ddiv_rn(__dsub_rn(E, xoff), yres));
        if (A == A) ++n_valid;
        if (rfl >= 0.0 && cfl >= 0.0 && rfl < (double)ysize && cfl < (double)xsize && A == A) {  // :48, nanmax drops NaN
            ++n_ingrid;
            atomicMax(keygrid + ((int64_t)rfl * xsize + (int64_t)cfl), vs_key64(A));
        }
    }
    flush_stats(stats, n_valid, n_ingrid, 0, 0);
}

int persistent_grid(const vs_ctx* ctx, int64_t work_items, int per_sm) {
    int64_t blocks = (work_items + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

void vs_make_raster_params(const vs_ctx* ctx, int H, int W, const double* M, RasterParams* p_out, PolyCoefs* pc_out) {
    const VsPoly& P = ctx->poly;
    RasterParams& p = *p_out;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) p.Mn[i][j] = (M[4 * i + j] - P.center[i] * M[12 + j]) * P.inv_half[i];
    for (int j = 0; j < 4; ++j) p.M3[j] = M[12 + j];
    p.half_z = 1.0 / P.inv_half[2];
    p.center_z = P.center[2];
    p.eps = ctx->ambiguity_eps;
    p.H = H;
    p.W = W;
    p.xsize = ctx->aoi.xsize;
    p.ysize = ctx->aoi.ysize;
    p.degree = P.degree;
    p.warp_agg = ctx->k1_warp_agg ? 1 : 0;
    p.touched = ctx->cur_touched;
    p.tiles_x = (ctx->aoi.xsize + VS_TILE_W - 1) / VS_TILE_W;
    p.w_magic = W > 0 ? (((1ull << 48) + (unsigned long long)W - 1) / (unsigned long long)W) : 0;
    PolyCoefs& pc = *pc_out;
    memset(&pc, 0, sizeof(pc));
    if (P.degree > 0) {
        if (P.d64 > 0) {
            for (int o = 0; o < 3; ++o) {
                memcpy(pc.c64[o], P.coefLo[o], sizeof(P.coefLo[o]));
                memcpy(pc.c32[o], P.coefR[o], sizeof(P.coefR[o]));
            }
        } else {
            memcpy(pc.c64, P.coef, sizeof(pc.c64));
        }
    }
}

extern "C" {

int vs_unproject_rasterize(vs_ctx* ctx, const float* depth, int32_t H, int32_t W, const double* M, uint32_t* keygrid,
                           int clear_first, float* height_map, uint64_t* stats, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_unproject_rasterize: NULL context");
    if (!ctx->aoi_set) {
        vs_set_error("vs_unproject_rasterize: call vs_set_aoi first");
        return VS_ERR_STATE;
    }
    VS_REQUIRE(H >= 0 && W >= 0, "vs_unproject_rasterize: negative image size");
    VS_REQUIRE(W < 65536 && (int64_t)H * W < ((int64_t)1 << 32) - 8, "vs_unproject_rasterize: image too large (W < 65536, H*W < 2^32)");
    VS_REQUIRE(M != nullptr && keygrid != nullptr, "vs_unproject_rasterize: NULL argument");
    VS_REQUIRE(((uintptr_t)depth & 15) == 0, "vs_unproject_rasterize: depth must be 16-byte aligned");
    VS_REQUIRE(height_map == nullptr || ((uintptr_t)height_map & 15) == 0,
               "vs_unproject_rasterize: height_map must be 16-byte aligned");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    const VsPoly& P = ctx->poly;
    const size_t cells = (size_t)ctx->aoi.xsize * ctx->aoi.ysize;
    if (clear_first) VS_CUDA(cudaMemsetAsync(keygrid, 0, cells * sizeof(uint32_t), stream));
    unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(stats);
    if (d_stats) VS_CUDA(cudaMemsetAsync(d_stats, 0, VS_NUM_STATS * sizeof(uint64_t), stream));
    const int64_t n_pix = (int64_t)H * W;
    if (n_pix == 0) return VS_OK;
    VS_REQUIRE(depth != nullptr, "vs_unproject_rasterize: depth is NULL");

    RasterParams p;
    PolyCoefs pc;
    vs_make_raster_params(ctx, H, W, M, &p, &pc);

    VsEllipsoidConsts c = vs_make_ellipsoid_consts();
    if (P.degree > 0) {
        const VsExactParams* ex = reinterpret_cast<const VsExactParams*>(ctx->d_exact);
        // One wave of resident CTAs (3 per SM for the d64 == 1 variant, 2 otherwise; __launch_bounds__ above): longer
        // grid-stride loops keep the prefetch pipeline full and avoid a partial last wave (measured on C2: 3.24 ms per
        // step with 3 CTAs per SM, 3.27 / 3.30 / 3.34 with 6 / 8 / 12).  VISSAT_K1_CTAS_PER_SM overrides.
        static const int k1_ctas_env = []() {
            const char* e = getenv("VISSAT_K1_CTAS_PER_SM");
            const int v = e ? atoi(e) : 0;
            return v < 0 ? 0 : (v > 32 ? 32 : v);
        }();
        const int grid = persistent_grid(ctx, (n_pix + PX - 1) / PX, k1_ctas_env ? k1_ctas_env : (P.d64 == 1 ? 3 : 2));
#define VS_LAUNCH_K1(DEG, LO) \
    k_unproject_scatter<DEG, LO, false><<<grid, kThreads, 0, stream>>>(p, pc, ex, depth, keygrid, height_map, d_stats)
#define VS_LAUNCH_K1_LEAN(DEG, LO) \
    k_unproject_scatter<DEG, LO, true><<<grid, kThreads, 0, stream>>>(p, pc, ex, depth, keygrid, height_map, d_stats)
        // degree 3 (every AOI up to a few km) also exists as a lean variant for the plain throughput call
        const bool lean = P.degree == 3 && d_stats == nullptr && height_map == nullptr && p.touched == nullptr && !p.warp_agg;
        switch (lean ? P.d64 : P.degree * 3 + P.d64) {
            case 0: VS_LAUNCH_K1_LEAN(3, 0); break;
            case 1: VS_LAUNCH_K1_LEAN(3, 1); break;
            case 2: VS_LAUNCH_K1_LEAN(3, 2); break;
            case 9: VS_LAUNCH_K1(3, 0); break;
            case 10: VS_LAUNCH_K1(3, 1); break;
            case 11: VS_LAUNCH_K1(3, 2); break;
            case 12: VS_LAUNCH_K1(4, 0); break;
            case 13: VS_LAUNCH_K1(4, 1); break;
            case 14: VS_LAUNCH_K1(4, 2); break;
            case 15: VS_LAUNCH_K1(5, 0); break;
            case 16: VS_LAUNCH_K1(5, 1); break;
            case 17: VS_LAUNCH_K1(5, 2); break;
            default: vs_set_error("vs_unproject_rasterize: bad polynomial degree"); return VS_ERR_STATE;
        }
#undef VS_LAUNCH_K1
#undef VS_LAUNCH_K1_LEAN
        VS_CHECK_LAUNCH(ctx, "k_unproject_scatter");
    } else {
        if (ctx->cur_touched != nullptr)   // exact-chain kernel (polynomial disabled): no per-point marking, every tile counts
            VS_CUDA(cudaMemsetAsync(ctx->cur_touched, 1, (size_t)p.tiles_x * ((ctx->aoi.ysize + VS_TILE_H - 1) / VS_TILE_H), stream));
        const int grid = persistent_grid(ctx, n_pix, 4);
        k_unproject_scatter_exact<<<grid, kThreads, 0, stream>>>(p, c, ctx->geo, P.center[0], 1.0 / P.inv_half[0],
                                                                P.center[1], 1.0 / P.inv_half[1], depth, keygrid,
                                                                height_map, d_stats);
        VS_CHECK_LAUNCH(ctx, "k_unproject_scatter_exact");
    }
    return VS_OK;
}

int vs_keygrid_clear(vs_ctx* ctx, void* keygrid, int64_t n_keys, int key_bytes, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_keygrid_clear: NULL context");
    VS_REQUIRE(n_keys >= 0 && (key_bytes == 4 || key_bytes == 8), "vs_keygrid_clear: bad size");
    if (n_keys == 0) return VS_OK;
    VS_REQUIRE(keygrid != nullptr, "vs_keygrid_clear: NULL keygrid");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    VS_CUDA(cudaMemsetAsync(keygrid, 0, (size_t)n_keys * key_bytes, (cudaStream_t)stream_));
    return VS_OK;
}

int vs_points_rasterize(vs_ctx* ctx, const double* points, int64_t n_points, double xoff, double yoff,
                        double xresolution, double yresolution, int32_t xsize, int32_t ysize, uint64_t* keygrid64,
                        int clear_first, uint64_t* stats, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_points_rasterize: NULL context");
    VS_REQUIRE(n_points >= 0, "vs_points_rasterize: negative point count");
    VS_REQUIRE(xsize > 0 && ysize > 0, "vs_points_rasterize: grid size must be positive");
    VS_REQUIRE(keygrid64 != nullptr, "vs_points_rasterize: keygrid is NULL");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (clear_first) VS_CUDA(cudaMemsetAsync(keygrid64, 0, (size_t)xsize * ysize * sizeof(uint64_t), stream));
    if (stats) VS_CUDA(cudaMemsetAsync(stats, 0, VS_NUM_STATS * sizeof(uint64_t), stream));
    if (n_points == 0) return VS_OK;
    VS_REQUIRE(points != nullptr, "vs_points_rasterize: points is NULL");
    const int grid = persistent_grid(ctx, n_points, 8);
    k_points_scatter<<<grid, kThreads, 0, stream>>>(points, n_points, xoff, yoff, xresolution, yresolution, xsize, ysize,
                                                    reinterpret_cast<unsigned long long*>(keygrid64),
                                                    reinterpret_cast<unsigned long long*>(stats));
    VS_CHECK_LAUNCH(ctx, "k_points_scatter");
    return VS_OK;
}

}  // extern "C"
