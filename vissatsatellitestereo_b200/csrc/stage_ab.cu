// Stages A and B co-scheduled in ONE kernel per view step: k_stage_ab.
//
// Stage A of a view (K1: aggregate_2p5d_util.py:75-98 + lib/proj_to_grid.py:42-61) is bound by the XU / FP64 / FMA
// pipes, stage B (K2: lib/proj_to_grid.py:62-79 + produce_dsm.py:58) by the ALU pipe.  As separate kernels they can
// only share the GPU SM by SM (a resident wave of K1 takes 94 % of the register file), so the pair runs at ~62 % of
// the issue rate.  Here every CTA of a persistent grid alternates between the two kinds of work, taken from two
// device-side queues:
//     launch j of an internal stream:   stage B of view j-1 (64x32 tiles of its key grid)
//                                     + stage A of view j   (batches of 4-pixel chunks of its depth map)
//                                     + zeroing of the key grid that launch j+1 will scatter into
// so that at any time the warps of an SM are spread over both instruction mixes.  Stage B of a view needs ALL of its
// stage A, which is why the two roles of one launch belong to consecutive views (the kernel boundary is the
// dependency), and three key grids rotate per internal stream.  The device code of both roles is the code of the
// separate kernels (rasterize_common.cuh: scatter_chunk_lean, finalize_tile.cuh: grid_finalize_tile); the results
// are bit-identical and the tests compare them.
//
// The depth of a batch is requested with cp.async BEFORE the CTA works on its tile and consumed after it, so the
// global-load latency of stage A hides behind stage B.
#include <stdlib.h>

#include "finalize_tile.cuh"
#include "rasterize_common.cuh"

namespace {

constexpr int kABThreads = 256;
constexpr int kABCtasPerSm = 3;
constexpr int kRing = 3;          // cp.async slots per thread

struct StageABParams {
    // stage A of view j (depth == nullptr: none)
    const float* depth;
    uint32_t* kg_scatter;
    int n_batches, iters;         // a batch = iters x 256 chunks of 4 pixels
    // stage B of view j-1 (kg_final == nullptr: none)
    const uint32_t* kg_final;
    float* dsm_out;
    unsigned long long* nan_count;
    int simd_cols, tiles_x, n_tiles;
    // key grid to zero for launch j+1 (nullptr: none)
    uint32_t* kg_clear;
    unsigned n_cells;
    int* counters;                // [0] next tile, [1] next batch (zeroed by the host before the launch)
};

template <int D, int D64, typename Sink>
__global__ void __launch_bounds__(kABThreads, kABCtasPerSm)
k_stage_ab(const __grid_constant__ StageABParams ab, const __grid_constant__ vsras::RasterParams p,
           const __grid_constant__ vsras::PolyCoefs pc, const VsExactParams* __restrict__ ex,
           const __grid_constant__ Sink sink) {
    using namespace vsras;
    __shared__ __align__(16) float4 s_pre[kRing][kABThreads];
    __shared__ int s_work[2];
    const int tid = threadIdx.x;

    // ---- zero the key grid of the next launch (fire-and-forget stores)
    if (ab.kg_clear != nullptr) {
        const unsigned n4 = ab.n_cells >> 2;
        uint4* __restrict__ c4 = reinterpret_cast<uint4*>(ab.kg_clear);
        for (unsigned i = blockIdx.x * kABThreads + tid; i < n4; i += gridDim.x * kABThreads) c4[i] = make_uint4(0u, 0u, 0u, 0u);
        if (blockIdx.x == 0 && tid < (int)(ab.n_cells & 3u)) ab.kg_clear[(n4 << 2) + tid] = 0u;
    }

    const int64_t n_pix = (int64_t)p.H * p.W;
    const unsigned n_chunks = ab.depth != nullptr ? (unsigned)((n_pix + PX - 1) / PX) : 0u;
    const unsigned n_full = (unsigned)(n_pix / PX);
    const bool simple = (p.W % PX) == 0;
    const float4* __restrict__ depth4 = reinterpret_cast<const float4*>(ab.depth);
    const unsigned slot_base = (unsigned)__cvta_generic_to_shared(&s_pre[0][tid]);
    constexpr unsigned kSlotStride = (unsigned)(sizeof(float4) * kABThreads);
    const int n_tiles = ab.kg_final != nullptr ? ab.n_tiles : 0;
    const int n_batches = ab.depth != nullptr ? ab.n_batches : 0;
    const unsigned batch_chunks = (unsigned)ab.iters * kABThreads;

    bool tiles_left = n_tiles > 0, batches_left = n_batches > 0;
    bool first = true;
    while (tiles_left || batches_left) {
        if (tid == 0) {
            // odd CTAs start with a batch only, so that the CTAs of an SM are out of phase from the first iteration on
            const bool skip_tile = first && (blockIdx.x & 1) && batches_left;
            s_work[0] = tiles_left ? (skip_tile ? -1 : atomicAdd(&ab.counters[0], 1)) : n_tiles;
            s_work[1] = batches_left ? atomicAdd(&ab.counters[1], 1) : n_batches;
        }
        first = false;
        __syncthreads();
        const int tile = s_work[0], batch = s_work[1];
        tiles_left = tile < n_tiles;
        batches_left = batch < n_batches;
        const bool do_tile = tile >= 0 && tiles_left;

        // ---- stage A, part 1: request the depth of the first two iterations of the batch
        const unsigned c0 = (unsigned)batch * batch_chunks + tid;
        if (batches_left) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const unsigned chunk = c0 + it * kABThreads;
                if (simple && it < ab.iters && chunk < n_full) cp_async16_s(slot_base + it * kSlotStride, depth4 + chunk);
                cp_async_commit();
            }
        }

        // ---- stage B: one tile of the previous view
        if (do_tile) {
            const int by = tile / ab.tiles_x, bx = tile - by * ab.tiles_x;
            vsfin::grid_finalize_tile<uint32_t, Sink>(bx, by, ab.kg_final, p.xsize, p.ysize, nullptr, ab.dsm_out, ab.simd_cols,
                                                      ab.nan_count, sink);
        }

        // ---- stage A, part 2: the batch
        if (batches_left) {
            int slot = 0;
            for (int it = 0; it < ab.iters; ++it) {
                const unsigned chunk = c0 + (unsigned)it * kABThreads;
                {
                    const int s2 = slot >= 1 ? slot - 1 : slot + 2;          // (slot + 2) % 3
                    const unsigned nxt = chunk + 2u * kABThreads;
                    if (simple && it + 2 < ab.iters && nxt < n_full) cp_async16_s(slot_base + s2 * kSlotStride, depth4 + nxt);
                    cp_async_commit();
                }
                asm volatile("cp.async.wait_group 2;" ::: "memory");
                const float4 cur = ld_shared_f4(slot_base + slot * kSlotStride);
                slot = slot == kRing - 1 ? 0 : slot + 1;
                if (chunk >= n_chunks) continue;
                const unsigned base = chunk * PX;
                // row = base / W by multiply-shift (magic = ceil(2^48 / W), exact for base < 2^32, W < 2^16)
                const unsigned row0 = (unsigned)(((unsigned long long)base * p.w_magic) >> 48);
                const unsigned col0 = base - row0 * (unsigned)p.W;
                if (simple ? (chunk >= n_full) : (col0 + PX > (unsigned)p.W || (int64_t)base + PX > n_pix)) {
                    unsigned cnt[4] = {0, 0, 0, 0};
                    scatter_chunk_generic<D, D64>(p, pc, ex, ab.depth, base, n_pix, ab.kg_scatter, nullptr, false, cnt);
                    continue;
                }
                const float4 t4 = simple ? cur : ld_stream_f4(reinterpret_cast<const float4*>(ab.depth + base));
                scatter_chunk_lean<D, D64>(p, pc, ex, t4, row0, col0, ab.kg_scatter);
            }
        }
        __syncthreads();   // s_work and the tile's shared arrays are reused by the next iteration
    }
}

template <int D, int D64, typename Sink>
int launch_ab(vs_ctx* ctx, const StageABParams& ab, const vsras::RasterParams& p, const vsras::PolyCoefs& pc, const Sink& sink,
              int grid, cudaStream_t stream) {
    const VsExactParams* ex = reinterpret_cast<const VsExactParams*>(ctx->d_exact);
    k_stage_ab<D, D64, Sink><<<grid, kABThreads, 0, stream>>>(ab, p, pc, ex, sink);
    VS_CHECK_LAUNCH(ctx, "k_stage_ab");
    return VS_OK;
}

}  // namespace

bool vs_peer_plan_for(const vs_ctx* ctx, const float* plane, int64_t plane_stride, VsPeerPlan* plan);

// Can this batch of views take the co-scheduled path?  (degree-3 polynomial with a float32 part, dense mode, no per-view
// counters / timing requested, TMA variant off.)  vs_set_coschedule(ctx, 0) / VISSAT_AB=0 switch the path off.
bool vs_stage_ab_eligible(const vs_ctx* ctx, bool want_stats, bool sparse) {
    if (!ctx->ab_on || want_stats || sparse || ctx->timing || !ctx->no_tma || ctx->k1_warp_agg || ctx->k2_mode == 2) return false;
    if (ctx->occ != nullptr) return false;
    if (ctx->xch_on && ctx->xch.occ_words > 0) return false;
    return ctx->poly.degree == 3 && (ctx->poly.d64 == 1 || ctx->poly.d64 == 2);
}

// Stages A + B of n_views views on `ns` internal streams (fork from / join into `stream` is done by the caller,
// pipeline.cu): stream s takes views s, s + ns, ...; launch j of a stream = stage B of its view j-1 + stage A of its view
// j.  Key grids: ctx->d_keygrid_ab[s][0..2], rotating.
int vs_stage_ab_run(vs_ctx* ctx, int n_views, const float* const* depth, const int32_t* H, const int32_t* W,
                    const double* inv_proj_mats, float* dsm_stack, int64_t plane_stride, int simd_lanes,
                    uint64_t* nan_counts, int ns, cudaStream_t* streams) {
    const int xs = ctx->aoi.xsize, ys = ctx->aoi.ysize;
    const size_t cells = (size_t)xs * ys;
    const int tiles_x = (xs + vsfin::TW - 1) / vsfin::TW, tiles_y = (ys + vsfin::TH - 1) / vsfin::TH;
    const int n_tiles = tiles_x * tiles_y;
    static const int ctas_env = []() {
        const char* e = getenv("VISSAT_AB_CTAS_PER_SM");
        const int v = e ? atoi(e) : 0;
        return v < 1 ? 0 : (v > kABCtasPerSm ? kABCtasPerSm : v);
    }();
    const int grid = ctx->sm_count * (ctas_env ? ctas_env : kABCtasPerSm);
    // device counters: 2 per launch, zeroed once per call on every stream's first use (the caller zeroes the whole array)
    int launch_no = 0;
    for (int s = 0; s < ns; ++s) {
        cudaStream_t st = streams[s];
        uint32_t** kg = ctx->d_keygrid_ab[s];
        int n_s = 0;
        for (int v = s; v < n_views; v += ns) ++n_s;
        if (n_s == 0) continue;
        VS_CUDA(cudaMemsetAsync(kg[0], 0, cells * sizeof(uint32_t), st));
        for (int j = 0; j <= n_s; ++j, ++launch_no) {
            StageABParams ab;
            memset(&ab, 0, sizeof(ab));
            vsras::RasterParams p;
            vsras::PolyCoefs pc;
            ab.counters = ctx->d_ab_counters + 2 * launch_no;
            ab.n_cells = (unsigned)cells;
            ab.tiles_x = tiles_x;
            ab.n_tiles = n_tiles;
            ab.simd_cols = vsfin::simd_cols_for(xs, simd_lanes);
            const int va = s + j * ns, vb = s + (j - 1) * ns;
            bool have_a = false;
            if (j < n_s) {
                const int64_t n_pix = (int64_t)H[va] * W[va];
                VS_REQUIRE(H[va] >= 0 && W[va] >= 0, "vs_views_to_dsm: negative image size");
                VS_REQUIRE(W[va] < 65536 && n_pix < ((int64_t)1 << 32) - 8, "vs_views_to_dsm: image too large (W < 65536, H*W < 2^32)");
                VS_REQUIRE(n_pix == 0 || depth[va] != nullptr, "vs_views_to_dsm: depth is NULL");
                VS_REQUIRE(((uintptr_t)depth[va] & 15) == 0, "vs_views_to_dsm: depth must be 16-byte aligned");
                vs_make_raster_params(ctx, H[va], W[va], inv_proj_mats + 16 * (size_t)va, &p, &pc);
                if (n_pix > 0) {
                    const int64_t n_chunks = (n_pix + vsras::PX - 1) / vsras::PX;
                    // as many batches as there are tiles (both queues drain together), but at least 4 per CTA
                    const int64_t want = n_tiles > 4 * grid ? n_tiles : 4 * (int64_t)grid;
                    int64_t iters = (n_chunks + want * kABThreads - 1) / (want * kABThreads);
                    if (iters < 1) iters = 1;
                    if (iters > 16) iters = 16;
                    ab.iters = (int)iters;
                    ab.n_batches = (int)((n_chunks + iters * kABThreads - 1) / (iters * kABThreads));
                    ab.depth = depth[va];
                    ab.kg_scatter = kg[j % 3];
                    have_a = true;
                }
            }
            if (!have_a) vs_make_raster_params(ctx, 0, 4, inv_proj_mats, &p, &pc);   // grid size fields only
            p.touched = nullptr;
            float* plane = nullptr;
            if (j >= 1) {
                plane = dsm_stack + (size_t)vb * plane_stride;
                ab.kg_final = kg[(j - 1) % 3];
                ab.dsm_out = plane;
                ab.nan_count = nan_counts ? reinterpret_cast<unsigned long long*>(nan_counts + vb) : nullptr;
                if (ab.nan_count) VS_CUDA(cudaMemsetAsync(ab.nan_count, 0, sizeof(uint64_t), st));
            }
            if (j + 1 < n_s) ab.kg_clear = kg[(j + 1) % 3];
            int rc;
            if (ctx->xch_on && j >= 1) {
                vsfin::PeerSink sink;
                if (!vs_peer_plan_for(ctx, plane, plane_stride, &sink.p)) {
                    vs_set_error("vs_views_to_dsm: dsm_stack plane is not a plane of vs_exchange.local_stack");
                    return VS_ERR_INVALID;
                }
                rc = ctx->poly.d64 == 1 ? launch_ab<3, 1>(ctx, ab, p, pc, sink, grid, st) : launch_ab<3, 2>(ctx, ab, p, pc, sink, grid, st);
            } else {
                vsfin::NoSink sink;
                rc = ctx->poly.d64 == 1 ? launch_ab<3, 1>(ctx, ab, p, pc, sink, grid, st) : launch_ab<3, 2>(ctx, ab, p, pc, sink, grid, st);
            }
            if (rc) return rc;
        }
    }
    return VS_OK;
}
