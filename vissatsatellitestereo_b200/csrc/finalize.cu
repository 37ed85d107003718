// Stage B kernels: key grid -> per-view DSM, and the stand-alone 3x3 median.
//
// K2  k_grid_finalize<Key>   lib/proj_to_grid.py:62-79 (decode, NaN-hole fill from the PRE-fill grid) fused with
//     produce_dsm.py:58 (astype(float32) + cv2.medianBlur(.,3)).  One CTA (256 threads) per 64x32 output tile.
//     The decoded tile with a 2-cell halo lives in shared memory, so each key is read from global memory once
//     per tile (halo re-reads are L2 hits) and each output is written once (16-byte stores): algorithmic traffic
//     4 B read + 4 B write per cell.  The kernel is bound by the ALU pipe (min/max and integer instructions
//     issue at half rate on sm_100), so the work per cell is what is optimised:
//       * holes are listed with one shared-memory atomic each and filled densely by the first n_holes threads;
//       * every thread blurs a 4-wide x 2-tall patch: 6 sorted 3-columns are shared by 4 windows
//         (21 min/max per output instead of 30), rows are fetched with 16-byte shared-memory loads;
//       * a tile without any NaN left skips every NaN test; otherwise each window with a NaN goes through the
//         exact emulation of OpenCV's 19-exchange network (SIMD / scalar column semantics).
// K4  k_median3x3            aggregate_2p5d.py:81 on a row band (halo rows supplied by the caller); same blur.
#include <stdlib.h>

#include "finalize_tile.cuh"

using namespace vsfin;

namespace {

template <typename Key, typename Sink>
__global__ void __launch_bounds__(kThreads, sizeof(Key) == 4 ? 5 : 1)   // float32 keys: 48 registers, 5 CTAs per SM
k_grid_finalize(const Key* __restrict__ keygrid, int W, int H, typename KeyTraits<Key>::value_t* __restrict__ filled_out,
                float* __restrict__ blur_out, int simd_cols, unsigned long long* __restrict__ nan_count,
                const __grid_constant__ Sink sink, Key* __restrict__ clear_other) {
    // vs_views_to_dsm alternates between two key grids per internal stream: while this launch consumes one, every CTA
    // zeroes its own 64x32 tile of the OTHER one (consumed by the previous launch of the stream, scattered into by the
    // next stage A), which replaces the whole-grid memset per view.  Tiles partition the grid, so no halo is involved.
    if (clear_other != nullptr) {
        const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
        if (sizeof(Key) == 4 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(clear_other) & 15) == 0 && tx0 + TW <= W) {
            static_assert(TW == 64 && TH == 32 && kThreads == 256, "two 16-byte stores per thread: rows r and r + 16");
            const int r = threadIdx.x >> 4, c4 = threadIdx.x & 15;
            Key* q = clear_other + (size_t)(ty0 + r) * W + tx0 + 4 * c4;
            if (ty0 + r < H) *reinterpret_cast<uint4*>(q) = make_uint4(0u, 0u, 0u, 0u);
            if (ty0 + r + 16 < H) *reinterpret_cast<uint4*>(q + (size_t)16 * W) = make_uint4(0u, 0u, 0u, 0u);
        } else {
            for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
                const int r = i / TW, c = i - r * TW;
                if (ty0 + r < H && tx0 + c < W) clear_other[(size_t)(ty0 + r) * W + tx0 + c] = 0;
            }
        }
    }
    grid_finalize_tile<Key, Sink>(blockIdx.x, blockIdx.y, keygrid, W, H, filled_out, blur_out, simd_cols, nan_count, sink);
}

// in: rows [in_row0, ...) of an H_total-row image; writes rows [row_begin, row_end)
__global__ void __launch_bounds__(kThreads)
k_median3x3(const float* __restrict__ in, int in_row0, int H, int W, int row_begin, int row_end,
            float* __restrict__ out, int simd_cols, unsigned long long* __restrict__ nan_count) {
    __shared__ __align__(16) float s_fill[TR * TS];
    __shared__ int s_has_nan;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = row_begin + blockIdx.y * TH;
    if (tid == 0) s_has_nan = 0;
    __syncthreads();
    bool saw_nan = false;
    const bool vec_ok = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && tx0 >= 4 && tx0 + TW + 4 <= W &&
                        ty0 >= 1 && ty0 + TH <= row_end && ty0 + TH <= H - 1;
    if (vec_ok) {   // interior tile: rows ty0-1 .. ty0+TH, grid columns tx0-4 .. tx0+67, 16-byte loads
        constexpr int VPR = TS / 4;
        constexpr int NROW = TH + 2;
        constexpr int NV4 = (NROW * VPR + kThreads - 1) / kThreads;
        float4 fv[NV4];
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            const int i = tid + q * kThreads;
            const int r = 1 + i / VPR, v4 = i - (r - 1) * VPR;
            fv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < NROW * VPR)
                fv[q] = *reinterpret_cast<const float4*>(in + (size_t)(ty0 - 2 + r - in_row0) * W + (tx0 - 4) + 4 * v4);
        }
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            const int i = tid + q * kThreads;
            const int r = 1 + i / VPR, v4 = i - (r - 1) * VPR;
            if (i < NROW * VPR) {
                *reinterpret_cast<float4*>(s_fill + r * TS + 4 * v4) = fv[q];
                const float sum = (fv[q].x + fv[q].y) + (fv[q].z + fv[q].w);
                saw_nan |= !(sum == sum);       // conservative (also the 2 unused columns on each side)
            }
        }
    } else {
        const int warp = tid >> 5, lane = tid & 31;
        constexpr int NIT = (TH + 2 + 7) / 8;
        float vals[NIT][3];
        // issue every global load first (memory-level parallelism), then store to shared memory
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int r = 1 + warp + 8 * it;
            const int gy = ty0 - 2 + r;
            // rows the caller supplies: max(row_begin-1, 0) .. min(row_end, H-1)
            const bool row_ok = r < TH + 3 && (unsigned)gy < (unsigned)H && gy >= row_begin - 1 && gy <= row_end;
            const float* __restrict__ row = in + (size_t)(row_ok ? gy - in_row0 : 0) * W + (tx0 - 2);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int c = 1 + lane + 32 * part;
                const int gx = tx0 - 2 + c;
                float v = CUDART_NAN_F;
                if (row_ok && (part < 2 || lane < TW + 2 - 64) && (unsigned)gx < (unsigned)W) v = row[c];
                vals[it][part] = v;
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int r = 1 + warp + 8 * it;
            if (r < TH + 3) {
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const int c = 1 + lane + 32 * part;
                    if (part < 2 || lane < TW + 2 - 64) {
                        const float v = vals[it][part];
                        const int gy = ty0 - 2 + r, gx = tx0 - 2 + c;
                        const bool inside = (unsigned)gy < (unsigned)H && gy >= row_begin - 1 && gy <= row_end &&
                                            (unsigned)gx < (unsigned)W;
                        saw_nan |= inside && (v != v);
                        s_fill[r * TS + OFF + c] = v;
                    }
                }
            }
        }
    }
    if (saw_nan) s_has_nan = 1;
    __syncthreads();
    replicate_border(s_fill, ty0, tx0, H, W);
    __syncthreads();
    const unsigned n_nan = blur_tile(s_fill, ty0, tx0, H, W, row_end, s_has_nan != 0, simd_cols != 0, out, row_begin);
    block_count_flush(n_nan, nan_count);
}

}  // namespace

// Two bit-identical implementations of stage B for float32 keys: the float-space kernels of this file
// (k_grid_finalize<uint32_t, .>, round 1) and the key-space kernels of finalize_keys.cu (round 2).  Measured: equal on
// C2 views (31.3 vs 31.5 us), the float-space one faster where holes dominate (C5: 137 vs 152 us, C4: 25.3 vs 26.3).
// So: dense calls use the float-space kernel; the sparse mode (occupancy bitmap / sparse exchange), which needs the
// touched-tile early-out, uses the key-space one.  VISSAT_K2_LEGACY=1 / VISSAT_K2_KEYS=1 force one of them.
int vs_launch_finalize_keys(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_cols,
                            unsigned long long* nan_count, cudaStream_t stream);
int vs_launch_finalize_keys_occ(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_cols,
                                unsigned long long* nan_count, const VsOccPlan& plan, const unsigned char* touched,
                                cudaStream_t stream);
int vs_launch_finalize_keys_peer(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_cols,
                                 unsigned long long* nan_count, const VsPeerPlan& plan, const unsigned char* touched,
                                 cudaStream_t stream);

// Stage B of one view with the peer stores of the multi-GPU exchange (called by vs_views_to_dsm, pipeline.cu).
int vs_grid_finalize_peer(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_lanes,
                          uint64_t* nan_count, const VsPeerPlan& plan, cudaStream_t stream, uint32_t* clear_other) {
    if (nan_count) VS_CUDA(cudaMemsetAsync(nan_count, 0, sizeof(uint64_t), stream));
    if (ctx->k2_mode == 2 || (ctx->k2_mode == 0 && plan.occ_words > 0))     // sparse exchange: key-space kernel
        return vs_launch_finalize_keys_peer(ctx, keygrid, xsize, ysize, dsm_out, simd_cols_for(xsize, simd_lanes),
                                            reinterpret_cast<unsigned long long*>(nan_count), plan, ctx->cur_touched, stream);
    PeerSink sink;
    sink.p = plan;
    dim3 grid((xsize + TW - 1) / TW, (ysize + TH - 1) / TH);
    k_grid_finalize<uint32_t, PeerSink><<<grid, kThreads, 0, stream>>>(keygrid, xsize, ysize, nullptr, dsm_out,
                                                                      simd_cols_for(xsize, simd_lanes),
                                                                      reinterpret_cast<unsigned long long*>(nan_count),
                                                                      sink, clear_other);
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize<u32, peer>");
    return VS_OK;
}

// Stage B of one view that also records the tile occupancy of the plane (vs_set_occupancy; called by vs_views_to_dsm).
int vs_grid_finalize_occ(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_lanes,
                         uint64_t* nan_count, const VsOccPlan& plan, cudaStream_t stream) {
    if (nan_count) VS_CUDA(cudaMemsetAsync(nan_count, 0, sizeof(uint64_t), stream));
    if (ctx->k2_mode != 1)                                                  // occupancy marking: key-space kernel
        return vs_launch_finalize_keys_occ(ctx, keygrid, xsize, ysize, dsm_out, simd_cols_for(xsize, simd_lanes),
                                           reinterpret_cast<unsigned long long*>(nan_count), plan, ctx->cur_touched, stream);
    OccSink sink;
    sink.o = plan;
    dim3 grid((xsize + TW - 1) / TW, (ysize + TH - 1) / TH);
    k_grid_finalize<uint32_t, OccSink><<<grid, kThreads, 0, stream>>>(keygrid, xsize, ysize, nullptr, dsm_out,
                                                                     simd_cols_for(xsize, simd_lanes),
                                                                     reinterpret_cast<unsigned long long*>(nan_count),
                                                                     sink, nullptr);
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize<u32, occ>");
    return VS_OK;
}

int vs_finalize_tma_try(vs_ctx* ctx, bool keys, const void* in, int in_rows, int in_row0, int W, int H, int row_begin,
                        int row_end, float* out, int simd_cols, unsigned long long* nan_count, cudaStream_t stream);

// Stage B of one view; clear_other: see k_grid_finalize (nullptr: nothing to zero).
int vs_grid_finalize_impl(vs_ctx* ctx, const uint32_t* keygrid, int32_t xsize, int32_t ysize, float* dsm_out, int simd_lanes,
                          uint64_t* nan_count, void* stream_, uint32_t* clear_other) {
    VS_REQUIRE(ctx != nullptr, "vs_grid_finalize: NULL context");
    VS_REQUIRE(xsize > 0 && ysize > 0, "vs_grid_finalize: grid size must be positive");
    VS_REQUIRE(keygrid != nullptr && dsm_out != nullptr, "vs_grid_finalize: NULL array");
    VS_REQUIRE(simd_lanes >= 0, "vs_grid_finalize: simd_lanes must be >= 0");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nan_count) VS_CUDA(cudaMemsetAsync(nan_count, 0, sizeof(uint64_t), stream));
    if (!ctx->no_tma && clear_other == nullptr) {   // opt-in: persistent CTAs fed by TMA (finalize_tma.cu)
        const int r = vs_finalize_tma_try(ctx, true, keygrid, ysize, 0, xsize, ysize, 0, ysize, dsm_out,
                                          simd_cols_for(xsize, simd_lanes),
                                          reinterpret_cast<unsigned long long*>(nan_count), stream);
        if (r < 0) return VS_ERR_CUDA;
        if (r > 0) return VS_OK;
    }
    if (ctx->k2_mode == 2 && clear_other == nullptr)                        // dense: the float-space kernel below is the default
        return vs_launch_finalize_keys(ctx, keygrid, xsize, ysize, dsm_out, simd_cols_for(xsize, simd_lanes),
                                       reinterpret_cast<unsigned long long*>(nan_count), stream);
    dim3 grid((xsize + TW - 1) / TW, (ysize + TH - 1) / TH);
    k_grid_finalize<uint32_t, NoSink><<<grid, kThreads, 0, stream>>>(keygrid, xsize, ysize, nullptr, dsm_out,
                                                                    simd_cols_for(xsize, simd_lanes),
                                                                    reinterpret_cast<unsigned long long*>(nan_count),
                                                                    NoSink(), clear_other);
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize<u32>");
    return VS_OK;
}

extern "C" {

int vs_grid_finalize(vs_ctx* ctx, const uint32_t* keygrid, int32_t xsize, int32_t ysize, float* dsm_out, int simd_lanes,
                     uint64_t* nan_count, void* stream_) {
    return vs_grid_finalize_impl(ctx, keygrid, xsize, ysize, dsm_out, simd_lanes, nan_count, stream_, nullptr);
}

int vs_grid_finalize64(vs_ctx* ctx, const uint64_t* keygrid64, int32_t xsize, int32_t ysize, double* filled64,
                       float* blurred32, int simd_lanes, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_grid_finalize64: NULL context");
    VS_REQUIRE(xsize > 0 && ysize > 0, "vs_grid_finalize64: grid size must be positive");
    VS_REQUIRE(keygrid64 != nullptr && (filled64 != nullptr || blurred32 != nullptr), "vs_grid_finalize64: NULL array");
    VS_REQUIRE(simd_lanes >= 0, "vs_grid_finalize64: simd_lanes must be >= 0");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    dim3 grid((xsize + TW - 1) / TW, (ysize + TH - 1) / TH);
    k_grid_finalize<unsigned long long, NoSink><<<grid, kThreads, 0, stream>>>(
        reinterpret_cast<const unsigned long long*>(keygrid64), xsize, ysize, filled64, blurred32,
        simd_cols_for(xsize, simd_lanes), nullptr, NoSink(), nullptr);
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize<u64>");
    return VS_OK;
}

int vs_median3x3(vs_ctx* ctx, const float* in, int32_t in_row0, int32_t H_total, int32_t W, int32_t row_begin,
                 int32_t row_end, float* out, int simd_lanes, uint64_t* nan_count, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_median3x3: NULL context");
    VS_REQUIRE(H_total > 0 && W > 0, "vs_median3x3: image size must be positive");
    VS_REQUIRE(row_begin >= 0 && row_end <= H_total && row_begin <= row_end, "vs_median3x3: bad row range");
    VS_REQUIRE(in_row0 >= 0 && in_row0 <= (row_begin > 0 ? row_begin - 1 : 0), "vs_median3x3: input must start at or before row_begin-1");
    VS_REQUIRE(simd_lanes >= 0, "vs_median3x3: simd_lanes must be >= 0");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nan_count) VS_CUDA(cudaMemsetAsync(nan_count, 0, sizeof(uint64_t), stream));
    if (row_begin == row_end) return VS_OK;
    VS_REQUIRE(in != nullptr && out != nullptr, "vs_median3x3: NULL array");
    if (!ctx->no_tma) {
        const int in_rows = (row_end + 1 < H_total ? row_end + 1 : H_total) - in_row0;
        const int r = vs_finalize_tma_try(ctx, false, in, in_rows, in_row0, W, H_total, row_begin, row_end, out,
                                          simd_cols_for(W, simd_lanes),
                                          reinterpret_cast<unsigned long long*>(nan_count), stream);
        if (r < 0) return VS_ERR_CUDA;
        if (r > 0) return VS_OK;
    }
    dim3 grid((W + TW - 1) / TW, (row_end - row_begin + TH - 1) / TH);
    k_median3x3<<<grid, kThreads, 0, stream>>>(in, in_row0, H_total, W, row_begin, row_end, out,
                                               simd_cols_for(W, simd_lanes),
                                               reinterpret_cast<unsigned long long*>(nan_count));
    VS_CHECK_LAUNCH(ctx, "k_median3x3");
    return VS_OK;
}

}  // extern "C"
