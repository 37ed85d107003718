// Stage B kernels: key grid -> per-view DSM, and the stand-alone 3x3 median.
//
// K2  k_grid_finalize<Key,T>   lib/proj_to_grid.py:62-79 (decode, NaN-hole fill from the PRE-fill grid) fused
//     with produce_dsm.py:58 (astype(float32) + cv2.medianBlur(.,3)).  One CTA per 32x32 output tile; the
//     decoded tile (+2 halo) and the hole-filled tile (+1 halo) live in shared memory, so each key is read
//     from global memory once per tile (halo re-reads are L2 hits) and each output is written once:
//     algorithmic traffic 4 B read + 4 B write per cell.
// K4  k_median3x3              aggregate_2p5d.py:81 on a row band (halo rows supplied by the caller).
#include <math_constants.h>

#include "median.cuh"
#include "vs_common.cuh"

namespace {

constexpr int TILE = 32;
constexpr int kThreads = 256;

template <typename Key> struct KeyTraits;
template <> struct KeyTraits<uint32_t> {
    typedef float value_t;
    static __device__ __forceinline__ float decode(uint32_t k) { return vs_unkey32(k); }
};
template <> struct KeyTraits<unsigned long long> {
    typedef double value_t;
    static __device__ __forceinline__ double decode(unsigned long long k) { return vs_unkey64(k); }
};

__device__ __forceinline__ void block_count_flush(unsigned local, unsigned long long* counter) {
    if (counter == nullptr) return;
    __shared__ unsigned s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    unsigned r = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && r) atomicAdd(&s_cnt, r);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(counter, (unsigned long long)s_cnt);
}

// blur of one output from a shared-memory tile of already-filled float values.
// `at(r, c)` returns the filled value at GLOBAL (r, c) (must be inside the tile's halo).
template <typename At>
__device__ __forceinline__ float blur_at(const At& at, int gy, int gx, int H, int W, bool simd_cols) {
    if (H == 1 || W == 1) {
        if (H == 1) return vs_median3_line(at(gy, max(gx - 1, 0)), at(gy, gx), at(gy, min(gx + 1, W - 1)));
        return vs_median3_line(at(max(gy - 1, 0), gx), at(gy, gx), at(min(gy + 1, H - 1), gx));
    }
    const int r0 = max(gy - 1, 0), r2 = min(gy + 1, H - 1);
    const int c0 = max(gx - 1, 0), c2 = min(gx + 1, W - 1);
    const float p0 = at(r0, c0), p1 = at(r0, gx), p2 = at(r0, c2);
    const float p3 = at(gy, c0), p4 = at(gy, gx), p5 = at(gy, c2);
    const float p6 = at(r2, c0), p7 = at(r2, gx), p8 = at(r2, c2);
    const bool simd = simd_cols && gx >= 1 && gx <= W - 2;
    return simd ? vs_median9_net<true>(p0, p1, p2, p3, p4, p5, p6, p7, p8)
                : vs_median9_net<false>(p0, p1, p2, p3, p4, p5, p6, p7, p8);
}

template <typename Key>
__global__ void __launch_bounds__(kThreads)
k_grid_finalize(const Key* __restrict__ keygrid, int W, int H, typename KeyTraits<Key>::value_t* __restrict__ filled_out,
                float* __restrict__ blur_out, int simd_cols, unsigned long long* __restrict__ nan_count) {
    typedef typename KeyTraits<Key>::value_t T;
    constexpr int RW = TILE + 4;  // raw tile width (halo 2)
    constexpr int FW = TILE + 2;  // filled tile width (halo 1)
    __shared__ T s_raw[RW * RW];
    __shared__ float s_fill[FW * FW];
    const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;

    // 1. decode keys (+2 halo); outside the grid -> NaN (hole fill only uses in-range neighbours, :73)
    for (int i = threadIdx.x; i < RW * RW; i += kThreads) {
        const int r = i / RW, c = i - r * RW;
        const int gy = ty0 - 2 + r, gx = tx0 - 2 + c;
        T v = (T)CUDART_NAN;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = KeyTraits<Key>::decode(keygrid[(size_t)gy * W + gx]);
        s_raw[i] = v;
    }
    __syncthreads();

    // 2. hole fill (+1 halo): NaN cell <- median of its non-NaN 3x3 neighbours in the pre-fill grid
    for (int i = threadIdx.x; i < FW * FW; i += kThreads) {
        const int r = i / FW, c = i - r * FW;
        const int gy = ty0 - 1 + r, gx = tx0 - 1 + c;
        const int rr = r + 1, rc = c + 1;  // position in s_raw
        T v = s_raw[rr * RW + rc];
        const bool inside = gy >= 0 && gy < H && gx >= 0 && gx < W;
        if (inside && v != v) {
            T nb[8] = {s_raw[(rr - 1) * RW + rc - 1], s_raw[(rr - 1) * RW + rc], s_raw[(rr - 1) * RW + rc + 1],
                       s_raw[rr * RW + rc - 1],                                   s_raw[rr * RW + rc + 1],
                       s_raw[(rr + 1) * RW + rc - 1], s_raw[(rr + 1) * RW + rc], s_raw[(rr + 1) * RW + rc + 1]};
            v = vs_median_of_valid8<T>(nb);
        }
        s_fill[i] = (float)v;  // produce_dsm.py:58 astype(np.float32)
        if (filled_out != nullptr && inside && r >= 1 && r <= TILE && c >= 1 && c <= TILE)
            filled_out[(size_t)gy * W + gx] = v;
    }
    if (blur_out == nullptr) return;
    __syncthreads();

    // 3. cv2.medianBlur(., 3) with replicated borders
    unsigned n_nan = 0;
    auto at = [&](int gy, int gx) -> float { return s_fill[(gy - (ty0 - 1)) * FW + (gx - (tx0 - 1))]; };
    for (int i = threadIdx.x; i < TILE * TILE; i += kThreads) {
        const int r = i / TILE, c = i - r * TILE;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy < H && gx < W) {
            const float m = blur_at(at, gy, gx, H, W, simd_cols != 0);
            blur_out[(size_t)gy * W + gx] = m;
            n_nan += (m != m);
        }
    }
    block_count_flush(n_nan, nan_count);
}

// in: rows [in_row0, ...) of an H_total-row image; writes rows [row_begin, row_end)
__global__ void __launch_bounds__(kThreads)
k_median3x3(const float* __restrict__ in, int in_row0, int H, int W, int row_begin, int row_end,
            float* __restrict__ out, int simd_cols, unsigned long long* __restrict__ nan_count) {
    constexpr int FW = TILE + 2;
    __shared__ float s_fill[FW * FW];
    const int tx0 = blockIdx.x * TILE, ty0 = row_begin + blockIdx.y * TILE;
    for (int i = threadIdx.x; i < FW * FW; i += kThreads) {
        const int r = i / FW, c = i - r * FW;
        const int gy = ty0 - 1 + r, gx = tx0 - 1 + c;
        float v = CUDART_NAN_F;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W && gy >= row_begin - 1 && gy <= row_end)
            v = in[(size_t)(gy - in_row0) * W + gx];
        s_fill[i] = v;
    }
    __syncthreads();
    unsigned n_nan = 0;
    auto at = [&](int gy, int gx) -> float { return s_fill[(gy - (ty0 - 1)) * FW + (gx - (tx0 - 1))]; };
    for (int i = threadIdx.x; i < TILE * TILE; i += kThreads) {
        const int r = i / TILE, c = i - r * TILE;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy < row_end && gx < W) {
            const float m = blur_at(at, gy, gx, H, W, simd_cols != 0);
            out[(size_t)(gy - row_begin) * W + gx] = m;
            n_nan += (m != m);
        }
    }
    block_count_flush(n_nan, nan_count);
}

inline int simd_cols_for(int W, int simd_lanes) { return (simd_lanes > 0 && W >= simd_lanes + 2) ? 1 : 0; }

}  // namespace

extern "C" {

int vs_grid_finalize(vs_ctx* ctx, const uint32_t* keygrid, int32_t xsize, int32_t ysize, float* dsm_out, int simd_lanes,
                     uint64_t* nan_count, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_grid_finalize: NULL context");
    VS_REQUIRE(xsize > 0 && ysize > 0, "vs_grid_finalize: grid size must be positive");
    VS_REQUIRE(keygrid != nullptr && dsm_out != nullptr, "vs_grid_finalize: NULL array");
    VS_REQUIRE(simd_lanes >= 0, "vs_grid_finalize: simd_lanes must be >= 0");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nan_count) VS_CUDA(cudaMemsetAsync(nan_count, 0, sizeof(uint64_t), stream));
    dim3 grid((xsize + TILE - 1) / TILE, (ysize + TILE - 1) / TILE);
    k_grid_finalize<uint32_t><<<grid, kThreads, 0, stream>>>(keygrid, xsize, ysize, nullptr, dsm_out,
                                                            simd_cols_for(xsize, simd_lanes),
                                                            reinterpret_cast<unsigned long long*>(nan_count));
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize<u32>");
    return VS_OK;
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
    dim3 grid((xsize + TILE - 1) / TILE, (ysize + TILE - 1) / TILE);
    k_grid_finalize<unsigned long long><<<grid, kThreads, 0, stream>>>(
        reinterpret_cast<const unsigned long long*>(keygrid64), xsize, ysize, filled64, blurred32,
        simd_cols_for(xsize, simd_lanes), nullptr);
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
    dim3 grid((W + TILE - 1) / TILE, (row_end - row_begin + TILE - 1) / TILE);
    k_median3x3<<<grid, kThreads, 0, stream>>>(in, in_row0, H_total, W, row_begin, row_end, out,
                                               simd_cols_for(W, simd_lanes),
                                               reinterpret_cast<unsigned long long*>(nan_count));
    VS_CHECK_LAUNCH(ctx, "k_median3x3");
    return VS_OK;
}

}  // extern "C"
