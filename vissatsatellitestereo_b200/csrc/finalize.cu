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
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    unsigned r = __reduce_add_sync(0xffffffffu, local);
    if ((tid & 31) == 0 && r) atomicAdd(&s_cnt, r);
    __syncthreads();
    if (tid == 0 && s_cnt) atomicAdd(counter, (unsigned long long)s_cnt);
}

// ---- 3x3 median on a shared-memory tile ----------------------------------------------------------------------
// For a window without NaN every correct median algorithm returns the same value as OpenCV's network, so the
// common case uses a cheap one (sort the three columns, then med3(max of minima, med3 of medians, min of
// maxima): 30 min/max).  Windows that contain a NaN go through the exact emulation of OpenCV's network.
__device__ __forceinline__ float med3f(float a, float b, float c) {
    return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}

__device__ __forceinline__ float median9_fast(float p0, float p1, float p2, float p3, float p4, float p5, float p6,
                                              float p7, float p8) {
    // columns (p0,p3,p6) (p1,p4,p7) (p2,p5,p8)
    const float a_lo = fminf(fminf(p0, p3), p6), a_hi = fmaxf(fmaxf(p0, p3), p6), a_mid = med3f(p0, p3, p6);
    const float b_lo = fminf(fminf(p1, p4), p7), b_hi = fmaxf(fmaxf(p1, p4), p7), b_mid = med3f(p1, p4, p7);
    const float c_lo = fminf(fminf(p2, p5), p8), c_hi = fmaxf(fmaxf(p2, p5), p8), c_mid = med3f(p2, p5, p8);
    return med3f(fmaxf(fmaxf(a_lo, b_lo), c_lo), med3f(a_mid, b_mid, c_mid), fminf(fminf(a_hi, b_hi), c_hi));
}

// s points at the window centre inside a tile with row stride S whose halo already holds the replicated
// border (so no clamping here).  CHECK_NAN: test the window and fall back to the exact network.
template <int S, bool CHECK_NAN>
__device__ __forceinline__ float median9_tile(const float* __restrict__ s, bool simd) {
    const float p0 = s[-S - 1], p1 = s[-S], p2 = s[-S + 1];
    const float p3 = s[-1], p4 = s[0], p5 = s[1];
    const float p6 = s[S - 1], p7 = s[S], p8 = s[S + 1];
    if (CHECK_NAN) {
        const float sum = ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7)) + p8;
        if (!(sum == sum)) {  // a NaN (or +inf and -inf) in the window: exact OpenCV semantics
            return simd ? vs_median9_net<true>(p0, p1, p2, p3, p4, p5, p6, p7, p8)
                        : vs_median9_net<false>(p0, p1, p2, p3, p4, p5, p6, p7, p8);
        }
    }
    return median9_fast(p0, p1, p2, p3, p4, p5, p6, p7, p8);
}

constexpr int TS = TILE + 5;  // row stride of the shared tiles (36 columns + 1 pad)
constexpr int TR = TILE + 4;  // rows of the shared tiles (halo 2)

// Blur phase shared by K2 and K4.  s_fill: TR x TS tile whose cell (r, c) is grid cell (ty0 - 2 + r, tx0 - 2 + c);
// rows/columns 1..34 must be final (hole-filled) and, outside the grid, replicated from the nearest inside cell.
// blockDim = (32, 8).  Writes out[(gy - out_row0) * W + gx] for gy in [ty0, min(ty0 + 32, row_limit)).
__device__ __forceinline__ unsigned blur_tile(const float* __restrict__ s_fill, int ty0, int tx0, int H, int W,
                                              int row_limit, bool tile_has_nan, bool simd_cols,
                                              float* __restrict__ out, int out_row0) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gx = tx0 + tx;
    unsigned n_nan = 0;
    if (H == 1 || W == 1) {  // OpenCV's 1-D special case (block-uniform)
#pragma unroll
        for (int k = 0; k < TILE / 8; ++k) {
            const int r = ty + 8 * k, gy = ty0 + r;
            if (gy < row_limit && gx < W) {
                const float* c = s_fill + (r + 2) * TS + (tx + 2);
                const float m = (H == 1) ? vs_median3_line(c[-1], c[0], c[1]) : vs_median3_line(c[-TS], c[0], c[TS]);
                out[(size_t)(gy - out_row0) * W + gx] = m;
                n_nan += (m != m);
            }
        }
        return n_nan;
    }
    const bool simd = simd_cols && gx >= 1 && gx <= W - 2;
#pragma unroll
    for (int k = 0; k < TILE / 8; ++k) {
        const int r = ty + 8 * k, gy = ty0 + r;
        if (gy < row_limit && gx < W) {
            const float* c = s_fill + (r + 2) * TS + (tx + 2);
            const float m = tile_has_nan ? median9_tile<TS, true>(c, simd) : median9_tile<TS, false>(c, simd);
            out[(size_t)(gy - out_row0) * W + gx] = m;
            n_nan += (m != m);
        }
    }
    return n_nan;
}

// Replicate the grid border into the part of the tile's 1-cell halo that lies outside the grid
// (cv2 BORDER_REPLICATE).  Only tiles touching the grid border have such cells.  rows_lo/rows_hi: first/last grid
// row that is valid to read (K4 row bands), normally 0 and H-1.
__device__ __forceinline__ void replicate_border(float* __restrict__ s_fill, int ty0, int tx0, int H, int W) {
    if (ty0 > 0 && tx0 > 0 && ty0 + TILE < H && tx0 + TILE < W) return;  // interior tile (block-uniform)
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < (TILE + 2) * (TILE + 2); i += kThreads) {
        const int r = 1 + i / (TILE + 2), c = 1 + i % (TILE + 2);
        const int gy = ty0 - 2 + r, gx = tx0 - 2 + c;
        if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
            const int cy = min(max(gy, 0), H - 1), cx = min(max(gx, 0), W - 1);
            const int rr = cy - (ty0 - 2), cc = cx - (tx0 - 2);
            if (rr >= 1 && rr <= TILE + 2 && cc >= 1 && cc <= TILE + 2) s_fill[r * TS + c] = s_fill[rr * TS + cc];
        }
    }
}

// blockDim = (32, 8); one CTA per 32x32 output tile.
template <typename Key>
__global__ void __launch_bounds__(kThreads)
k_grid_finalize(const Key* __restrict__ keygrid, int W, int H, typename KeyTraits<Key>::value_t* __restrict__ filled_out,
                float* __restrict__ blur_out, int simd_cols, unsigned long long* __restrict__ nan_count) {
    typedef typename KeyTraits<Key>::value_t T;
    constexpr bool kSameTile = sizeof(T) == sizeof(float);   // float32 keys: one tile serves fill and blur
    __shared__ T s_raw[TR * TS];                    // decoded keys; holes are patched in place after phase 2
    __shared__ float s_fill32[kSameTile ? 1 : TR * TS];   // float32 copy (what cv2.medianBlur sees) for the f64 path
    __shared__ unsigned short s_hole_pos[(TILE + 2) * (TILE + 2)];
    __shared__ T s_hole_val[(TILE + 2) * (TILE + 2)];
    __shared__ int s_nholes, s_has_nan;
    float* s_fill = kSameTile ? reinterpret_cast<float*>(s_raw) : s_fill32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 32 + tx;
    const int tx0 = blockIdx.x * TILE, ty0 = blockIdx.y * TILE;
    if (tid == 0) {
        s_nholes = 0;
        s_has_nan = 0;
    }
    __syncthreads();

    // 1. decode keys (+2 halo).  Key 0 (= empty, and what is used outside the grid) decodes to NaN, and the fill
    //    only uses in-range neighbours (:73).  Holes = empty cells inside the grid within the 1-cell halo.
    for (int i = tid; i < TR * TR; i += kThreads) {
        const int r = i / TR, c = i - r * TR;
        const int gy = ty0 - 2 + r, gx = tx0 - 2 + c;
        const bool inside = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
        Key k = 0;
        if (inside) k = keygrid[(size_t)gy * W + gx];
        const T v = KeyTraits<Key>::decode(k);
        const int pos = r * TS + c;
        s_raw[pos] = v;
        if (!kSameTile) s_fill[pos] = (float)v;     // produce_dsm.py:58 astype(np.float32)
        if (inside && k == 0 && (unsigned)(r - 1) < (unsigned)(TILE + 2) && (unsigned)(c - 1) < (unsigned)(TILE + 2))
            s_hole_pos[atomicAdd(&s_nholes, 1)] = (unsigned short)pos;
    }
    __syncthreads();

    // 2. hole fill: NaN cell <- median of its non-NaN 3x3 neighbours in the PRE-fill grid (dense over the list;
    //    results are staged so that the fill does not cascade, lib/proj_to_grid.py:65)
    const int nholes = s_nholes;
    for (int h = tid; h < nholes; h += kThreads) {
        const T* c = s_raw + s_hole_pos[h];
        T nb[8] = {c[-TS - 1], c[-TS], c[-TS + 1], c[-1], c[1], c[TS - 1], c[TS], c[TS + 1]};
        const T v = vs_median_of_valid8<T>(nb);
        s_hole_val[h] = v;
        if (v != v) s_has_nan = 1;
    }
    __syncthreads();
    for (int h = tid; h < nholes; h += kThreads) {
        const int pos = s_hole_pos[h];
        const T v = s_hole_val[h];
        s_raw[pos] = v;
        if (!kSameTile) s_fill[pos] = (float)v;
    }
    __syncthreads();
    if (filled_out != nullptr) {
#pragma unroll
        for (int k = 0; k < TILE / 8; ++k) {
            const int r = ty + 8 * k, gy = ty0 + r, gx = tx0 + tx;
            if (gy < H && gx < W) filled_out[(size_t)gy * W + gx] = s_raw[(r + 2) * TS + (tx + 2)];
        }
    }
    if (blur_out == nullptr) return;

    // 3. cv2.medianBlur(., 3) with replicated borders
    replicate_border(s_fill, ty0, tx0, H, W);
    __syncthreads();
    const unsigned n_nan = blur_tile(s_fill, ty0, tx0, H, W, H, s_has_nan != 0, simd_cols != 0, blur_out, 0);
    block_count_flush(n_nan, nan_count);
}

// in: rows [in_row0, ...) of an H_total-row image; writes rows [row_begin, row_end)
__global__ void __launch_bounds__(kThreads)
k_median3x3(const float* __restrict__ in, int in_row0, int H, int W, int row_begin, int row_end,
            float* __restrict__ out, int simd_cols, unsigned long long* __restrict__ nan_count) {
    __shared__ float s_fill[TR * TS];
    __shared__ int s_has_nan;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tx0 = blockIdx.x * TILE, ty0 = row_begin + blockIdx.y * TILE;
    if (ty == 0 && tx == 0) s_has_nan = 0;
    __syncthreads();
    bool saw_nan = false;
#pragma unroll
    for (int it = 0; it < (TR + 7) / 8; ++it) {
        const int r = ty + 8 * it;
        if (r >= 1 && r <= TILE + 2) {
            const int gy = ty0 - 2 + r;
            // rows the caller supplies: max(row_begin-1, 0) .. min(row_end, H-1)
            const bool row_ok = gy >= 0 && gy < H && gy >= row_begin - 1 && gy <= row_end;
            const float* __restrict__ row = in + (size_t)(row_ok ? gy - in_row0 : 0) * W;
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c = pass * 32 + tx;
                if (pass == 0 || tx < 4) {
                    const int gx = tx0 - 2 + c;
                    const bool inside = row_ok && gx >= 0 && gx < W;
                    const float v = inside ? row[gx] : CUDART_NAN_F;
                    s_fill[r * TS + c] = v;
                    saw_nan |= inside && (v != v) && c >= 1 && c <= TILE + 2;
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, saw_nan) && tx == 0) s_has_nan = 1;
    __syncthreads();
    replicate_border(s_fill, ty0, tx0, H, W);
    __syncthreads();
    const unsigned n_nan = blur_tile(s_fill, ty0, tx0, H, W, row_end, s_has_nan != 0, simd_cols != 0, out, row_begin);
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
    k_grid_finalize<uint32_t><<<grid, dim3(32, 8), 0, stream>>>(keygrid, xsize, ysize, nullptr, dsm_out,
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
    k_grid_finalize<unsigned long long><<<grid, dim3(32, 8), 0, stream>>>(
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
    k_median3x3<<<grid, dim3(32, 8), 0, stream>>>(in, in_row0, H_total, W, row_begin, row_end, out,
                                               simd_cols_for(W, simd_lanes),
                                               reinterpret_cast<unsigned long long*>(nan_count));
    VS_CHECK_LAUNCH(ctx, "k_median3x3");
    return VS_OK;
}

}  // extern "C"
