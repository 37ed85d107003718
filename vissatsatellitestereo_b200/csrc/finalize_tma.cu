// Stage B, Blackwell pipeline: persistent CTAs fed by TMA.
//
// k_tile_pipeline_tma<KEYS>   the same per-tile work as finalize.cu (K2 when KEYS, K4 otherwise), but each CTA is
// persistent and its 36x72 input box (tile + 2-cell halo, start column a multiple of 4 as TMA requires) arrives by
// one `cp.async.bulk.tensor.2d` (TMA) per tile into a double-buffered shared tile, signalled through an mbarrier:
// the load of tile k+1 overlaps the hole fill + 3x3 median of tile k, which the plain-load kernels (one tile per
// CTA: load -> barrier -> compute) cannot do.  TMA's out-of-bounds zero fill is exactly the "empty" key (decodes
// to NaN) for K2; for K4 the cells outside the grid are overwritten by the border replication anyway.
//
// Requirements: row pitch (W * 4 bytes) and base address multiples of 16 bytes; otherwise the callers in finalize.cu
// fall back to the plain-load kernels.
#include <cuda.h>

#include "finalize_common.cuh"

using namespace vsfin;

namespace {

constexpr int BOX_W = TS;                 // 72 columns: tile column c sits at box column OFF + c; box column 0 is grid
                                          // column tx0 - 4 (TMA needs the innermost coordinate to be a multiple of 16 bytes)
constexpr int BOX_H = TR;                 // 36 rows
constexpr uint32_t kBoxBytes = BOX_W * BOX_H * sizeof(uint32_t);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// KEYS: input = uint32 key grid (decode + hole fill + blur).  !KEYS: input = float32 image rows (blur only).
template <bool KEYS>
__global__ void __launch_bounds__(kThreads)
k_tile_pipeline_tma(const __grid_constant__ CUtensorMap map, int W, int H, int row_begin, int row_end, int in_row0,
                    int tiles_x, int n_tiles, float* __restrict__ out, int simd_cols,
                    unsigned long long* __restrict__ nan_count) {
    __shared__ __align__(128) uint32_t s_buf[2][BOX_H * BOX_W];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ unsigned short s_hole_pos[KEYS ? MAX_HOLES : 1];
    __shared__ float s_hole_val[KEYS ? MAX_HOLES : 1];
    __shared__ int s_has_nan;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // one elected thread arms the barrier and issues the TMA; the descriptor is addressed as the kernel's
    // __grid_constant__ parameter (it must not be copied to local memory, so no lambda/by-value helper here)
#define VS_ISSUE_TILE(TILE, BUF)                                                  \
    do {                                                                          \
        const int bx_ = (TILE) % tiles_x, by_ = (TILE) / tiles_x;                 \
        const int x_ = bx_ * TW - 2 - OFF;                                        \
        const int y_ = row_begin + by_ * TH - 2 - in_row0;                        \
        mbar_arrive_expect_tx(&s_bar[(BUF)], kBoxBytes);                          \
        tma_load_2d(s_buf[(BUF)], &map, &s_bar[(BUF)], x_, y_);                   \
    } while (0)

    int tile = blockIdx.x;
    if (tile < n_tiles && tid == 0) VS_ISSUE_TILE(tile, 0);
    unsigned n_nan = 0;
    for (int k = 0; tile < n_tiles; ++k, tile += gridDim.x) {
        const int buf = k & 1;
        const int next = tile + gridDim.x;
        if (tid == 0) {
            s_has_nan = 0;
            if (next < n_tiles) {
                fence_proxy_async();   // buffer buf^1 was last touched through the generic proxy (iteration k-1)
                VS_ISSUE_TILE(next, buf ^ 1);
            }
        }
        mbar_wait(&s_bar[buf], (k >> 1) & 1);
        __syncthreads();               // also publishes the s_has_nan reset

        const int bx = tile % tiles_x, by = tile / tiles_x;
        const int tx0 = bx * TW, ty0 = row_begin + by * TH;
        uint32_t* raw = s_buf[buf];
        float* s_fill = reinterpret_cast<float*>(raw);

        // ---- pass over the box: decode keys in place + list holes (KEYS) / look for NaN (!KEYS).
        // Every warp keeps its own hole list (its rows only): the count lives in a register, no atomics.
        bool saw_nan = false;
        int wcnt = 0;                                  // warp-uniform
        unsigned short* my_pos = s_hole_pos + (KEYS ? warp * HSEG : 0);
        float* my_val = s_hole_val + (KEYS ? warp * HSEG : 0);
        for (int r = warp; r < BOX_H; r += kThreads / 32) {
            const int gy = ty0 - 2 + r;
            const bool row_in = (unsigned)gy < (unsigned)H && (unsigned)(r - 1) < (unsigned)(TH + 2);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int bc = lane + 32 * part;       // box column
                bool hole = false;
                const int pos = r * BOX_W + bc;
                if (bc < BOX_W) {
                    const int c = bc - OFF;            // tile column (grid column tx0 - 2 + c)
                    const int gx = tx0 - 2 + c;
                    const bool inner = row_in && (unsigned)gx < (unsigned)W && (unsigned)(c - 1) < (unsigned)(TW + 2);
                    if (KEYS) {
                        const uint32_t key = raw[pos];
                        s_fill[pos] = vs_unkey32(key);                 // key 0 (empty / outside the grid) -> NaN
                        hole = inner && key == 0;
                    } else {
                        const float v = s_fill[pos];
                        saw_nan |= inner && (v != v);
                    }
                }
                if (KEYS) {
                    const unsigned m = __ballot_sync(0xffffffffu, hole);
                    if (m) {   // warp-uniform
                        if (hole) my_pos[wcnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)pos;
                        wcnt += __popc(m);
                    }
                }
            }
        }
        if (!KEYS && saw_nan) s_has_nan = 1;
        __syncthreads();

        // ---- hole fill (lib/proj_to_grid.py:65-79), staged so that it reads the PRE-fill grid
        if (KEYS) {
            bool nan_left = false;
            for (int i = lane; i < wcnt; i += 32) {
                const float* c = s_fill + my_pos[i];
                float nb[8] = {c[-BOX_W - 1], c[-BOX_W], c[-BOX_W + 1], c[-1], c[1], c[BOX_W - 1], c[BOX_W], c[BOX_W + 1]};
                const float v = vs_median_of_valid8<float>(nb);
                my_val[i] = v;
                nan_left |= (v != v);
            }
            if (nan_left) s_has_nan = 1;
            __syncthreads();
            for (int i = lane; i < wcnt; i += 32) s_fill[my_pos[i]] = my_val[i];
            __syncthreads();
        }

        // ---- cv2.medianBlur(., 3) with replicated borders
        replicate_border(s_fill, ty0, tx0, H, W);
        __syncthreads();
        n_nan += blur_tile(s_fill, ty0, tx0, H, W, row_end, s_has_nan != 0, simd_cols != 0, out, row_begin);
        __syncthreads();   // everyone is done with this buffer before it is refilled (prefetch of iteration k+1)
    }
#undef VS_ISSUE_TILE
    block_count_flush(n_nan, nan_count);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool make_map(CUtensorMap* map, const void* base, bool is_float, int W, int rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BOX_W, (cuuint32_t)BOX_H};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, is_float ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2,
              const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Called by finalize.cu.  Return: 1 = launched, 0 = not applicable (caller uses the plain-load kernel), <0 = error.
int vs_finalize_tma_try(vs_ctx* ctx, bool keys, const void* in, int in_rows, int in_row0, int W, int H, int row_begin,
                        int row_end, float* out, int simd_cols, unsigned long long* nan_count, cudaStream_t stream) {
    if (H == 1 || W == 1) return 0;
    if ((W & 3) != 0 || (reinterpret_cast<uintptr_t>(in) & 15) != 0) return 0;
    CUtensorMap map;
    if (!make_map(&map, in, !keys, W, in_rows)) return 0;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (row_end - row_begin + TH - 1) / TH;
    const int n_tiles = tiles_x * tiles_y;
    if (n_tiles == 0) return 1;
    int grid = ctx->sm_count * 4;
    if (grid > n_tiles) grid = n_tiles;
    if (keys)
        k_tile_pipeline_tma<true><<<grid, kThreads, 0, stream>>>(map, W, H, row_begin, row_end, in_row0, tiles_x, n_tiles,
                                                                 out, simd_cols, nan_count);
    else
        k_tile_pipeline_tma<false><<<grid, kThreads, 0, stream>>>(map, W, H, row_begin, row_end, in_row0, tiles_x,
                                                                  n_tiles, out, simd_cols, nan_count);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        vs_cuda_fail(e, "k_tile_pipeline_tma");
        return -1;
    }
    ctx->launches++;
    return 1;
}
