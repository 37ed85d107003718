// Stage C kernel K3: cross-view robust fusion, aggregate_2p5d.py:65-78, one thread per grid cell.
//
//   num = V - #NaN ; cells with num <= 2 -> NaN                                        (:69-71)
//   med = nanmedian(x)         = (s[(k-1)/2] + s[k/2]) / 2 in float32 over the k sorted valid values   (:74)
//   mad = nanmedian(|x - med|) same rule                                                (:75)
//   reject |x - med| > mad (strict; NaN compares false)                                 (:76-77)
//   mean = nanmean(x) = float32 pairwise sum (numpy's add.reduce order over the V-long view axis, rejected
//          and NaN entries contributing +0) / count                                     (:78)
//
// Layout: V planes of rows*W float32 (per-view DSMs, sorted view order).  Adjacent threads own adjacent cells,
// so every per-view load is a coalesced 128-byte line per warp; each plane element is read from HBM once
// (the re-reads of passes 2 and 3 hit L1/L2).  Algorithmic traffic: 4V bytes read + 4 bytes written per cell.
//
// V <= 64: values live in registers and are sorted with a Batcher merge-exchange network (sortnets_gen.cuh),
//          branch-free.  V > 64: per-thread quickselect on a shared-memory column (valid values compacted).
#include <math_constants.h>

#include "sortnets_gen.cuh"
#include "vs_common.cuh"

namespace {

constexpr int kBlockSmall = 128;

template <int N> struct SortNet;
#define VS_DEF_SORTNET(N)                                                     \
    template <> struct SortNet<N> {                                           \
        static __device__ __forceinline__ void sort(float (&a)[N]) {          \
            VS_SORTNET_##N(VS_CE_REG)                                         \
        }                                                                     \
    };
#define VS_CE_REG(i, j) { const float lo = fminf(a[i], a[j]); const float hi = fmaxf(a[i], a[j]); a[i] = lo; a[j] = hi; }
VS_DEF_SORTNET(8)
VS_DEF_SORTNET(16)
VS_DEF_SORTNET(24)
VS_DEF_SORTNET(32)
VS_DEF_SORTNET(40)
VS_DEF_SORTNET(48)
VS_DEF_SORTNET(56)
VS_DEF_SORTNET(64)
#undef VS_DEF_SORTNET

// (s[(k-1)/2] + s[k/2]) / 2 with static register indexing
template <int N>
__device__ __forceinline__ float middle_of_sorted(const float (&s)[N], int k) {
    // k <= N, so both middle positions are <= N/2: only the lower half of the sorted array is ever read and
    // the compiler drops every compare-exchange that feeds the upper half only.
    const int ilo = (k - 1) >> 1, ihi = k >> 1;
    float lo = s[0], hi = s[0];
#pragma unroll
    for (int i = 1; i <= N / 2; ++i) {
        lo = (i == ilo) ? s[i] : lo;
        hi = (i == ihi) ? s[i] : hi;
    }
    return __fdiv_rn(__fadd_rn(lo, hi), 2.0f);
}

// numpy pairwise sum over y[0..n) for n <= 128 (one leaf), y given by a functor with STATIC indices 0..N-1
template <int N, typename F>
__device__ __forceinline__ float pairwise_leaf_static(const F& y, int n) {
    if (n < 8) {
        float res = 0.0f;
#pragma unroll
        for (int i = 0; i < (N < 8 ? N : 7); ++i)
            if (i < n) res = __fadd_rn(res, y(i));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = y(j);
    const int nfull = n - (n & 7);
#pragma unroll
    for (int i = 8; i < N; i += 8) {
        if (i < nfull) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], y(i + j));
        }
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
#pragma unroll
    for (int i = 8; i < N; ++i)
        if (i >= nfull && i < n) res = __fadd_rn(res, y(i));
    return res;
}

template <int NV>
__global__ void __launch_bounds__(kBlockSmall)
k_fuse_small(const float* __restrict__ views, int64_t plane_stride, int V, int64_t n_cells, float* __restrict__ out) {
    const int64_t cell = blockIdx.x * (int64_t)kBlockSmall + threadIdx.x;
    if (cell >= n_cells) return;
    float x[NV];
    float s[NV];
    int k = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        float t = CUDART_NAN_F;
        if (v < V) t = __ldg(views + (int64_t)v * plane_stride + cell);
        x[v] = t;
        const bool ok = (t == t);
        k += ok;
        s[v] = ok ? t : CUDART_INF_F;
    }
    if (k <= 2) {  // :69-71
        out[cell] = CUDART_NAN_F;
        return;
    }
    SortNet<NV>::sort(s);
    const float med = middle_of_sorted<NV>(s, k);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const float d = fabsf(__fsub_rn(x[v], med));  // NaN stays NaN
        s[v] = (d == d) ? d : CUDART_INF_F;
    }
    SortNet<NV>::sort(s);
    const float mad = middle_of_sorted<NV>(s, k);
    int cnt = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const float d = fabsf(__fsub_rn(x[v], med));
        const bool keep = (x[v] == x[v]) && !(d > mad);  // :76-77
        cnt += keep;
        x[v] = keep ? x[v] : 0.0f;  // nanmean: NaN -> 0, then sum over the whole axis
    }
    auto y = [&](int i) -> float { return x[i]; };
    const float tot = pairwise_leaf_static<NV>(y, V);
    out[cell] = __fdiv_rn(tot, (float)cnt);
}

// ---------------------------------------------------------------------------------------------------------
// generic path (V > 64): shared-memory column per thread, stride = blockDim.x
// ---------------------------------------------------------------------------------------------------------
struct SmemCol {
    float* base;
    int stride;
    __device__ __forceinline__ float& operator[](int i) const { return base[i * stride]; }
};

// after the call: c[n] holds the n-th smallest of c[0..k), everything before it is <= c[n]
__device__ __forceinline__ void quickselect(const SmemCol& c, int k, int n) {
    int lo = 0, hi = k - 1;
    while (hi > lo) {
        if (hi - lo < 8) {  // insertion sort of the small remaining window
            for (int i = lo + 1; i <= hi; ++i) {
                const float t = c[i];
                int j = i - 1;
                while (j >= lo && c[j] > t) {
                    c[j + 1] = c[j];
                    --j;
                }
                c[j + 1] = t;
            }
            return;
        }
        // median-of-3 pivot
        const int mid = lo + ((hi - lo) >> 1);
        float a = c[lo], b = c[mid], d = c[hi];
        if (a > b) { const float t = a; a = b; b = t; }
        if (b > d) { const float t = b; b = d; d = t; }
        if (a > b) { const float t = a; a = b; b = t; }
        c[lo] = a; c[mid] = b; c[hi] = d;
        const float pivot = b;
        int i = lo, j = hi;
        // Hoare partition (stops on equal keys, so runs of equal heights split evenly)
        while (true) {
            do { ++i; } while (c[i] < pivot);
            do { --j; } while (c[j] > pivot);
            if (i >= j) break;
            const float t = c[i]; c[i] = c[j]; c[j] = t;
        }
        // now c[lo..j] <= pivot <= c[j+1..hi]
        if (n <= j) hi = j; else lo = j + 1;
    }
}

__device__ __forceinline__ float middle_by_select(const SmemCol& c, int k) {
    const int ilo = (k - 1) >> 1, ihi = k >> 1;
    quickselect(c, k, ihi);
    const float hi = c[ihi];
    float lo = hi;
    if (ilo != ihi) {
        lo = c[0];
        for (int i = 1; i < ihi; ++i) lo = fmaxf(lo, c[i]);
    }
    return __fdiv_rn(__fadd_rn(lo, hi), 2.0f);
}

struct KeepFn {
    const float* p;
    int64_t stride;
    float med, mad;
    __device__ __forceinline__ float operator()(int v) const {
        const float t = __ldg(p + (int64_t)v * stride);
        const float d = fabsf(__fsub_rn(t, med));
        return ((t == t) && !(d > mad)) ? t : 0.0f;
    }
};

__device__ __forceinline__ float pairwise_leaf_dyn(const KeepFn& y, int off, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, y(off + i));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = y(off + j);
    int i = 8;
    for (; i < n - (n & 7); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], y(off + i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, y(off + i));
    return res;
}

// numpy pairwise_sum: n > 128 -> split at n/2 rounded down to a multiple of 8
template <int DEPTH>
__device__ __forceinline__ float pairwise_dyn(const KeepFn& y, int off, int n) {
    if (n <= 128) return pairwise_leaf_dyn(y, off, n);
    int n2 = n >> 1;
    n2 -= n2 & 7;
    return __fadd_rn(pairwise_dyn<DEPTH - 1>(y, off, n2), pairwise_dyn<DEPTH - 1>(y, off + n2, n - n2));
}
template <>
__device__ __forceinline__ float pairwise_dyn<0>(const KeepFn& y, int off, int n) {
    return pairwise_leaf_dyn(y, off, n);  // unreachable for V <= 128 * 2^DEPTH (checked on the host)
}
constexpr int kPairwiseDepth = 5;  // V <= 4096

__global__ void k_fuse_generic(const float* __restrict__ views, int64_t plane_stride, int V, int64_t n_cells,
                               float* __restrict__ out) {
    extern __shared__ float s_col[];
    const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    const SmemCol c{s_col + threadIdx.x, (int)blockDim.x};
    const float* p = views + cell;
    int k = 0;
    for (int v = 0; v < V; ++v) {
        const float t = __ldg(p + (int64_t)v * plane_stride);
        if (t == t) c[k++] = t;
    }
    if (k <= 2) {
        out[cell] = CUDART_NAN_F;
        return;
    }
    const float med = middle_by_select(c, k);
    int j = 0;
    for (int v = 0; v < V; ++v) {
        const float t = __ldg(p + (int64_t)v * plane_stride);
        if (t == t) c[j++] = fabsf(__fsub_rn(t, med));
    }
    const float mad = middle_by_select(c, k);
    int cnt = 0;
    for (int v = 0; v < V; ++v) {
        const float t = __ldg(p + (int64_t)v * plane_stride);
        const float d = fabsf(__fsub_rn(t, med));
        cnt += ((t == t) && !(d > mad));
    }
    const KeepFn y{p, plane_stride, med, mad};
    const float tot = pairwise_dyn<kPairwiseDepth>(y, 0, V);
    out[cell] = __fdiv_rn(tot, (float)cnt);
}

template <int NV>
int launch_small(vs_ctx* ctx, const float* views, int64_t plane_stride, int V, int64_t n_cells, float* out,
                 cudaStream_t stream) {
    const int64_t blocks = (n_cells + kBlockSmall - 1) / kBlockSmall;
    k_fuse_small<NV><<<(unsigned)blocks, kBlockSmall, 0, stream>>>(views, plane_stride, V, n_cells, out);
    VS_CHECK_LAUNCH(ctx, "k_fuse_small");
    return VS_OK;
}

}  // namespace

extern "C" int vs_fuse_views(vs_ctx* ctx, const float* views, int64_t plane_stride, int32_t n_views, int32_t rows,
                             int32_t W, float* out_mean, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_fuse_views: NULL context");
    VS_REQUIRE(n_views >= 1 && n_views <= 128 * (1 << kPairwiseDepth), "vs_fuse_views: n_views must be 1..4096");
    VS_REQUIRE(rows >= 0 && W >= 0, "vs_fuse_views: negative size");
    const int64_t n_cells = (int64_t)rows * W;
    if (n_cells == 0) return VS_OK;
    VS_REQUIRE(views != nullptr && out_mean != nullptr, "vs_fuse_views: NULL array");
    VS_REQUIRE(plane_stride >= n_cells, "vs_fuse_views: plane_stride smaller than a plane");
    VS_REQUIRE(n_cells < ((int64_t)1 << 31) * kBlockSmall, "vs_fuse_views: too many cells");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    const int V = n_views;
    if (V <= 8) return launch_small<8>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 16) return launch_small<16>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 24) return launch_small<24>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 32) return launch_small<32>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 40) return launch_small<40>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 48) return launch_small<48>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 56) return launch_small<56>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 64) return launch_small<64>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    // generic: one shared-memory column of V floats per thread
    int block = 128;
    while (block > 32 && (size_t)block * V * sizeof(float) > 200 * 1024) block >>= 1;
    const size_t smem = (size_t)block * V * sizeof(float);
    VS_REQUIRE(smem <= 227 * 1024, "vs_fuse_views: too many views for one shared-memory column per cell");
    VS_CUDA(cudaFuncSetAttribute(k_fuse_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (n_cells + block - 1) / block;
    k_fuse_generic<<<(unsigned)blocks, block, smem, stream>>>(views, plane_stride, V, n_cells, out_mean);
    VS_CHECK_LAUNCH(ctx, "k_fuse_generic");
    return VS_OK;
}
