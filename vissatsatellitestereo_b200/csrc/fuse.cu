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
// V <= 32:  k_fuse_small  -- original and sorted values both in registers, Batcher merge-exchange network
//            (sortnets_gen.cuh), branch-free.
// V <= 128: k_fuse_medium -- only the sorted copy in registers, originals re-read from L2 for the sum.
// V > 128:  k_fuse_large  -- 4, 8 or 32 threads per cell, sorted runs in shared memory, bitwise bisection.
#include <math_constants.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "sortnets_gen.cuh"
#include "vs_common.cuh"

namespace {

constexpr int kBlockSmall = 128;

template <int N> struct SortNet;
#define VS_DEF_SORTNET(N)                                                     \
    template <> struct SortNet<N> {                                           \
        template <int M>                                                      \
        static __device__ __forceinline__ void sort(float (&a)[M]) {          \
            VS_SORTNET_##N(VS_CE_REG)                                         \
        }                                                                     \
    };
#define VS_CE_REG(i, j) { const float lo = fminf(a[i], a[j]); const float hi = fmaxf(a[i], a[j]); a[i] = lo; a[j] = hi; }
VS_DEF_SORTNET(8)
VS_DEF_SORTNET(16)
VS_DEF_SORTNET(24)
VS_DEF_SORTNET(32)
VS_DEF_SORTNET(36)
VS_DEF_SORTNET(40)
VS_DEF_SORTNET(44)
VS_DEF_SORTNET(48)
VS_DEF_SORTNET(52)
VS_DEF_SORTNET(56)
VS_DEF_SORTNET(60)
VS_DEF_SORTNET(64)
VS_DEF_SORTNET(80)
VS_DEF_SORTNET(88)
VS_DEF_SORTNET(96)
VS_DEF_SORTNET(104)
VS_DEF_SORTNET(112)
VS_DEF_SORTNET(120)
VS_DEF_SORTNET(128)
#undef VS_DEF_SORTNET

// Bitonic merge (ascending) of a bitonic sequence held in N registers; N is padded to a power of two with
// virtual +inf wires, whose compare-exchanges are no-ops and are skipped at compile time.
template <int N>
__device__ __forceinline__ void bitonic_merge_regs(float (&a)[N]) {
    constexpr int N2 = N <= 8 ? 8 : N <= 16 ? 16 : N <= 32 ? 32 : N <= 64 ? 64 : 128;
#pragma unroll
    for (int j = N2 / 2; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int l = i ^ j;
            if (l > i && l < N) {
                const float lo = fminf(a[i], a[l]);
                const float hi = fmaxf(a[i], a[l]);
                a[i] = lo;
                a[l] = hi;
            }
        }
    }
}

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
    for (int i = 8; i + 8 <= N; i += 8) {   // N need not be a multiple of 8: a full block of 8 never starts past N - 8
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
    // |s - med| over the SORTED values falls, then rises (the +inf pads stay +inf at the end): a bitonic
    // sequence, so one bitonic merge sorts it -- no second full sort.
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = fabsf(__fsub_rn(s[v], med));
    bitonic_merge_regs<NV>(s);
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
// medium V (33..128): still one thread per cell and a register sorting network, but only the SORTED copy lives
// in registers; the original-order values needed by the pairwise sum are re-read from L2 (KeepFn).
// ---------------------------------------------------------------------------------------------------------
struct KeepFn {
    const float* p;
    int64_t stride;
    float med, mad;
    __device__ __forceinline__ bool kept(int v) const {
        const float t = __ldg(p + (int64_t)v * stride);
        const float d = fabsf(__fsub_rn(t, med));
        return (t == t) && !(d > mad);          // aggregate_2p5d.py:76-77
    }
    __device__ __forceinline__ float operator()(int v) const {
        const float t = __ldg(p + (int64_t)v * stride);
        const float d = fabsf(__fsub_rn(t, med));
        return ((t == t) && !(d > mad)) ? t : 0.0f;   // nanmean: NaN / rejected -> +0
    }
};

template <int NV>
__global__ void __launch_bounds__(kBlockSmall, (NV > 64 && NV <= 112) ? 4 : 1)
k_fuse_medium(const float* __restrict__ views, int64_t plane_stride, int V, int64_t n_cells, float* __restrict__ out) {
    const int64_t cell = blockIdx.x * (int64_t)kBlockSmall + threadIdx.x;
    if (cell >= n_cells) return;
    float s[NV];
    int k = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        float t = CUDART_NAN_F;
        if (v < V) t = __ldg(views + (int64_t)v * plane_stride + cell);
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
    for (int v = 0; v < NV; ++v) s[v] = fabsf(__fsub_rn(s[v], med));
    bitonic_merge_regs<NV>(s);
    const float mad = middle_of_sorted<NV>(s, k);
    // numpy pairwise sum for 8 <= V <= 128: one leaf, 8 strided accumulators, then the tail
    const KeepFn y{views + cell, plane_stride, med, mad};
    int cnt = 0;
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        r[j] = y(j);
        cnt += y.kept(j);
    }
    const int nfull = V - (V & 7);
    for (int i = 8; i < nfull; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r[j] = __fadd_rn(r[j], y(i + j));
            cnt += y.kept(i + j);
        }
    }
    float tot = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (int i = nfull; i < V; ++i) {
        tot = __fadd_rn(tot, y(i));
        cnt += y.kept(i);
    }
    out[cell] = __fdiv_rn(tot, (float)cnt);
}

// ---------------------------------------------------------------------------------------------------------
// large V (129..2048): LANES threads per cell.
//
// 1. The CTA loads its cells x V tile with coalesced 128-byte rows into shared memory.
// 2. Lane l of a cell takes values v = l, l+LANES, ... (<= NVL of them), sorts them in registers with the same
//    merge-exchange network as the small path and writes them back as a sorted run of order-preserving keys.
// 3. Order statistics of the union of the LANES runs come from a 32-step bitwise bisection on the key; each
//    step is one branch-free lower_bound per lane (<= 7 shared-memory probes) and a shuffle reduction.
// 4. For the MAD each lane turns its sorted run into |x - med| (bitonic), merges it in registers, and the same
//    bisection runs on those runs.
// 5. The float32 sum in numpy's pairwise order is accumulated by lanes 0..7 of the cell (numpy's 8 strided
//    accumulators), re-reading the views from L2.
// ---------------------------------------------------------------------------------------------------------
constexpr int kLargeThreads = 256;

template <int LANES>
__device__ __forceinline__ int group_sum(int v, unsigned mask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <int LANES>
__device__ __forceinline__ uint32_t group_max(uint32_t v, unsigned mask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(mask, v, o));
    return v;
}

// Runs are stored padded with 0xffffffff up to NVLP = the power of two above NVL, so the branch-free lower bound
// needs no range checks.
template <int NVL> struct RunPad { static constexpr int value = NVL < 8 ? 8 : NVL < 16 ? 16 : NVL < 32 ? 32 : NVL < 64 ? 64 : 128; };

// number of keys < T in this lane's sorted run (stride LANES)
template <int LANES, int NVL>
__device__ __forceinline__ int run_lower_bound(const uint32_t* __restrict__ run, uint32_t T) {
    constexpr int P = RunPad<NVL>::value;
    int lb = 0;
#pragma unroll
    for (int step = P / 2; step >= 1; step >>= 1)
        if (run[(lb + step - 1) * LANES] < T) lb += step;
    return lb;
}

// The same lower bound with the first three levels of the search done on 7 splitter keys held in registers
// (sp[j] = run element (j+1)*P/8 - 1): three dependent shared-memory probes fewer per call.  The bisection below calls
// it 31..32 times per order statistic and is bound by exactly that dependent-load latency.
template <int LANES, int NVL>
__device__ __forceinline__ void run_load_splitters(const uint32_t* __restrict__ run, uint32_t (&sp)[7]) {
    constexpr int B = RunPad<NVL>::value / 8;
#pragma unroll
    for (int j = 0; j < 7; ++j) sp[j] = run[((j + 1) * B - 1) * LANES];
}
template <int LANES, int NVL>
__device__ __forceinline__ int run_lower_bound_sp(const uint32_t* __restrict__ run, const uint32_t (&sp)[7], uint32_t T) {
    constexpr int B = RunPad<NVL>::value / 8;
    const bool c1 = sp[3] < T;
    const uint32_t s2 = c1 ? sp[5] : sp[1];
    const bool c2 = s2 < T;
    const uint32_t s3 = c1 ? (c2 ? sp[6] : sp[4]) : (c2 ? sp[2] : sp[0]);
    const bool c3 = s3 < T;
    int lb = (c1 ? 4 * B : 0) + (c2 ? 2 * B : 0) + (c3 ? B : 0);
#pragma unroll
    for (int step = B / 2; step >= 1; step >>= 1)
        if (run[(lb + step - 1) * LANES] < T) lb += step;
    return lb;
}

// keys of rank r_hi and r_hi-1 (0-based) in the union of the group's runs; k_lo_out = k_hi if want_lo is false
template <int LANES, int NVL>
__device__ __forceinline__ void group_select(const uint32_t* __restrict__ run, int r_hi, bool want_lo, unsigned mask,
                                             uint32_t start_key, int start_bit, uint32_t& k_hi_out, uint32_t& k_lo_out) {
    uint32_t sp[7];
    run_load_splitters<LANES, NVL>(run, sp);
    uint32_t K = start_key;
    for (int b = start_bit; b >= 0; --b) {  // largest K with count(keys < K) <= r_hi  ==  the r_hi-th smallest key
        const uint32_t T = K | (1u << b);
        const int c = group_sum<LANES>(run_lower_bound_sp<LANES, NVL>(run, sp, T), mask);
        if (c <= r_hi) K = T;
    }
    k_hi_out = K;
    k_lo_out = K;
    if (want_lo) {
        const int lb = run_lower_bound_sp<LANES, NVL>(run, sp, K);
        const int below = group_sum<LANES>(lb, mask);
        if (below == r_hi) {  // rank r_hi-1 is the largest key below K
            const uint32_t mine = lb > 0 ? run[(lb - 1) * LANES] : 0u;
            k_lo_out = group_max<LANES>(mine, mask);
        }
    }
}

template <int LANES>
__device__ __forceinline__ uint32_t group_min(uint32_t v, unsigned mask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(mask, v, o));
    return v;
}

// The same two order statistics by a SPLIT search instead of a bit-wise bisection (round 2; opt-in, see g_fuse_split).  Lane l keeps p_l = how many
// of its sorted keys belong to the r_hi smallest of the union (sum of p_l = r_hi).  The split is right when the largest
// key left of the splits, A, is <= the smallest key right of them, B; then B has rank r_hi and A rank r_hi - 1.  The runs
// of a cell are interleaved views (lane = view mod LANES), i.e. statistically alike, so the balanced start p_l ~ r_hi /
// LANES is a few exchanges away from the answer (2 probes per iteration against 32 bisection steps of 3 probes);
// every iteration moves keys from lanes that hold too many of the small ones to lanes that hold too few.  An input that
// needs more than kSplitIters exchanges (views that alternate systematically between two levels) finishes with the
// bisection above, restricted to the bracket [B, A] the exchanges have reached, so the worst case stays bounded.  Keys: 0 is below every key, the run pads 0xffffffff above.
constexpr int kSplitIters = 16;
__device__ int g_fuse_split = 0;      // VISSAT_FUSE_SPLIT=1: try the split search first (measured: 1.34 vs 1.40 ms at V = 400 when
                                      // the lanes' runs are alike, 1.69 ms when they are not -- so it is opt-in)
template <int LANES, int NVL>
__device__ __forceinline__ void group_select_split(const uint32_t* __restrict__ run, int r_hi, bool want_lo, unsigned mask,
                                                   uint32_t& k_hi_out, uint32_t& k_lo_out) {
    if (!g_fuse_split) {              // uniform over the grid
        group_select<LANES, NVL>(run, r_hi, want_lo, mask, 0u, 31, k_hi_out, k_lo_out);
        return;
    }
    const int me = threadIdx.x & 31;
    const int lane = me & (LANES - 1);
    int p = r_hi / LANES + (lane < (r_hi & (LANES - 1)) ? 1 : 0);
    uint32_t A = 0xffffffffu, B = 0u;
    for (int it = 0; it < kSplitIters; ++it) {
        const uint32_t a = p > 0 ? run[(p - 1) * LANES] : 0u;
        const uint32_t b = run[p * LANES];                   // p <= NVL < RunPad: a pad at worst
        A = group_max<LANES>(a, mask);
        B = group_min<LANES>(b, mask);
        if (A <= B) {                                        // group-uniform
            k_hi_out = B;
            k_lo_out = want_lo ? A : B;
            return;
        }
        // A > B.  Every lane whose top-left key exceeds B would gain from handing it over, every lane whose first right
        // key is below A from taking one: m = min(#over, #under) lanes of each kind move their split by one, so that
        // the sum of the p_l stays r_hi (if exactly one pair is out of order this is the single exchange A <-> B).
        // Any sequence of such moves is safe -- only the exit test above decides -- and all lanes adjust in parallel:
        // the iteration count is the largest per-lane offset from the balanced start, not the sum.
        const unsigned over = __ballot_sync(mask, a > B), under = __ballot_sync(mask, b < A);
        const int m = min(__popc(over), __popc(under));
        const unsigned below_me = (1u << me) - 1u;
        const bool dec = a > B && __popc(over & below_me) < m;
        const bool inc = b < A && __popc(under & below_me) < m;
        p += (inc ? 1 : 0) - (dec ? 1 : 0);
    }
    // Not there yet.  With r_hi keys left of the splits, A among them and B < A right of them: fewer than r_hi keys are
    // below B and more than r_hi are <= A, so B <= answer <= A, and the bisection only has to resolve the bits below the
    // common prefix of the two.
    const int hb = 31 - __clz(A ^ B);                        // highest differing bit (A != B here)
    const uint32_t prefix = A & ~((2u << hb) - 1u);
    group_select<LANES, NVL>(run, r_hi, want_lo, mask, prefix, hb, k_hi_out, k_lo_out);
}

// nanmean of the survivors of one cell in numpy's pairwise order, computed by the LANES threads of the cell's group
// (lane j < 8 owns numpy's accumulator r[j]); the views are re-read from L2.  Every lane returns the mean.
// `present(v)`: whether view v takes part at all (dense stacks: always; sparse fusion: the tile's occupancy bit -- an
// absent view counts as NaN, contributes +0 to the sum like any rejected value, and its memory is never read).
struct AllPresent {
    __device__ __forceinline__ bool operator()(int) const { return true; }
};
struct BitsPresent {
    const uint32_t* bits;   // shared memory: the tile's occupancy words
    __device__ __forceinline__ bool operator()(int v) const { return (bits[v >> 5] >> (v & 31)) & 1u; }
};

template <int LANES, typename Present>
__device__ __forceinline__ float sum_survivors(const float* __restrict__ views, int64_t cell, int64_t plane_stride, int V,
                                               float med, float mad, int lane, unsigned gmask, const Present& present) {
    const KeepFn y{views + cell, plane_stride, med, mad};
    // numpy's 8 strided accumulators r[0..7] are spread over the first AL = min(LANES, 8) lanes of the group:
    // lane q < AL owns r[q], r[q + AL], ...
    constexpr int AL = LANES < 8 ? LANES : 8;
    constexpr int ACC = 8 / AL;
    const bool acc_lane = lane < AL;
    int cnt = 0;
    float total = 0.0f;
    {
        // iterative walk over numpy's recursion tree (leaves of <= 128 elements, splits at multiples of 8),
        // left to right; partial sums are combined with an explicit stack exactly as the recursion would
        float stack_val[12];
        int stack_lvl[12];
        int sp = 0;
        int seg_off[12], seg_n[12], seg_lvl[12];
        int tp = 0;
        seg_off[0] = 0; seg_n[0] = V; seg_lvl[0] = 0; tp = 1;
        while (tp > 0) {
            --tp;
            const int off = seg_off[tp], n = seg_n[tp], lvl = seg_lvl[tp];
            if (n > 128) {
                int n2 = n >> 1;
                n2 -= n2 & 7;
                // push right first so that left is processed first
                seg_off[tp] = off + n2; seg_n[tp] = n - n2; seg_lvl[tp] = lvl + 1; ++tp;
                seg_off[tp] = off; seg_n[tp] = n2; seg_lvl[tp] = lvl + 1; ++tp;
                continue;
            }
            // ---- leaf
            float leaf;
            if (n < 8) {
                leaf = 0.0f;
                for (int i = 0; i < n; ++i) {   // every lane computes the same value
                    if (!present(off + i)) continue;
                    leaf = __fadd_rn(leaf, y(off + i));
                    if (lane == 0) cnt += y.kept(off + i);
                }
            } else {
                float r[ACC];
#pragma unroll
                for (int m = 0; m < ACC; ++m) r[m] = 0.0f;
                const int nfull = n - (n & 7);
                if (acc_lane) {   // every view index of the leaf's full blocks is visited exactly once: count here too
#pragma unroll
                    for (int m = 0; m < ACC; ++m) {
                        const int v0 = off + lane + m * AL;
                        const float* q = y.p + (int64_t)v0 * y.stride;
                        const int64_t step8 = 8 * y.stride;
                        float t = present(v0) ? __ldg(q) : CUDART_NAN_F;
                        bool keep = (t == t) && !(fabsf(__fsub_rn(t, y.med)) > y.mad);
                        r[m] = keep ? t : 0.0f;
                        cnt += keep;
                        int i = 8;
                        for (; i + 24 < nfull; i += 32) {        // four blocks of 8 per batch: the loads go out together
                            float tt[4];
                            bool pr[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                pr[j] = present(v0 + i + 8 * j);
                                tt[j] = pr[j] ? __ldg(q + (j + 1) * step8) : CUDART_NAN_F;
                            }
                            q += 4 * step8;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {        // same order of additions as one by one
                                keep = (tt[j] == tt[j]) && !(fabsf(__fsub_rn(tt[j], y.med)) > y.mad);
                                if (pr[j]) r[m] = __fadd_rn(r[m], keep ? tt[j] : 0.0f);   // absent view: + 0, skipped
                                cnt += keep;
                            }
                        }
                        for (; i < nfull; i += 8) {
                            q += step8;
                            if (!present(v0 + i)) continue;       // + 0: no-op
                            t = __ldg(q);
                            keep = (t == t) && !(fabsf(__fsub_rn(t, y.med)) > y.mad);
                            r[m] = __fadd_rn(r[m], keep ? t : 0.0f);
                            cnt += keep;
                        }
                    }
                }
                // gather r[0..7] (accumulator a lives in lane a % AL, slot a / AL) and combine as numpy does:
                // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7))
                float a8[8];
#pragma unroll
                for (int a = 0; a < 8; ++a) a8[a] = __shfl_sync(gmask, r[a / AL], a % AL, LANES);
                leaf = __fadd_rn(__fadd_rn(__fadd_rn(a8[0], a8[1]), __fadd_rn(a8[2], a8[3])),
                                 __fadd_rn(__fadd_rn(a8[4], a8[5]), __fadd_rn(a8[6], a8[7])));
                for (int i = nfull; i < n; ++i) {
                    if (!present(off + i)) continue;
                    leaf = __fadd_rn(leaf, y(off + i));
                    if (lane == 0) cnt += y.kept(off + i);
                }
            }
            // ---- combine with finished left siblings
            float val = leaf;
            int l = lvl;
            while (sp > 0 && stack_lvl[sp - 1] == l) {
                val = __fadd_rn(stack_val[sp - 1], val);
                --sp;
                --l;
            }
            stack_val[sp] = val;
            stack_lvl[sp] = l;
            ++sp;
        }
        total = stack_val[0];
    }
    cnt = group_sum<LANES>(cnt, gmask);
    return __fdiv_rn(total, (float)cnt);
}

template <int LANES, int NVL>
__global__ void __launch_bounds__(kLargeThreads)
k_fuse_large(const float* __restrict__ views, int64_t plane_stride, int V, int64_t n_cells, int VS,
             float* __restrict__ out) {
    constexpr int CELLS = kLargeThreads / LANES;  // cells per CTA
    extern __shared__ uint32_t s_tile[];          // CELLS x VS words: floats first, then keys
    const int tid = threadIdx.x;
    const int64_t cell0 = blockIdx.x * (int64_t)CELLS;

    // 1. coalesced load: consecutive threads read consecutive cells of one view plane.  Thread t owns cell t % CELLS and
    //    views t / CELLS, + kLargeThreads / CELLS, ...; the loads of a batch of kLoadBatch views are all issued before
    //    the first result is stored (the kernel has ~22 warps per SM to hide the latency of this phase with).
    {
        constexpr int VSTEP = kLargeThreads / CELLS;       // views between two loads of a thread
        constexpr int kLoadBatch = 8;
        const int c = tid % CELLS;
        const bool cell_ok = cell0 + c < n_cells;
        const float* __restrict__ src = views + cell0 + c;
        uint32_t* __restrict__ dst = s_tile + c * VS;
        for (int v0 = tid / CELLS; v0 < V; v0 += VSTEP * kLoadBatch) {
            float t[kLoadBatch];
#pragma unroll
            for (int j = 0; j < kLoadBatch; ++j) {
                const int v = v0 + j * VSTEP;
                t[j] = CUDART_NAN_F;
                if (cell_ok && v < V) t[j] = __ldg(src + (int64_t)v * plane_stride);
            }
#pragma unroll
            for (int j = 0; j < kLoadBatch; ++j) {
                const int v = v0 + j * VSTEP;
                if (v < V) dst[v] = __float_as_uint(t[j]);
            }
        }
    }
    __syncthreads();

    const int c_local = tid / LANES, lane = tid % LANES;
    const int64_t cell = cell0 + c_local;
    const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << ((tid & 31) / LANES * LANES));
    uint32_t* run = s_tile + c_local * VS + lane;  // element i of this lane's run: run[i * LANES]

    // 2. per-lane sorted run
    float s[NVL];
    int k_lane = 0;
#pragma unroll
    for (int i = 0; i < NVL; ++i) {
        const int v = lane + i * LANES;
        float t = CUDART_INF_F;
        if (v < V) {
            const float x = __uint_as_float(run[i * LANES]);
            if (x == x) {
                t = x;
                ++k_lane;
            }
        }
        s[i] = t;
    }
    const int k = group_sum<LANES>(k_lane, gmask);
    if (cell >= n_cells) return;       // whole group leaves together
    if (k <= 2) {                      // aggregate_2p5d.py:69-71
        if (lane == 0) out[cell] = CUDART_NAN_F;
        return;
    }
    SortNet<NVL>::sort(s);
#pragma unroll
    for (int i = 0; i < NVL; ++i) run[i * LANES] = vs_key32(s[i]);
#pragma unroll
    for (int i = NVL; i < RunPad<NVL>::value; ++i) run[i * LANES] = 0xffffffffu;   // pads: never below any threshold
    __syncwarp(gmask);

    // 3. median
    const int ilo = (k - 1) >> 1, ihi = k >> 1;
    uint32_t khi, klo;
    group_select_split<LANES, NVL>(run, ihi, ilo != ihi, gmask, khi, klo);
    const float med = __fdiv_rn(__fadd_rn(vs_unkey32(klo), vs_unkey32(khi)), 2.0f);
    __syncwarp(gmask);

    // 4. MAD: |sorted run - med| is bitonic per lane -> merge in registers -> runs of keys again
#pragma unroll
    for (int i = 0; i < NVL; ++i) s[i] = fabsf(__fsub_rn(s[i], med));
    bitonic_merge_regs<NVL>(s);
#pragma unroll
    for (int i = 0; i < NVL; ++i) run[i * LANES] = vs_key32(s[i]);
    __syncwarp(gmask);
    group_select_split<LANES, NVL>(run, ihi, ilo != ihi, gmask, khi, klo);   // deviations are >= 0
    const float mad = __fdiv_rn(__fadd_rn(vs_unkey32(klo), vs_unkey32(khi)), 2.0f);

    // 5. nanmean of the survivors in numpy's pairwise order
    const float mean = sum_survivors<LANES>(views, cell, plane_stride, V, med, mad, lane, gmask, AllPresent());
    if (lane == 0) out[cell] = mean;
}

// ---------------------------------------------------------------------------------------------------------
// Sparse fusion (vs_fuse_views_sparse): large AOIs, where a view covers a fraction of the grid.
//
// Stage B records per 64x32 tile which views hold data there (occupancy bitmap, vs_set_occupancy / vs_exchange.occ).
// A planning kernel bins the tiles by their view count; one kernel per bin then runs the register-network path
// (or the multi-lane path above 128 views) over that bin's tiles, gathering only the marked planes: an unmarked
// (tile, view) pair is NaN by definition and its memory is never read.  The survivors are summed in numpy's order
// over the ORIGINAL view axis (the tree depends on V and on the view indices, not on how many views are present),
// with absent views skipped -- they would add +0.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSparseBins = 19;
__constant__ int c_bin_cap[kSparseBins] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 208, 256, 416, 512, 1024, 2048};
static const int h_bin_cap[kSparseBins] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 208, 256, 416, 512, 1024, 2048};
constexpr int kMaxOccWords = 64;   // 2048 views
constexpr int kMaxLeaves = 32;     // 2048 views / 64-element minimum leaf

struct SparseGeom {
    const float* views;
    int64_t plane_stride;
    int V, rows, W, row0;      // planes hold grid rows [row0, row0 + rows)
    const uint32_t* occ;
    int occ_words, tiles_x;
    int ty_first, n_units;     // tile rows ty_first .. ; units = tile rows x tile columns
    const int* bin_count;      // [kSparseBins]
    const int* bin_list;       // [kSparseBins][n_units]
    float* out;
    // numpy's pairwise-sum tree over the view axis [0, V), leaves left to right (host-built, make_sum_tree):
    // leaf l covers views [leaf_off[l], leaf_off[l] + leaf_n[l]); after it, leaf_pops[l] partial sums are folded
    int n_leaves;
    short leaf_off[kMaxLeaves], leaf_n[kMaxLeaves];
    signed char leaf_pops[kMaxLeaves];
};

// Leaves of numpy's float32 pairwise add.reduce over n elements: blocks of <= 128 elements, split at
// n/2 rounded down to a multiple of 8.  pops[l] = how many finished left siblings are added after leaf l.
static void make_sum_tree(int V, SparseGeom& g) {
    struct Seg { int off, n, lvl; };
    Seg todo[64];
    int tp = 0, nl = 0;
    int lvl_stack[32], sp = 0;
    todo[tp++] = {0, V, 0};
    while (tp > 0) {
        const Seg sgm = todo[--tp];
        if (sgm.n > 128) {
            int n2 = sgm.n >> 1;
            n2 -= n2 & 7;
            todo[tp++] = {sgm.off + n2, sgm.n - n2, sgm.lvl + 1};
            todo[tp++] = {sgm.off, n2, sgm.lvl + 1};
            continue;
        }
        int pops = 0, l = sgm.lvl;
        while (sp > 0 && lvl_stack[sp - 1] == l) { --sp; --l; ++pops; }
        lvl_stack[sp++] = l;
        g.leaf_off[nl] = (short)sgm.off;
        g.leaf_n[nl] = (short)sgm.n;
        g.leaf_pops[nl] = (signed char)pops;
        ++nl;
    }
    g.n_leaves = nl;
}

__global__ void k_fuse_plan(const uint32_t* __restrict__ occ, int occ_words, int tiles_x, int ty_first, int n_units,
                            int* __restrict__ bin_count, int* __restrict__ bin_list) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const int ty = ty_first + u / tiles_x, tx = u - (u / tiles_x) * tiles_x;
    const uint32_t* w = occ + ((size_t)ty * tiles_x + tx) * occ_words;
    int n = 0;
    for (int i = 0; i < occ_words; ++i) n += __popc(w[i]);
    int b = 0;
    while (b < kSparseBins - 1 && n > c_bin_cap[b]) ++b;
    bin_list[(size_t)b * n_units + atomicAdd(bin_count + b, 1)] = u;
}

// One warp turns the tile's occupancy words into the ascending list of its views (shared memory).
__device__ __forceinline__ int build_view_list(const uint32_t* __restrict__ occ_tile, int occ_words, uint32_t* s_bits,
                                               int* s_list, int cap) {
    __shared__ int s_n;
    const int tid = threadIdx.x;
    if (tid < 32) {
        int base = 0;
        for (int w0 = 0; w0 < occ_words; w0 += 32) {
            const int w = w0 + tid;
            const uint32_t bits = w < occ_words ? occ_tile[w] : 0u;
            if (w < kMaxOccWords) s_bits[w] = bits;
            const int c = __popc(bits);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            int pos = base + incl - c;
            uint32_t b = bits;
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                if (pos < cap) s_list[pos] = w * 32 + bit;
                ++pos;
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (tid == 0) s_n = base;
    }
    __syncthreads();
    return s_n;
}

// numpy's pairwise sum over the ORIGINAL view axis [0, V), driven by the tile's view list.  Absent views would add +0,
// so only the listed ones are visited.  The list is regrouped once per tile (build_sum_groups, one warp): entries of
// leaf l that fall into numpy's accumulator r[j] (j = (v - leaf offset) % 8 inside the leaf's full blocks of 8) form
// group 9 l + j, the leaf's tail elements group 9 l + 8, every group in ascending view order (stable) -- the order
// numpy adds them in.  The per-cell loop then runs over contiguous ranges with the accumulator in a fixed register.
constexpr int kMaxGroups = kMaxLeaves * 9;

struct SumGroups {
    const int* perm;             // shared: plane indices regrouped
    const unsigned short* gend;  // shared: end offset of every group (cumulative)
};

// s_perm: >= n ints, s_gend: kMaxGroups ushorts, s_cnt: kMaxGroups ints (scratch).  All threads call it; one warp works.
__device__ __forceinline__ void build_sum_groups(const SparseGeom& g, const int* __restrict__ s_list, int n, int* s_perm,
                                                 unsigned short* s_gend, int* s_cnt) {
    const int tid = threadIdx.x;
    const int n_groups = g.n_leaves * 9;
    for (int i = tid; i < n_groups; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    if (tid < 32) {
        const unsigned lt = (1u << tid) - 1u;
        // pass 1: group sizes
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + tid;
            if (i < n) {
                const int v = s_list[i];
                int L = 0;
                while (L + 1 < g.n_leaves && v >= g.leaf_off[L] + g.leaf_n[L]) ++L;
                const int pos = v - g.leaf_off[L], nl = g.leaf_n[L];
                const int j = (nl >= 8 && pos < nl - (nl & 7)) ? (pos & 7) : 8;
                atomicAdd(&s_cnt[L * 9 + j], 1);
            }
        }
        __syncwarp();
        // exclusive prefix over the groups -> s_cnt becomes the running write position, s_gend the end offsets
        int base = 0;
        for (int q0 = 0; q0 < n_groups; q0 += 32) {
            const int q = q0 + tid;
            const int c = q < n_groups ? s_cnt[q] : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            if (q < n_groups) {
                s_cnt[q] = base + incl - c;
                s_gend[q] = (unsigned short)(base + incl);
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        // pass 2: stable placement, 32 entries at a time (entries of one group keep their ascending order)
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + tid;
            int key = -1 - tid, v = 0;       // inactive lanes: unique keys
            if (i < n) {
                v = s_list[i];
                int L = 0;
                while (L + 1 < g.n_leaves && v >= g.leaf_off[L] + g.leaf_n[L]) ++L;
                const int pos = v - g.leaf_off[L], nl = g.leaf_n[L];
                key = L * 9 + ((nl >= 8 && pos < nl - (nl & 7)) ? (pos & 7) : 8);
            }
            const unsigned m = __match_any_sync(0xffffffffu, key);
            if (i < n) {
                const int rank = __popc(m & lt);
                s_perm[s_cnt[key] + rank] = v;
            }
            __syncwarp();
            if (i < n && (m & lt) == 0) s_cnt[key] += __popc(m);      // the group's first lane advances its position
            __syncwarp();
        }
    }
    __syncthreads();
}

// One thread sums accumulators [J0, J1) of every leaf (and, if TAIL, adds the tails and folds the leaves); PART = the
// whole sum when J0 = 0, J1 = 8, TAIL = true.  Returns the partial block sums through r_out when !TAIL.
__device__ __forceinline__ float kept_value(const float* __restrict__ cellp, int64_t stride, int v, float med, float mad,
                                            int& cnt) {
    const float t = __ldg(cellp + (int64_t)v * stride);
    const bool keep = (t == t) && !(fabsf(__fsub_rn(t, med)) > mad);   // aggregate_2p5d.py:76-77
    cnt += keep;
    return keep ? t : 0.0f;
}

__device__ __forceinline__ float sparse_sum_1t(const SparseGeom& g, const SumGroups& sg, const float* __restrict__ cellp,
                                               float med, float mad) {
    int e = 0, cnt = 0, sp = 0;
    float st[8];
    for (int L = 0; L < g.n_leaves; ++L) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r[j] = 0.0f;
            const int end = sg.gend[L * 9 + j];
            for (; e < end; ++e) r[j] = __fadd_rn(r[j], kept_value(cellp, g.plane_stride, sg.perm[e], med, mad, cnt));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        const int end = sg.gend[L * 9 + 8];
        for (; e < end; ++e) res = __fadd_rn(res, kept_value(cellp, g.plane_stride, sg.perm[e], med, mad, cnt));
        for (int q = 0; q < g.leaf_pops[L]; ++q) res = __fadd_rn(st[--sp], res);
        st[sp++] = res;
    }
    return __fdiv_rn(st[0], (float)cnt);
}

// The same sum by the two lanes of a pair: lane A owns accumulators 0..3, lane B 4..7 (numpy combines
// ((r0+r1)+(r2+r3)) + ((r4+r5)+(r6+r7)), so each lane reduces its half and A adds B's); A adds the tails and folds.
// Returns the mean on lane A.
__device__ __forceinline__ float sparse_sum_pair(const SparseGeom& g, const SumGroups& sg, const float* __restrict__ cellp,
                                                 float med, float mad, bool is_b, unsigned pm) {
    int cnt = 0, sp = 0;
    float st[8];
    for (int L = 0; L < g.n_leaves; ++L) {
        const int j0 = is_b ? 4 : 0;
        int e = (L * 9 + j0) > 0 ? sg.gend[L * 9 + j0 - 1] : 0;
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            r[j] = 0.0f;
            const int end = sg.gend[L * 9 + j0 + j];
            for (; e < end; ++e) r[j] = __fadd_rn(r[j], kept_value(cellp, g.plane_stride, sg.perm[e], med, mad, cnt));
        }
        const float half = __fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3]));
        const float other = __shfl_xor_sync(pm, half, 1);
        float res = __fadd_rn(half, other);              // on A: (A's half) + (B's half), numpy's order
        if (!is_b) {
            e = sg.gend[L * 9 + 7];
            const int end = sg.gend[L * 9 + 8];
            for (; e < end; ++e) res = __fadd_rn(res, kept_value(cellp, g.plane_stride, sg.perm[e], med, mad, cnt));
            for (int q = 0; q < g.leaf_pops[L]; ++q) res = __fadd_rn(st[--sp], res);
            st[sp++] = res;
        }
    }
    cnt += __shfl_xor_sync(pm, cnt, 1);
    return __fdiv_rn(st[0], (float)cnt);
}

// rows of unit u (a tile clipped to the planes' rows), split into 4 slabs of 8 tile rows
__device__ __forceinline__ void unit_geometry(const SparseGeom& g, int u, int& ty, int& tx) {
    ty = g.ty_first + u / g.tiles_x;
    tx = u - (u / g.tiles_x) * g.tiles_x;
}

template <int NV>
__global__ void __launch_bounds__(kBlockSmall, NV <= 32 ? 6 : NV <= 64 ? 5 : NV <= 112 ? 4 : 2)
k_fuse_sparse_regs(const __grid_constant__ SparseGeom g, int bin, int sl) {
    __shared__ uint32_t s_bits[kMaxOccWords];
    __shared__ int s_list[NV], s_perm[NV], s_cnt[kMaxGroups];
    __shared__ unsigned short s_gend[kMaxGroups];
    const SumGroups sg{s_perm, s_gend};
    const int n_items = g.bin_count[bin] << sl;      // 2^sl slabs of (32 >> sl) rows per tile
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int u = g.bin_list[(size_t)bin * g.n_units + (item >> sl)];
        int ty, tx;
        unit_geometry(g, u, ty, tx);
        __syncthreads();   // the previous item's readers of the shared lists are done
        const int n = build_view_list(g.occ + ((size_t)ty * g.tiles_x + tx) * g.occ_words, g.occ_words, s_bits, s_list, NV);
        build_sum_groups(g, s_list, n, s_perm, s_gend, s_cnt);
        const int slab = item & ((1 << sl) - 1), slab_rows = VS_TILE_H >> sl;
        const int gy0 = max(ty * VS_TILE_H + slab * slab_rows, g.row0);
        const int gy1 = min(min(ty * VS_TILE_H + (slab + 1) * slab_rows, g.row0 + g.rows), (ty + 1) * VS_TILE_H);
        for (int c = threadIdx.x; c < VS_TILE_W * (gy1 - gy0); c += kBlockSmall) {
            const int gy = gy0 + (c >> 6), gx = tx * VS_TILE_W + (c & 63);
            if (gx >= g.W) continue;
            const int64_t cell = (int64_t)(gy - g.row0) * g.W + gx;
            float s[NV];
            int k = 0;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                float t = CUDART_NAN_F;
                if (i < n) t = __ldg(g.views + (int64_t)s_list[i] * g.plane_stride + cell);
                const bool ok = (t == t);
                k += ok;
                s[i] = ok ? t : CUDART_INF_F;
            }
            if (k <= 2) {  // aggregate_2p5d.py:69-71
                g.out[cell] = CUDART_NAN_F;
                continue;
            }
            SortNet<NV>::sort(s);
            const float med = middle_of_sorted<NV>(s, k);
#pragma unroll
            for (int i = 0; i < NV; ++i) s[i] = fabsf(__fsub_rn(s[i], med));
            bitonic_merge_regs<NV>(s);
            const float mad = middle_of_sorted<NV>(s, k);
            g.out[cell] = sparse_sum_1t(g, sg, g.views + cell, med, mad);
        }
    }
}

// More than 128 views in a tile: the multi-lane path of k_fuse_large over the tile's view list.
template <int LANES, int NVL>
__global__ void __launch_bounds__(kLargeThreads)
k_fuse_sparse_large(const __grid_constant__ SparseGeom g, int bin, int VS, int sl) {
    constexpr int CELLS = kLargeThreads / LANES;   // cells per pass: 64, 32 or 8 consecutive columns of one tile row
    constexpr int PASSES_PER_ROW = VS_TILE_W / CELLS;
    extern __shared__ uint32_t s_tile[];           // CELLS x VS words
    __shared__ uint32_t s_bits[kMaxOccWords];
    __shared__ int s_list[LANES * NVL];
    const int tid = threadIdx.x;
    const int n_items = g.bin_count[bin] << sl;      // 2^sl slabs of (32 >> sl) rows per tile
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int u = g.bin_list[(size_t)bin * g.n_units + (item >> sl)];
        int ty, tx;
        unit_geometry(g, u, ty, tx);
        __syncthreads();
        const int n = build_view_list(g.occ + ((size_t)ty * g.tiles_x + tx) * g.occ_words, g.occ_words, s_bits, s_list,
                                      LANES * NVL);
        const int slab = item & ((1 << sl) - 1), slab_rows = VS_TILE_H >> sl;
        const int gy0 = max(ty * VS_TILE_H + slab * slab_rows, g.row0);
        const int gy1 = min(min(ty * VS_TILE_H + (slab + 1) * slab_rows, g.row0 + g.rows), (ty + 1) * VS_TILE_H);
        const BitsPresent present{s_bits};
        for (int pass = 0; pass < (gy1 - gy0) * PASSES_PER_ROW; ++pass) {
            const int gy = gy0 + pass / PASSES_PER_ROW;
            const int gx0 = tx * VS_TILE_W + (pass % PASSES_PER_ROW) * CELLS;
            if (gx0 >= g.W) continue;                       // block-uniform
            const int64_t cell0 = (int64_t)(gy - g.row0) * g.W + gx0;
            const int n_cols = min(CELLS, g.W - gx0);
            __syncthreads();                                // previous pass done with s_tile
            {   // loads of a batch of listed views in flight before the first store (see k_fuse_large)
                constexpr int VSTEP = kLargeThreads / CELLS;
                constexpr int kLoadBatch = 8;
                const int c = tid % CELLS;
                const bool cell_ok = c < n_cols;
                const float* __restrict__ src = g.views + cell0 + c;
                uint32_t* __restrict__ dst = s_tile + c * VS;
                for (int s0 = tid / CELLS; s0 < n; s0 += VSTEP * kLoadBatch) {
                    float t[kLoadBatch];
#pragma unroll
                    for (int j = 0; j < kLoadBatch; ++j) {
                        const int slot = s0 + j * VSTEP;
                        t[j] = CUDART_NAN_F;
                        if (cell_ok && slot < n) t[j] = __ldg(src + (int64_t)s_list[slot] * g.plane_stride);
                    }
#pragma unroll
                    for (int j = 0; j < kLoadBatch; ++j) {
                        const int slot = s0 + j * VSTEP;
                        if (slot < n) dst[slot] = __float_as_uint(t[j]);
                    }
                }
            }
            __syncthreads();
            const int c_local = tid / LANES, lane = tid % LANES;
            const int64_t cell = cell0 + c_local;
            const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << ((tid & 31) / LANES * LANES));
            uint32_t* run = s_tile + c_local * VS + lane;
            float s[NVL];
            int k_lane = 0;
#pragma unroll
            for (int i = 0; i < NVL; ++i) {
                const int slot = lane + i * LANES;
                float t = CUDART_INF_F;
                if (slot < n) {
                    const float x = __uint_as_float(run[i * LANES]);
                    if (x == x) {
                        t = x;
                        ++k_lane;
                    }
                }
                s[i] = t;
            }
            const int k = group_sum<LANES>(k_lane, gmask);
            if (c_local >= n_cols) continue;     // whole group skips together (no block barrier below)
            if (k <= 2) {
                if (lane == 0) g.out[cell] = CUDART_NAN_F;
                continue;
            }
            SortNet<NVL>::sort(s);
#pragma unroll
            for (int i = 0; i < NVL; ++i) run[i * LANES] = vs_key32(s[i]);
#pragma unroll
            for (int i = NVL; i < RunPad<NVL>::value; ++i) run[i * LANES] = 0xffffffffu;
            __syncwarp(gmask);
            const int ilo = (k - 1) >> 1, ihi = k >> 1;
            uint32_t khi, klo;
            group_select_split<LANES, NVL>(run, ihi, ilo != ihi, gmask, khi, klo);
            const float med = __fdiv_rn(__fadd_rn(vs_unkey32(klo), vs_unkey32(khi)), 2.0f);
            __syncwarp(gmask);
#pragma unroll
            for (int i = 0; i < NVL; ++i) s[i] = fabsf(__fsub_rn(s[i], med));
            bitonic_merge_regs<NVL>(s);
#pragma unroll
            for (int i = 0; i < NVL; ++i) run[i * LANES] = vs_key32(s[i]);
            __syncwarp(gmask);
            group_select_split<LANES, NVL>(run, ihi, ilo != ihi, gmask, khi, klo);
            const float mad = __fdiv_rn(__fadd_rn(vs_unkey32(klo), vs_unkey32(khi)), 2.0f);
            const float mean = sum_survivors<LANES>(g.views, cell, g.plane_stride, g.V, med, mad, lane, gmask, present);
            if (lane == 0) g.out[cell] = mean;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 129..208 views per cell: TWO threads per cell (adjacent lanes), each sorting up to NVL values in registers with the
// same networks as the one-thread path; the order statistics of the union come from one cross step -- thread A keeps
// min(a[i], b[NVL-1-i]) (the NVL smallest of the union, a bitonic sequence), thread B the max (the NVL largest) -- and
// one bitonic merge per thread.  |x - med| is bitonic over A's sorted half and already ascending over B's, so the MAD
// costs one more merge + cross step + merge.  No shared memory, no bisection.  (The multi-lane shared-memory kernel
// above remains for more than 208 views.)
// ---------------------------------------------------------------------------------------------------------
constexpr int kPairThreads = 128;   // 64 cells per CTA pass

// Bitonic merge into DESCENDING order of a "mountain" (rises, then falls) held in N registers; N is padded to a power
// of two with virtual -inf wires at the end, which continue the falling part and never move.
template <int N>
__device__ __forceinline__ void bitonic_merge_regs_desc(float (&a)[N]) {
    constexpr int N2 = N <= 8 ? 8 : N <= 16 ? 16 : N <= 32 ? 32 : N <= 64 ? 64 : 128;
#pragma unroll
    for (int j = N2 / 2; j > 0; j >>= 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int l = i ^ j;
            if (l > i && l < N) {
                const float hi = fmaxf(a[i], a[l]);
                const float lo = fminf(a[i], a[l]);
                a[i] = hi;
                a[l] = lo;
            }
        }
    }
}

template <int NVL>
__device__ __forceinline__ float pick_reg(const float (&s)[NVL], int idx) {
    float r = s[0];
#pragma unroll
    for (int i = 1; i < NVL; ++i) r = (i == idx) ? s[i] : r;
    return r;
}

// One cell, computed by the two lanes of a pair (lane A even, lane B odd); med / mad are returned on both lanes.
// slot(i) -> plane index of the i-th listed view (i < n), cellp = views + cell.
//
// A takes the even list slots, B the odd ones, and B works on NEGATED values: after both sorted ascending,
// B's element i is -(its (NVL-1-i)-th smallest), so the classic split "A keeps min(a[i], b[NVL-1-i]), B keeps the max"
// becomes the SAME instruction on both lanes, t[i] = min(own[i], -partner[i]), with no index reversal; both results are
// mountains, sorted by the same descending merge.  A then holds the NVL smallest values of the union in descending
// order (rank r at index NVL-1-r), B minus the NVL largest (the smallest of them at index 0).
template <int NVL, typename SlotFn>
__device__ __forceinline__ void fuse_cell_pair(const SlotFn& slot, int n, const float* __restrict__ cellp, int64_t stride,
                                               bool is_b, unsigned pm, float& med_out, float& mad_out, int& k_out) {
    const int lane_a = (threadIdx.x & 31) & ~1;
    const uint32_t sign = is_b ? 0x80000000u : 0u;
    float s[NVL];
    int k_lane = 0;
#pragma unroll
    for (int i = 0; i < NVL; ++i) {
        const int sl = 2 * i + (is_b ? 1 : 0);
        float t = CUDART_NAN_F;
        if (sl < n) t = __ldg(cellp + (int64_t)slot(sl) * stride);
        const bool ok = (t == t);
        k_lane += ok;
        s[i] = __uint_as_float(__float_as_uint(ok ? t : CUDART_INF_F) ^ sign);     // B: negated
    }
    const int k = k_lane + __shfl_xor_sync(pm, k_lane, 1);
    k_out = k;
    if (k <= 2) return;                      // both lanes of the pair agree (aggregate_2p5d.py:69-71)
    SortNet<NVL>::sort(s);
    float t2[NVL];
#pragma unroll
    for (int i = 0; i < NVL; ++i) t2[i] = fminf(s[i], -__shfl_xor_sync(pm, s[i], 1));
    bitonic_merge_regs_desc<NVL>(t2);
    const int ilo = (k - 1) >> 1, ihi = k >> 1;        // ranks in the union; ilo <= NVL - 1 because k <= 2 NVL
    float up0 = -__shfl_xor_sync(pm, t2[0], 1);        // on A: the smallest of the upper half
    float lo = pick_reg<NVL>(t2, NVL - 1 - ilo);
    float hi = (ihi < NVL) ? pick_reg<NVL>(t2, NVL - 1 - ihi) : up0;
    float med = __shfl_sync(pm, __fdiv_rn(__fadd_rn(lo, hi), 2.0f), lane_a);
    // |x - med|: on A a valley over the descending half (falls to the median, rises below it; +inf pads in front),
    // on B (true values -t2[i] >= med, ascending in i) already ascending.  Same instructions on both lanes.
    const float smed = __uint_as_float(__float_as_uint(med) ^ sign);
#pragma unroll
    for (int i = 0; i < NVL; ++i) s[i] = fabsf(__fsub_rn(t2[i], smed));
    bitonic_merge_regs<NVL>(s);                          // ascending (+inf pads); a no-op on B's ascending run
    // second split: A keeps its ascending run, B turns its run into minus the reversed run
#pragma unroll
    for (int i = 0; i < NVL; ++i) t2[i] = is_b ? -s[NVL - 1 - i] : s[i];
#pragma unroll
    for (int i = 0; i < NVL; ++i) s[i] = fminf(t2[i], -__shfl_xor_sync(pm, t2[i], 1));
    bitonic_merge_regs_desc<NVL>(s);
    up0 = -__shfl_xor_sync(pm, s[0], 1);
    lo = pick_reg<NVL>(s, NVL - 1 - ilo);
    hi = (ihi < NVL) ? pick_reg<NVL>(s, NVL - 1 - ihi) : up0;
    mad_out = __shfl_sync(pm, __fdiv_rn(__fadd_rn(lo, hi), 2.0f), lane_a);
    med_out = med;
}

struct ListSlot {
    const int* list;
    __device__ __forceinline__ int operator()(int i) const { return list[i]; }
};
struct IdentitySlot {
    __device__ __forceinline__ int operator()(int i) const { return i; }
};

template <int NVL>
__global__ void __launch_bounds__(kPairThreads, 4)
k_fuse_sparse_pair(const __grid_constant__ SparseGeom g, int bin, int sl) {
    __shared__ uint32_t s_bits[kMaxOccWords];
    __shared__ int s_list[2 * NVL], s_perm[2 * NVL], s_cnt[kMaxGroups];
    __shared__ unsigned short s_gend[kMaxGroups];
    const SumGroups sg{s_perm, s_gend};
    const int n_items = g.bin_count[bin] << sl;      // 2^sl slabs of (32 >> sl) rows per tile
    const bool is_b = threadIdx.x & 1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int u = g.bin_list[(size_t)bin * g.n_units + (item >> sl)];
        int ty, tx;
        unit_geometry(g, u, ty, tx);
        __syncthreads();
        const int n = build_view_list(g.occ + ((size_t)ty * g.tiles_x + tx) * g.occ_words, g.occ_words, s_bits, s_list, 2 * NVL);
        build_sum_groups(g, s_list, n, s_perm, s_gend, s_cnt);
        const int slab = item & ((1 << sl) - 1), slab_rows = VS_TILE_H >> sl;
        const int gy0 = max(ty * VS_TILE_H + slab * slab_rows, g.row0);
        const int gy1 = min(min(ty * VS_TILE_H + (slab + 1) * slab_rows, g.row0 + g.rows), (ty + 1) * VS_TILE_H);
        const ListSlot slot{s_list};
        const unsigned pm = 3u << ((threadIdx.x & 31) & ~1);      // the two lanes of this cell
        for (int row = gy0; row < gy1; ++row) {
            const int gx = tx * VS_TILE_W + (threadIdx.x >> 1);
            if (gx >= g.W) continue;                     // the pair leaves together
            const int64_t cell = (int64_t)(row - g.row0) * g.W + gx;
            float med = 0.f, mad = 0.f;
            int k = 0;
            fuse_cell_pair<NVL>(slot, n, g.views + cell, g.plane_stride, is_b, pm, med, mad, k);
            if (k <= 2) {
                if (!is_b) g.out[cell] = CUDART_NAN_F;
                continue;
            }
            const float mean = sparse_sum_pair(g, sg, g.views + cell, med, mad, is_b, pm);
            if (!is_b) g.out[cell] = mean;
        }
    }
}

// dense stacks, 129..208 views: the same pair path with the identity list
template <int NVL>
__global__ void __launch_bounds__(kPairThreads, 4)
k_fuse_pair(const __grid_constant__ SparseGeom g, int64_t n_cells) {
    __shared__ int s_list[2 * NVL], s_perm[2 * NVL], s_cnt[kMaxGroups];
    __shared__ unsigned short s_gend[kMaxGroups];
    for (int i = threadIdx.x; i < 2 * NVL; i += kPairThreads) s_list[i] = i;
    __syncthreads();
    build_sum_groups(g, s_list, g.V, s_perm, s_gend, s_cnt);
    const SumGroups sg{s_perm, s_gend};
    const bool is_b = threadIdx.x & 1;
    const int64_t cell = blockIdx.x * (int64_t)(kPairThreads / 2) + (threadIdx.x >> 1);
    if (cell >= n_cells) return;
    float med = 0.f, mad = 0.f;
    int k = 0;
    const unsigned pm = 3u << ((threadIdx.x & 31) & ~1);
    fuse_cell_pair<NVL>(IdentitySlot(), g.V, g.views + cell, g.plane_stride, is_b, pm, med, mad, k);
    if (k <= 2) {
        if (!is_b) g.out[cell] = CUDART_NAN_F;
        return;
    }
    const float mean = sparse_sum_pair(g, sg, g.views + cell, med, mad, is_b, pm);
    if (!is_b) g.out[cell] = mean;
}

template <int NVL>
int launch_sparse_pair(vs_ctx* ctx, const SparseGeom& g, int bin, cudaStream_t stream) {
    const int blocks = std::min(g.n_units << 2, ctx->sm_count * 8);
    k_fuse_sparse_pair<NVL><<<blocks, kPairThreads, 0, stream>>>(g, bin, 2);
    VS_CHECK_LAUNCH(ctx, "k_fuse_sparse_pair");
    return VS_OK;
}

template <int NVL>
int launch_pair(vs_ctx* ctx, const float* views, int64_t plane_stride, int V, int64_t n_cells, float* out,
                cudaStream_t stream) {
    SparseGeom g;
    memset(&g, 0, sizeof(g));
    g.views = views;
    g.plane_stride = plane_stride;
    g.V = V;
    g.out = out;
    make_sum_tree(V, g);
    const int64_t blocks = (n_cells + kPairThreads / 2 - 1) / (kPairThreads / 2);
    k_fuse_pair<NVL><<<(unsigned)blocks, kPairThreads, 0, stream>>>(g, n_cells);
    VS_CHECK_LAUNCH(ctx, "k_fuse_pair");
    return VS_OK;
}

// Work items are slabs of a tile: 2^sl slabs of (32 >> sl) rows.  Measured on C3 (200 views, 8192^2 cells): 4 slabs of 8
// rows for every bin 34.6 ms; thinner slabs for the bins with many views (8 / 16 per tile) plus the bins spread over three
// streams 36.2 ms -- the per-item list building costs more than the shorter tails save.
template <int NV>
int launch_sparse_regs(vs_ctx* ctx, const SparseGeom& g, int bin, cudaStream_t stream) {
    // persistent grid: the number of tiles of this bin is only known on the device
    constexpr int sl = 2;
    const int blocks = std::min(g.n_units << sl, ctx->sm_count * 8);
    k_fuse_sparse_regs<NV><<<blocks, kBlockSmall, 0, stream>>>(g, bin, sl);
    VS_CHECK_LAUNCH(ctx, "k_fuse_sparse_regs");
    return VS_OK;
}

template <int LANES, int NVL>
int launch_sparse_large(vs_ctx* ctx, const SparseGeom& g, int bin, cudaStream_t stream) {
    constexpr int CELLS = kLargeThreads / LANES;
    int VS = LANES * RunPad<NVL>::value;
    VS += (9 - (VS & 31) + 32) & 31;
    const size_t smem = (size_t)CELLS * VS * sizeof(uint32_t);
    VS_CUDA(cudaFuncSetAttribute(k_fuse_sparse_large<LANES, NVL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = std::min(g.n_units << 2, ctx->sm_count * 4);
    k_fuse_sparse_large<LANES, NVL><<<blocks, kLargeThreads, smem, stream>>>(g, bin, VS, 2);
    VS_CHECK_LAUNCH(ctx, "k_fuse_sparse_large");
    return VS_OK;
}

static int fuse_select_mode_once() {
    static const int rc = []() {
        const char* e = getenv("VISSAT_FUSE_SPLIT");
        const int v = (e && e[0] == '1') ? 1 : 0;
        return v ? (int)cudaMemcpyToSymbol(g_fuse_split, &v, sizeof(int)) : 0;
    }();
    return rc;
}

template <int NV>
int launch_small(vs_ctx* ctx, const float* views, int64_t plane_stride, int V, int64_t n_cells, float* out,
                 cudaStream_t stream) {
    const int64_t blocks = (n_cells + kBlockSmall - 1) / kBlockSmall;
    k_fuse_small<NV><<<(unsigned)blocks, kBlockSmall, 0, stream>>>(views, plane_stride, V, n_cells, out);
    VS_CHECK_LAUNCH(ctx, "k_fuse_small");
    return VS_OK;
}

template <int NV>
int launch_medium(vs_ctx* ctx, const float* views, int64_t plane_stride, int V, int64_t n_cells, float* out,
                  cudaStream_t stream) {
    const int64_t blocks = (n_cells + kBlockSmall - 1) / kBlockSmall;
    k_fuse_medium<NV><<<(unsigned)blocks, kBlockSmall, 0, stream>>>(views, plane_stride, V, n_cells, out);
    VS_CHECK_LAUNCH(ctx, "k_fuse_medium");
    return VS_OK;
}

template <int LANES, int NVL>
int launch_large_t(vs_ctx* ctx, const float* views, int64_t plane_stride, int V, int64_t n_cells, float* out,
                   cudaStream_t stream) {
    constexpr int CELLS = kLargeThreads / LANES;
    int VS = LANES * RunPad<NVL>::value;
    VS += (9 - (VS & 31) + 32) & 31;  // row stride = 9 (mod 32): conflict-free tile stores, few conflicts on the runs
    const size_t smem = (size_t)CELLS * VS * sizeof(uint32_t);
    fuse_select_mode_once();
    VS_CUDA(cudaFuncSetAttribute(k_fuse_large<LANES, NVL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (n_cells + CELLS - 1) / CELLS;
    k_fuse_large<LANES, NVL><<<(unsigned)blocks, kLargeThreads, smem, stream>>>(views, plane_stride, V, n_cells, VS, out);
    VS_CHECK_LAUNCH(ctx, "k_fuse_large");
    return VS_OK;
}

int launch_large(vs_ctx* ctx, const float* views, int64_t plane_stride, int V, int64_t n_cells, float* out,
                 cudaStream_t stream) {
    // A register-only variant (cross-lane bitonic merge tree with shuffles instead of the bisection) was measured
    // slower (1.43 vs 1.28 ms for 200 views x 1M cells): three times the min/max instructions, and FMNMX issues on the
    // half-rate ALU pipe.
#define VS_TRY(LANES, NVL) \
    if (V <= LANES * NVL) return launch_large_t<LANES, NVL>(ctx, views, plane_stride, V, n_cells, out, stream);
    VS_TRY(4, 40) VS_TRY(4, 48) VS_TRY(4, 52) VS_TRY(4, 56) VS_TRY(4, 64)
    VS_TRY(8, 40) VS_TRY(8, 48) VS_TRY(8, 52) VS_TRY(8, 56) VS_TRY(8, 64)
    VS_TRY(32, 24) VS_TRY(32, 32) VS_TRY(32, 48) VS_TRY(32, 64)
#undef VS_TRY
    vs_set_error("vs_fuse_views: more than 2048 views");
    return VS_ERR_INVALID;
}

}  // namespace

extern "C" int vs_fuse_views(vs_ctx* ctx, const float* views, int64_t plane_stride, int32_t n_views, int32_t rows,
                             int32_t W, float* out_mean, void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_fuse_views: NULL context");
    VS_REQUIRE(n_views >= 1 && n_views <= 2048, "vs_fuse_views: n_views must be 1..2048");
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
    // network sizes in steps of 4 above 32 views (the network for 52 wires has 9 % fewer comparators than the one
    // for 56), steps of 8 above 64
    // Above 32 views the one-array kernel (sorted copy in registers, originals re-read from L2) wins: 75..80 registers
    // and 24 warps per SM against 113..188 registers for the two-array kernel (measured, 2048^2 cells: V = 40
    // 356 vs 383 us, V = 50 482 vs 585 us, V = 64 653 vs 966 us).  Network sizes in steps of 4 up to 64 views (the
    // network for 52 wires has 9 % fewer comparators than the one for 56), steps of 8 above.
    if (V <= 36) return launch_medium<36>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 40) return launch_medium<40>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 44) return launch_medium<44>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 48) return launch_medium<48>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 52) return launch_medium<52>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 56) return launch_medium<56>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 60) return launch_medium<60>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 64) return launch_medium<64>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 80) return launch_medium<80>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 88) return launch_medium<88>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 96) return launch_medium<96>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 104) return launch_medium<104>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 112) return launch_medium<112>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 120) return launch_medium<120>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    if (V <= 128) return launch_medium<128>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    // 129..208 views: two threads per cell, register networks + one cross step (k_fuse_pair); beyond: multi-lane runs in
    // shared memory (k_fuse_large).  VISSAT_FUSE_PAIR=0 keeps the round-1 path for A/B measurements.
    static const bool use_pair = []() { const char* e = getenv("VISSAT_FUSE_PAIR"); return !(e && e[0] == '0'); }();
    if (use_pair) {
        if (V <= 160) return launch_pair<80>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
        if (V <= 176) return launch_pair<88>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
        if (V <= 192) return launch_pair<96>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
        if (V <= 208) return launch_pair<104>(ctx, views, plane_stride, V, n_cells, out_mean, stream);
    }
    return launch_large(ctx, views, plane_stride, V, n_cells, out_mean, stream);
}


extern "C" int vs_fuse_views_sparse(vs_ctx* ctx, const float* views, int64_t plane_stride, int32_t n_views, int32_t rows,
                                    int32_t W, int32_t row0, const uint32_t* occ, int32_t occ_words, float* out_mean,
                                    void* stream_) {
    VS_REQUIRE(ctx != nullptr, "vs_fuse_views_sparse: NULL context");
    VS_REQUIRE(n_views >= 1 && n_views <= 2048, "vs_fuse_views_sparse: n_views must be 1..2048");
    VS_REQUIRE(rows >= 0 && W >= 0 && row0 >= 0, "vs_fuse_views_sparse: negative size");
    const int64_t n_cells = (int64_t)rows * W;
    if (n_cells == 0) return VS_OK;
    VS_REQUIRE(views != nullptr && out_mean != nullptr && occ != nullptr, "vs_fuse_views_sparse: NULL array");
    VS_REQUIRE(plane_stride >= n_cells, "vs_fuse_views_sparse: plane_stride smaller than a plane");
    VS_REQUIRE(occ_words > 0 && occ_words <= kMaxOccWords && (int64_t)occ_words * 32 >= n_views,
               "vs_fuse_views_sparse: occ_words must cover n_views (and at most 2048 views)");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    cudaStream_t stream = (cudaStream_t)stream_;
    SparseGeom g;
    g.views = views;
    g.plane_stride = plane_stride;
    g.V = n_views;
    g.rows = rows;
    g.W = W;
    g.row0 = row0;
    g.occ = occ;
    g.occ_words = occ_words;
    g.tiles_x = (W + VS_TILE_W - 1) / VS_TILE_W;
    g.ty_first = row0 / VS_TILE_H;
    const int ty_last = (row0 + rows - 1) / VS_TILE_H;
    g.n_units = (ty_last - g.ty_first + 1) * g.tiles_x;
    g.out = out_mean;
    make_sum_tree(n_views, g);
    // scratch: per-bin counters + per-bin unit lists (grown outside of stream capture: the first call sizes it)
    const size_t need = (size_t)kSparseBins * (1 + (size_t)g.n_units);
    if (ctx->fuse_plan_ints < need) {
        if (ctx->d_fuse_plan) cudaFree(ctx->d_fuse_plan);
        ctx->d_fuse_plan = nullptr;
        ctx->fuse_plan_ints = 0;
        VS_CUDA(cudaMalloc(&ctx->d_fuse_plan, need * sizeof(int)));
        ctx->fuse_plan_ints = need;
    }
    int* bin_count = ctx->d_fuse_plan;
    int* bin_list = ctx->d_fuse_plan + kSparseBins;
    g.bin_count = bin_count;
    g.bin_list = bin_list;
    VS_CUDA(cudaMemsetAsync(bin_count, 0, kSparseBins * sizeof(int), stream));
    k_fuse_plan<<<(g.n_units + 255) / 256, 256, 0, stream>>>(occ, occ_words, g.tiles_x, g.ty_first, g.n_units, bin_count,
                                                             bin_list);
    VS_CHECK_LAUNCH(ctx, "k_fuse_plan");
    // one launch per bin that this view count can reach (a tile never has more than n_views views), on the caller's stream
    int rc = VS_OK;
    for (int b = 0; b < kSparseBins && rc == VS_OK; ++b) {
        if (b > 0 && h_bin_cap[b - 1] >= n_views) break;
        switch (b) {
            case 0: rc = launch_sparse_regs<8>(ctx, g, b, stream); break;
            case 1: rc = launch_sparse_regs<16>(ctx, g, b, stream); break;
            case 2: rc = launch_sparse_regs<24>(ctx, g, b, stream); break;
            case 3: rc = launch_sparse_regs<32>(ctx, g, b, stream); break;
            case 4: rc = launch_sparse_regs<40>(ctx, g, b, stream); break;
            case 5: rc = launch_sparse_regs<48>(ctx, g, b, stream); break;
            case 6: rc = launch_sparse_regs<56>(ctx, g, b, stream); break;
            case 7: rc = launch_sparse_regs<64>(ctx, g, b, stream); break;
            case 8: rc = launch_sparse_regs<80>(ctx, g, b, stream); break;
            case 9: rc = launch_sparse_regs<96>(ctx, g, b, stream); break;
            case 10: rc = launch_sparse_regs<112>(ctx, g, b, stream); break;
            case 11: rc = launch_sparse_pair<64>(ctx, g, b, stream); break;     // 113..128: two threads x 64 (255 registers otherwise)
            case 12: rc = launch_sparse_pair<80>(ctx, g, b, stream); break;
            case 13: rc = launch_sparse_pair<104>(ctx, g, b, stream); break;
            case 14: rc = launch_sparse_large<4, 64>(ctx, g, b, stream); break;
            case 15: rc = launch_sparse_large<8, 52>(ctx, g, b, stream); break;
            case 16: rc = launch_sparse_large<8, 64>(ctx, g, b, stream); break;
            case 17: rc = launch_sparse_large<32, 32>(ctx, g, b, stream); break;
            case 18: rc = launch_sparse_large<32, 64>(ctx, g, b, stream); break;
        }
    }
    return rc;
}
