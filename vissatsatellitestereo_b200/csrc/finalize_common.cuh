// Shared device code of the stage-B kernels (finalize.cu: plain loads; finalize_tma.cu: TMA-fed persistent pipeline).
#pragma once

#include <math_constants.h>

#include "median.cuh"
#include "vs_common.cuh"

namespace vsfin {

constexpr int TW = 64;            // tile width  (outputs)
constexpr int TH = 32;            // tile height (outputs)
constexpr int kThreads = 256;
constexpr int OFF = 2;            // column offset of the tile inside a shared row.  With OFF = 2 the row starts at grid
                                  // column tx0 - 4: a multiple of 4 elements, which TMA requires of the innermost
                                  // box coordinate (16-byte granularity; verified on B200: x = 61 faults, x = 60 and
                                  // negative multiples of 4 are fine), and the middle 4 floats of every 6-float blur
                                  // window are 16-byte aligned.
constexpr int TS = 72;            // shared row stride in elements (>= OFF + TW + 4, multiple of 4)
constexpr int TR = TH + 4;        // shared rows (halo 2)
constexpr int TC = TW + 4;        // shared columns in use (halo 2)
constexpr int HSEG = ((TR + 7) / 8) * (TW + 2) + 6;   // hole-list capacity per warp (its rows x inner columns), 336
constexpr int HSEG2 = 392;                             // capacity per warp in finalize.cu (its vector path lists up to 3*32*4 cells per warp)
constexpr int MAX_HOLES = 8 * HSEG2;

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
    const int tid = threadIdx.x;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    unsigned r = __reduce_add_sync(0xffffffffu, local);
    if ((tid & 31) == 0 && r) atomicAdd(&s_cnt, r);
    __syncthreads();
    if (tid == 0 && s_cnt) atomicAdd(counter, (unsigned long long)s_cnt);
}

// ---- 3x3 median pieces ----------------------------------------------------------------------------------------
struct Col3 {
    float lo, mid, hi;
};
__device__ __forceinline__ Col3 sort_col3(float a, float b, float c) {
    Col3 r;
    const float ab_lo = fminf(a, b), ab_hi = fmaxf(a, b);
    r.lo = fminf(ab_lo, c);
    r.hi = fmaxf(ab_hi, c);
    r.mid = fmaxf(ab_lo, fminf(ab_hi, c));
    return r;
}
__device__ __forceinline__ float med3f(float a, float b, float c) {
    return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}
// median of 9 from three sorted columns: med3(max of minima, med3 of medians, min of maxima)
__device__ __forceinline__ float median_of_cols(const Col3& a, const Col3& b, const Col3& c) {
    return med3f(fmaxf(fmaxf(a.lo, b.lo), c.lo), med3f(a.mid, b.mid, c.mid), fminf(fminf(a.hi, b.hi), c.hi));
}
__device__ __forceinline__ bool has_nan9(const float (&r0)[3], const float (&r1)[3], const float (&r2)[3]) {
    const float sum = ((r0[0] + r0[1]) + (r0[2] + r1[0])) + ((r1[1] + r1[2]) + (r2[0] + r2[1])) + r2[2];
    return !(sum == sum);   // NaN (or +inf with -inf): take the exact path
}
__device__ __forceinline__ float median9_exact(const float* r0, const float* r1, const float* r2, bool simd) {
    return simd ? vs_median9_net<true>(r0[0], r0[1], r0[2], r1[0], r1[1], r1[2], r2[0], r2[1], r2[2])
                : vs_median9_net<false>(r0[0], r0[1], r0[2], r1[0], r1[1], r1[2], r2[0], r2[1], r2[2]);
}

// ---- output sinks ----------------------------------------------------------------------------------------------
// Besides the caller's array, stage B can store every output row into the row-band stacks of the ranks that fuse it
// (multi-GPU exchange through peer-mapped memory).  NoSink compiles to nothing.
struct NoSink {
    __device__ __forceinline__ void prepare(int, int, int) const {}         // every thread, once per tile, before a barrier
    __device__ __forceinline__ void store4(int, int, int, const float4&) const {}
    __device__ __forceinline__ void store1(int, int, int, float) const {}
    __device__ __forceinline__ void mark_tile(int, int, int) const {}       // one thread per CTA calls it
    __device__ __forceinline__ bool skips_empty_tiles() const { return false; }
    __device__ __forceinline__ bool skips_empty_tiles_or_none() const { return true; }   // nothing goes to peers
};
// Records which (tile, view) pairs hold data (vs_set_occupancy); no peer stores.
struct OccSink {
    VsOccPlan o;
    __device__ __forceinline__ void prepare(int, int, int) const {}
    __device__ __forceinline__ void store4(int, int, int, const float4&) const {}
    __device__ __forceinline__ void store1(int, int, int, float) const {}
    __device__ __forceinline__ void mark_tile(int tx, int ty, int) const {
        atomicOr(o.occ + ((size_t)ty * o.tiles_x + tx) * o.occ_words + (o.view >> 5), 1u << (o.view & 31));
    }
    __device__ __forceinline__ bool skips_empty_tiles() const { return false; }
    __device__ __forceinline__ bool skips_empty_tiles_or_none() const { return true; }
};
struct PeerSink {
    VsPeerPlan p;
    // sparse exchange: an all-empty tile is not stored into the band stacks; a tile with data sets its (tile, view)
    // bit in the bitmap of every rank whose band (+ halo) it touches (system-scope atomics over NVLink)
    __device__ __forceinline__ bool skips_empty_tiles() const { return p.occ_words > 0; }
    __device__ __forceinline__ bool skips_empty_tiles_or_none() const { return p.occ_words > 0; }
    __device__ __forceinline__ void mark_tile(int tx, int ty, int H) const {
        if (p.occ_words <= 0) return;
        const int y0 = ty * TH, y1 = min(y0 + TH, H) - 1;
        const int j0 = max(band_of(y0) - 1, 0), j1 = min(band_of(y1) + 1, p.n - 1);
        for (int j = j0; j <= j1; ++j) {
            if (p.row0[j + 1] == p.row0[j] || p.occ[j] == nullptr) continue;
            if (y1 >= p.row0[j] - p.halo && y0 < p.row0[j + 1] + p.halo)
                atomicOr_system(p.occ[j] + ((size_t)ty * p.tiles_x + tx) * p.occ_words + (p.view >> 5), 1u << (p.view & 31));
        }
    }
    // band j of row y: row0[j] <= y < row0[j+1].  Bands are near-uniform (numpy.array_split), so y*n/H is at most
    // one band off; the two loops make it exact (and step over empty bands when H < n).
    __device__ __forceinline__ int band_of(int y) const {
        int j = min((int)__umulhi((unsigned)y, p.inv), p.n - 1);
        while (y < p.row0[j]) --j;
        while (y >= p.row0[j + 1]) ++j;
        return j;
    }
    __device__ __forceinline__ float* at(int j, int y, int x, int W) const {
        const int h0 = max(p.row0[j] - p.halo, 0);
        return p.plane[j] + (size_t)(y - h0) * W + x;
    }
    template <typename F>
    __device__ __forceinline__ void each_dest(int y, F f) const {
        const int j = band_of(y);
        f(j);
        // the first `halo` rows of band j are the lower halo of band j-1, the last ones the upper halo of band j+1
        if (j > 0 && y - p.row0[j] < p.halo && p.row0[j] > p.row0[j - 1]) f(j - 1);
        if (j + 1 < p.n && p.row0[j + 1] - y <= p.halo && p.row0[j + 2] > p.row0[j + 1]) f(j + 1);
    }
    // The 16-byte stores of a tile go through a per-tile table in shared memory: for each of its 32 rows the row starts
    // in the (up to three) band stacks that receive it, worked out once per tile by 32 threads (prepare) instead of by
    // every thread for every segment (band search + halo tests: ~40 instructions per store, +20 % on the kernel).
    static __device__ __forceinline__ float** row_table() {
        __shared__ float* t[TH * 3];
        return t;
    }
    static __device__ __forceinline__ int* row_info() {   // [r] = destinations of tile row r, [TH] = first grid row of the tile
        __shared__ int t[TH + 1];
        return t;
    }
    __device__ __forceinline__ void prepare(int ty0, int H, int W) const {
        const int r = threadIdx.x;
        if (r < TH) {
            int n = 0;
            const int y = ty0 + r;
            if (y < H) each_dest(y, [&](int j) { row_table()[3 * r + n++] = at(j, y, 0, W); });
            row_info()[r] = n;
        }
        if (r == 0) row_info()[TH] = ty0;
    }
    __device__ __forceinline__ void store4(int y, int x, int W, const float4& v) const {
        const int r = y - row_info()[TH];
        const int nd = row_info()[r];
        float* const* t = row_table() + 3 * r;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < nd) {
                float* d = t[k] + x;
                if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
                    *reinterpret_cast<float4*>(d) = v;
                } else {   // odd row pitch
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
            }
        }
    }
    __device__ __forceinline__ void store1(int y, int x, int W, float v) const {
        each_dest(y, [&](int j) { *at(j, y, x, W) = v; });
    }
};

// Blur phase shared by K2 and K4.  s_fill: TR x TS tile; element (r, OFF + c) is grid cell (ty0 - 2 + r, tx0 - 2 + c).
// Rows/columns within the 1-cell halo must be final (hole-filled) and, outside the grid, replicated from the
// nearest inside cell.  Writes out[(gy - out_row0) * W + gx] for gy in [ty0, min(ty0 + TH, row_limit)).
template <typename Sink = NoSink>
__device__ __forceinline__ unsigned blur_tile(const float* __restrict__ s_fill, int ty0, int tx0, int H, int W,
                                              int row_limit, bool tile_has_nan, bool simd_cols,
                                              float* __restrict__ out, int out_row0, const Sink& sink = Sink()) {
    const int tid = threadIdx.x;
    unsigned n_nan = 0;
    if (H == 1 || W == 1) {  // OpenCV's 1-D special case (block-uniform): 3-tap median along the line
        for (int i = tid; i < TW * TH; i += kThreads) {
            const int r = i / TW, c = i - r * TW;
            const int gy = ty0 + r, gx = tx0 + c;
            if (gy < row_limit && gx < W) {
                const float* p = s_fill + (r + 2) * TS + (OFF + c + 2);
                const float m = (H == 1) ? vs_median3_line(p[-1], p[0], p[1]) : vs_median3_line(p[-TS], p[0], p[TS]);
                out[(size_t)(gy - out_row0) * W + gx] = m;
                sink.store1(gy, gx, W, m);
                n_nan += (m != m);
            }
        }
        return n_nan;
    }
    // patch of 4 columns x 2 rows per thread
    const int k = tid & 15, rp = tid >> 4;          // strip 0..15, row pair 0..15
    const int r = 2 * rp, c = 4 * k;
    const int gy = ty0 + r, gx = tx0 + c;
    if (gy >= row_limit || gx >= W) return 0;
    // rows r-1 .. r+2 of the tile, columns c-1 .. c+4  ->  shared rows r+1 .. r+4, columns OFF+c+1 .. OFF+c+6
    float win[4][6];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float* p = s_fill + (r + 1 + j) * TS + (OFF + c + 1);   // p + 1 is 16-byte aligned
        const float4 a = *reinterpret_cast<const float4*>(p + 1);
        win[j][0] = p[0]; win[j][1] = a.x; win[j][2] = a.y; win[j][3] = a.z; win[j][4] = a.w; win[j][5] = p[5];
    }
    float res[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        Col3 col[6];
#pragma unroll
        for (int x = 0; x < 6; ++x) col[x] = sort_col3(win[j][x], win[j + 1][x], win[j + 2][x]);
#pragma unroll
        for (int x = 0; x < 4; ++x) res[j][x] = median_of_cols(col[x], col[x + 1], col[x + 2]);
    }
    if (tile_has_nan) {  // block-uniform; rare
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const float r0[3] = {win[j][x], win[j][x + 1], win[j][x + 2]};
                const float r1[3] = {win[j + 1][x], win[j + 1][x + 1], win[j + 1][x + 2]};
                const float r2[3] = {win[j + 2][x], win[j + 2][x + 1], win[j + 2][x + 2]};
                if (has_nan9(r0, r1, r2)) {
                    const int x_g = gx + x;
                    res[j][x] = median9_exact(r0, r1, r2, simd_cols && x_g >= 1 && x_g <= W - 2);
                }
            }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int y_g = gy + j;
        if (y_g < row_limit) {
            float* o = out + (size_t)(y_g - out_row0) * W + gx;
            if (gx + 3 < W && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                const float4 v4 = make_float4(res[j][0], res[j][1], res[j][2], res[j][3]);
                *reinterpret_cast<float4*>(o) = v4;
                sink.store4(y_g, gx, W, v4);
                if (tile_has_nan) {   // block-uniform; a tile without a NaN going in has none coming out
#pragma unroll
                    for (int x = 0; x < 4; ++x) n_nan += (res[j][x] != res[j][x]);
                }
            } else {
#pragma unroll
                for (int x = 0; x < 4; ++x)
                    if (gx + x < W) {
                        o[x] = res[j][x];
                        sink.store1(y_g, gx + x, W, res[j][x]);
                        n_nan += (tile_has_nan && res[j][x] != res[j][x]);
                    }
            }
        }
    }
    return n_nan;
}

// Replicate the grid border into the part of the tile's 1-cell halo that lies outside the grid
// (cv2 BORDER_REPLICATE).  Only tiles touching the grid border have such cells.
__device__ __forceinline__ void replicate_border(float* __restrict__ s_fill, int ty0, int tx0, int H, int W) {
    if (ty0 > 0 && tx0 > 0 && ty0 + TH < H && tx0 + TW < W) return;  // interior tile (block-uniform)
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += kThreads) {
        const int r = 1 + i / (TW + 2), c = 1 + i % (TW + 2);
        const int gy = ty0 - 2 + r, gx = tx0 - 2 + c;
        if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
            const int cy = min(max(gy, 0), H - 1), cx = min(max(gx, 0), W - 1);
            const int rr = cy - (ty0 - 2), cc = cx - (tx0 - 2);
            if (rr >= 1 && rr <= TH + 2 && cc >= 1 && cc <= TW + 2) s_fill[r * TS + OFF + c] = s_fill[rr * TS + OFF + cc];
        }
    }
}


inline int simd_cols_for(int W, int simd_lanes) { return (simd_lanes > 0 && W >= simd_lanes + 2) ? 1 : 0; }

}  // namespace vsfin
