// Stage B, float32-key path (K2): key grid -> per-view DSM, computed in KEY space.
//
//   lib/proj_to_grid.py:62-79 (decode, 3x3 NaN-hole fill from the PRE-fill grid) + produce_dsm.py:58
//   (astype(float32) + cv2.medianBlur(., 3)), one CTA (256 threads) per 64x32 output tile, like finalize.cu.
//
// The kernel is bound by the ALU pipe (min/max, compares, selects), so what is optimised is the instruction count
// per cell (round 1: 136).  The order-preserving uint32 keys are never decoded on the way in: min/max/median of keys
// are the keys of the min/max/median of the floats, and integer arithmetic gives what float arithmetic cannot:
//   * the middle of three is  a + b + c - min3 - max3  EXACTLY (modular arithmetic), so a sorted 3-column costs
//     2 three-input min/max (VIMNMX3) + 2 three-input adds instead of 6 two-input float min/max, and the median of
//     three costs 4 instructions;
//   * the hole fill needs "the middle one / the two middle ones of the k valid values among 8 neighbours".  The empty
//     neighbours (key 0) are re-labelled alternately 0 (below everything) and 0xffffffff (above everything): the
//     valid values then sit in the middle of the 8, and their median is ALWAYS at sorted positions 3 and 4 (k even)
//     or 4 (k odd).  So a fixed selection network for positions 3/4 replaces "sort 8 + select by count", and the
//     parity of k is the state of the alternation -- no counter.
//   * holes are listed once per thread (12-bit mask of its 3 x 4 loaded keys, one warp scan) instead of one ballot
//     per key.
// Only the values that leave the kernel are decoded.  Windows that still contain a NaN after the fill take the
// exact float emulation of OpenCV's network (SIMD / scalar column semantics), as before.
// Bit-identical to finalize.cu's k_grid_finalize<uint32_t> (tests compare both with cv2 / the reference goldens).
#include "finalize_common.cuh"

using namespace vsfin;

namespace {

__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u32(a, b, c); }
__device__ __forceinline__ uint32_t umax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u32(a, b, c); }
// middle of three keys, exact in modular arithmetic
__device__ __forceinline__ uint32_t umed3(uint32_t a, uint32_t b, uint32_t c) {
    return a + b + c - umin3(a, b, c) - umax3(a, b, c);
}
struct KCol {
    uint32_t lo, mid, hi;
};
__device__ __forceinline__ KCol ksort3(uint32_t a, uint32_t b, uint32_t c) {
    KCol r;
    r.lo = umin3(a, b, c);
    r.hi = umax3(a, b, c);
    r.mid = a + b + c - r.lo - r.hi;
    return r;
}

// The k valid keys among 8 neighbours -> key of np.median(valid) (float32 arithmetic for the 2-value mean, see
// median.cuh: vs_mean2), 0 if there is none.  has_none is set when k == 0.
__device__ __forceinline__ uint32_t kmedian_of_valid8(uint32_t (&v)[8], bool& has_none) {
    uint32_t pad = 0u;   // label of the next empty neighbour: 0, 0xffffffff, 0, ...
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool emp = v[i] == 0u;
        v[i] = emp ? pad : v[i];
        pad = emp ? ~pad : pad;
    }
    const bool k_odd = pad != 0u;          // 8 - k labels were handed out; pad flipped once per label
    // Batcher odd-even merge sort for 8 inputs (19 exchanges); only outputs 3 and 4 are read, so the compiler drops the
    // exchanges that feed the others only
#define VS_KCE(i, j) { const uint32_t lo = min(v[i], v[j]); const uint32_t hi = max(v[i], v[j]); v[i] = lo; v[j] = hi; }
    VS_KCE(0, 1) VS_KCE(2, 3) VS_KCE(4, 5) VS_KCE(6, 7)
    VS_KCE(0, 2) VS_KCE(1, 3) VS_KCE(4, 6) VS_KCE(5, 7)
    VS_KCE(1, 2) VS_KCE(5, 6)
    VS_KCE(0, 4) VS_KCE(1, 5) VS_KCE(2, 6) VS_KCE(3, 7)
    VS_KCE(2, 4) VS_KCE(3, 5)
    VS_KCE(1, 2) VS_KCE(3, 4) VS_KCE(5, 6)
#undef VS_KCE
    // k == 0: four labels of each kind -> v[3] = 0, v[4] = 0xffffffff (never the key of a finite float)
    has_none = v[4] == 0xffffffffu;
    const uint32_t avg = vs_key32(vs_mean2(vs_unkey32(v[3]), vs_unkey32(v[4])));
    const uint32_t med = k_odd ? v[4] : avg;
    return has_none ? 0u : med;
}

// cv2 BORDER_REPLICATE on the key tile (only tiles touching the grid border have such cells)
__device__ __forceinline__ void replicate_border_keys(uint32_t* __restrict__ s_key, int ty0, int tx0, int H, int W) {
    if (ty0 > 0 && tx0 > 0 && ty0 + TH < H && tx0 + TW < W) return;  // interior tile (block-uniform)
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += kThreads) {
        const int r = 1 + i / (TW + 2), c = 1 + i % (TW + 2);
        const int gy = ty0 - 2 + r, gx = tx0 - 2 + c;
        if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
            const int cy = min(max(gy, 0), H - 1), cx = min(max(gx, 0), W - 1);
            const int rr = cy - (ty0 - 2), cc = cx - (tx0 - 2);
            if (rr >= 1 && rr <= TH + 2 && cc >= 1 && cc <= TW + 2) s_key[r * TS + OFF + c] = s_key[rr * TS + OFF + cc];
        }
    }
}

template <typename Sink>
__device__ __forceinline__ unsigned blur_tile_keys(const uint32_t* __restrict__ s_key, int ty0, int tx0, int H, int W,
                                                   bool tile_has_nan, bool simd_cols, float* __restrict__ out,
                                                   const Sink& sink) {
    const int tid = threadIdx.x;
    unsigned n_nan = 0;
    if (H == 1 || W == 1) {  // OpenCV's 1-D special case (block-uniform): 3-tap median along the line
        for (int i = tid; i < TW * TH; i += kThreads) {
            const int r = i / TW, c = i - r * TW;
            const int gy = ty0 + r, gx = tx0 + c;
            if (gy < H && gx < W) {
                const uint32_t* p = s_key + (r + 2) * TS + (OFF + c + 2);
                const float m = (H == 1) ? vs_median3_line(vs_unkey32(p[-1]), vs_unkey32(p[0]), vs_unkey32(p[1]))
                                         : vs_median3_line(vs_unkey32(p[-TS]), vs_unkey32(p[0]), vs_unkey32(p[TS]));
                out[(size_t)gy * W + gx] = m;
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
    if (gy >= H || gx >= W) return 0;
    // rows r-1 .. r+2 of the tile, columns c-1 .. c+4  ->  shared rows r+1 .. r+4, columns OFF+c+1 .. OFF+c+6
    uint32_t win[4][6];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t* p = s_key + (r + 1 + j) * TS + (OFF + c + 1);   // p + 1 is 16-byte aligned
        const uint4 a = *reinterpret_cast<const uint4*>(p + 1);
        win[j][0] = p[0]; win[j][1] = a.x; win[j][2] = a.y; win[j][3] = a.z; win[j][4] = a.w; win[j][5] = p[5];
    }
    float res[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        KCol col[6];
#pragma unroll
        for (int x = 0; x < 6; ++x) col[x] = ksort3(win[j][x], win[j + 1][x], win[j + 2][x]);
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const uint32_t a = umax3(col[x].lo, col[x + 1].lo, col[x + 2].lo);
            const uint32_t b = umed3(col[x].mid, col[x + 1].mid, col[x + 2].mid);
            const uint32_t cc = umin3(col[x].hi, col[x + 1].hi, col[x + 2].hi);
            res[j][x] = vs_unkey32(umed3(a, b, cc));
            if (tile_has_nan) {  // block-uniform; tiles with NaNs left after the fill
                const uint32_t wmin = umin3(col[x].lo, col[x + 1].lo, col[x + 2].lo);
                if (wmin == 0u) {   // the window holds an empty cell: OpenCV's network on the floats, exactly
                    const uint32_t wmax = umax3(col[x].hi, col[x + 1].hi, col[x + 2].hi);
                    float m = CUDART_NAN_F;   // nothing but NaNs in, NaN out
                    if (wmax != 0u) {
                        const float r0[3] = {vs_unkey32(win[j][x]), vs_unkey32(win[j][x + 1]), vs_unkey32(win[j][x + 2])};
                        const float r1[3] = {vs_unkey32(win[j + 1][x]), vs_unkey32(win[j + 1][x + 1]), vs_unkey32(win[j + 1][x + 2])};
                        const float r2[3] = {vs_unkey32(win[j + 2][x]), vs_unkey32(win[j + 2][x + 1]), vs_unkey32(win[j + 2][x + 2])};
                        const int x_g = gx + x;
                        m = median9_exact(r0, r1, r2, simd_cols && x_g >= 1 && x_g <= W - 2);
                    }
                    res[j][x] = m;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int y_g = gy + j;
        if (y_g < H) {
            float* o = out + (size_t)y_g * W + gx;
            if (gx + 3 < W && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                const float4 v4 = make_float4(res[j][0], res[j][1], res[j][2], res[j][3]);
                *reinterpret_cast<float4*>(o) = v4;
                sink.store4(y_g, gx, W, v4);
                if (tile_has_nan) {
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

template <typename Sink>
__device__ __forceinline__ void write_nan_tile_keys(float* __restrict__ blur_out, int ty0, int tx0, int H, int W,
                                                    unsigned long long* __restrict__ nan_count, const Sink& sink) {
    unsigned n = 0;
    // full tile of a 16-byte-pitch grid with nothing to forward to peers: 2 float4 stores per thread
    if (sink.skips_empty_tiles_or_none() && (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(blur_out) & 15) == 0) &&
        tx0 + TW <= W && ty0 + TH <= H) {
        const float4 nan4 = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
        for (int i = threadIdx.x; i < TW * TH / 4; i += kThreads) {
            const int r = i / (TW / 4), c4 = i - r * (TW / 4);
            *reinterpret_cast<float4*>(blur_out + (size_t)(ty0 + r) * W + tx0 + 4 * c4) = nan4;
        }
        if (nan_count != nullptr && threadIdx.x == 0) atomicAdd(nan_count, (unsigned long long)(TW * TH));
        return;
    }
    for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
        const int r = i / TW, c = i - r * TW;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy < H && gx < W) {
            blur_out[(size_t)gy * W + gx] = CUDART_NAN_F;
            if (!sink.skips_empty_tiles()) sink.store1(gy, gx, W, CUDART_NAN_F);
            ++n;
        }
    }
    block_count_flush(n, nan_count);
}

// Sparse mode: zero the keys of the tiles stage A touched and reset their marks -- replaces the whole-grid memset
// between views (on a large AOI a view touches about a third of the tiles).
__global__ void __launch_bounds__(kThreads)
k_clear_touched(uint32_t* __restrict__ keygrid, unsigned char* __restrict__ touched, int W, int H, int tiles_x, int n_tiles) {
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const bool on = touched[t] != 0;
        __syncthreads();                     // everybody has read the mark before thread 0 resets it
        if (!on) continue;                   // block-uniform
        const int ty0 = (t / tiles_x) * TH, tx0 = (t - (t / tiles_x) * tiles_x) * TW;
        if ((W & 3) == 0 && ((reinterpret_cast<uintptr_t>(keygrid) & 15) == 0) && tx0 + TW <= W && ty0 + TH <= H) {
            for (int i = threadIdx.x; i < TW * TH / 4; i += kThreads) {
                const int r = i / (TW / 4), c4 = i - r * (TW / 4);
                *reinterpret_cast<uint4*>(keygrid + (size_t)(ty0 + r) * W + tx0 + 4 * c4) = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
                const int r = i / TW, c = i - r * TW;
                if (ty0 + r < H && tx0 + c < W) keygrid[(size_t)(ty0 + r) * W + tx0 + c] = 0u;
            }
        }
        if (threadIdx.x == 0) touched[t] = 0;
    }
}

template <typename Sink>
__global__ void __launch_bounds__(kThreads, 5)
k_grid_finalize_keys(const uint32_t* __restrict__ keygrid, int W, int H, float* __restrict__ blur_out, int simd_cols,
                     unsigned long long* __restrict__ nan_count, const unsigned char* __restrict__ touched,
                     const __grid_constant__ Sink sink) {
    __shared__ __align__(16) uint32_t s_key[TR * TS];   // keys of the tile + 2-cell halo; holes are patched in place
    __shared__ unsigned short s_hole_pos[MAX_HOLES];
    __shared__ uint32_t s_hole_val[MAX_HOLES];
    __shared__ int s_has_nan;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    if (tid == 0) s_has_nan = 0;
    sink.prepare(ty0, H, W);                                 // PeerSink: per-row destination table of this tile
    __syncthreads();
    if (touched != nullptr) {   // sparse mode: a tile whose 3x3 tile neighbourhood got no key is empty without looking
        int any = 0;
        if (tid < 9) {
            const int ty = (int)blockIdx.y + tid / 3 - 1, tx = (int)blockIdx.x + tid % 3 - 1;
            if (ty >= 0 && ty < (int)gridDim.y && tx >= 0 && tx < (int)gridDim.x) any = touched[ty * gridDim.x + tx];
        }
        if (!__syncthreads_or(any)) {
            write_nan_tile_keys(blur_out, ty0, tx0, H, W, nan_count, sink);
            return;
        }
    }

    // 1. load the keys (+2 halo; key 0 = empty, also used outside the grid) and list the holes = empty cells inside
    //    the grid within the 1-cell halo of the tile.  Every warp keeps its own list (count in a register).
    const int warp = tid >> 5, lane = tid & 31;
    int wcnt = 0;   // holes listed by this warp (warp-uniform)
    const bool vec_ok = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(keygrid) & 15) == 0) && tx0 >= 4 &&
                        tx0 + TW + 4 <= W && ty0 >= 2 && ty0 + TH + 2 <= H;
    if (vec_ok) {   // interior tile, 16-byte row pitch: the 36 x 72 box (grid columns tx0-4 .. tx0+67) in 16-byte loads
        constexpr int VPR = TS / 4;                 // 18 vectors per row
        constexpr int NV4 = (TR * VPR + kThreads - 1) / kThreads;   // 3 per thread
        uint4 kv[NV4];
        int pos0[NV4];
#pragma unroll
        for (int q = 0; q < NV4; ++q) {             // all loads first
            const int i = tid + q * kThreads;
            const int r = i / VPR, v4 = i - r * VPR;
            pos0[q] = r * TS + 4 * v4;              // box column 4*v4 <-> tile column 4*v4 - OFF
            kv[q] = make_uint4(0u, 0u, 0u, 0u);
            if (i < TR * VPR)
                kv[q] = *reinterpret_cast<const uint4*>(keygrid + (size_t)(ty0 - 2 + r) * W + (tx0 - 4) + 4 * v4);
        }
        {   // a tile whose whole box is empty (large AOIs: most tiles of most views) is all-NaN: skip the work
            unsigned nz = 0;
#pragma unroll
            for (int q = 0; q < NV4; ++q) nz |= kv[q].x | kv[q].y | kv[q].z | kv[q].w;
            if (!__syncthreads_or(nz != 0)) {
                write_nan_tile_keys(blur_out, ty0, tx0, H, W, nan_count, sink);
                return;
            }
        }
        unsigned mask = 0;                          // bit 4q + e: key e of vector q is a hole to fill
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            const int i = tid + q * kThreads;
            const int r = i / VPR, v4 = i - r * VPR;
            const bool act = i < TR * VPR;
            if (act) *reinterpret_cast<uint4*>(s_key + pos0[q]) = kv[q];
            const bool row_in = act && (unsigned)(r - 1) < (unsigned)(TH + 2);
            const uint32_t kk[4] = {kv[q].x, kv[q].y, kv[q].z, kv[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * v4 + e - OFF;     // tile column
                const bool hole = row_in && kk[e] == 0u && (unsigned)(c - 1) < (unsigned)(TW + 2);
                mask |= hole ? (1u << (4 * q + e)) : 0u;
            }
        }
        // one exclusive warp scan of the per-thread hole counts, then every thread appends its own holes
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        wcnt = __shfl_sync(0xffffffffu, incl, 31);
        unsigned short* dst = s_hole_pos + warp * HSEG2 + (incl - cnt);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const int q = b >> 2;
            const int p0 = q == 0 ? pos0[0] : (q == 1 ? pos0[1] : pos0[NV4 - 1]);
            *dst++ = (unsigned short)(p0 + (b & 3));
        }
    } else {
        constexpr int NIT = (TR + 7) / 8;   // rows per warp
        uint32_t keys[NIT][3];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int r = warp + 8 * it;
            const int gy = ty0 - 2 + r;
            const bool row_ok = r < TR && (unsigned)gy < (unsigned)H;
            const uint32_t* __restrict__ row = keygrid + (size_t)(row_ok ? gy : 0) * W + (tx0 - 2);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int c = lane + 32 * part;
                const int gx = tx0 - 2 + c;
                uint32_t key = 0;
                if (row_ok && (part < 2 || lane < TC - 64) && (unsigned)gx < (unsigned)W) key = row[c];
                keys[it][part] = key;
            }
        }
        {
            bool nz = false;
#pragma unroll
            for (int it = 0; it < NIT; ++it)
#pragma unroll
                for (int part = 0; part < 3; ++part) nz |= (keys[it][part] != 0);
            if (!__syncthreads_or(nz)) {
                write_nan_tile_keys(blur_out, ty0, tx0, H, W, nan_count, sink);
                return;
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int r = warp + 8 * it;
            if (r < TR) {   // warp-uniform
                const int gy = ty0 - 2 + r;
                const bool row_in = (unsigned)gy < (unsigned)H && (unsigned)(r - 1) < (unsigned)(TH + 2);
#pragma unroll
                for (int part = 0; part < 3; ++part) {
                    const int c = lane + 32 * part;
                    const int pos = r * TS + OFF + c;
                    bool hole = false;
                    if (part < 2 || lane < TC - 64) {
                        const uint32_t key = keys[it][part];
                        s_key[pos] = key;
                        const int gx = tx0 - 2 + c;
                        hole = key == 0 && row_in && (unsigned)gx < (unsigned)W && (unsigned)(c - 1) < (unsigned)(TW + 2);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, hole);
                    if (m) {   // warp-uniform
                        if (hole) s_hole_pos[warp * HSEG2 + wcnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)pos;
                        wcnt += __popc(m);
                    }
                }
            }
        }
    }
    if (tid == 0) sink.mark_tile(blockIdx.x, blockIdx.y, H);   // this (tile, view) holds data
    __syncthreads();

    // 2. hole fill: empty cell <- median of its non-empty 3x3 neighbours in the PRE-fill grid (lib/proj_to_grid.py:65
    //    reads a copy), dense over the warp's list; results are staged and patched in after a barrier.
    {
        bool nan_left = false;
        for (int i = lane; i < wcnt; i += 32) {
            const uint32_t* c = s_key + s_hole_pos[warp * HSEG2 + i];
            uint32_t nb[8] = {c[-TS - 1], c[-TS], c[-TS + 1], c[-1], c[1], c[TS - 1], c[TS], c[TS + 1]};
            bool none;
            s_hole_val[warp * HSEG2 + i] = kmedian_of_valid8(nb, none);
            nan_left |= none;
        }
        if (nan_left) s_has_nan = 1;
        __syncthreads();
        for (int i = lane; i < wcnt; i += 32) s_key[s_hole_pos[warp * HSEG2 + i]] = s_hole_val[warp * HSEG2 + i];
        __syncthreads();
    }

    // 3. cv2.medianBlur(., 3) with replicated borders
    replicate_border_keys(s_key, ty0, tx0, H, W);
    __syncthreads();
    const unsigned n_nan = blur_tile_keys(s_key, ty0, tx0, H, W, s_has_nan != 0, simd_cols != 0, blur_out, sink);
    block_count_flush(n_nan, nan_count);
}

}  // namespace

// Launchers used by finalize.cu's entry points (one per sink).
int vs_launch_clear_touched(vs_ctx* ctx, uint32_t* keygrid, unsigned char* touched, int xsize, int ysize, cudaStream_t stream) {
    const int tiles_x = (xsize + TW - 1) / TW, n_tiles = tiles_x * ((ysize + TH - 1) / TH);
    const int blocks = n_tiles < ctx->sm_count * 8 ? n_tiles : ctx->sm_count * 8;
    k_clear_touched<<<blocks, kThreads, 0, stream>>>(keygrid, touched, xsize, ysize, tiles_x, n_tiles);
    VS_CHECK_LAUNCH(ctx, "k_clear_touched");
    return VS_OK;
}
int vs_launch_finalize_keys(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_cols,
                            unsigned long long* nan_count, cudaStream_t stream) {
    dim3 grid((xsize + TW - 1) / TW, (ysize + TH - 1) / TH);
    k_grid_finalize_keys<NoSink><<<grid, kThreads, 0, stream>>>(keygrid, xsize, ysize, dsm_out, simd_cols, nan_count, nullptr, NoSink());
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize_keys");
    return VS_OK;
}
int vs_launch_finalize_keys_occ(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_cols,
                                unsigned long long* nan_count, const VsOccPlan& plan, const unsigned char* touched,
                                cudaStream_t stream) {
    OccSink sink;
    sink.o = plan;
    dim3 grid((xsize + TW - 1) / TW, (ysize + TH - 1) / TH);
    k_grid_finalize_keys<OccSink><<<grid, kThreads, 0, stream>>>(keygrid, xsize, ysize, dsm_out, simd_cols, nan_count, touched, sink);
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize_keys<occ>");
    return VS_OK;
}
int vs_launch_finalize_keys_peer(vs_ctx* ctx, const uint32_t* keygrid, int xsize, int ysize, float* dsm_out, int simd_cols,
                                 unsigned long long* nan_count, const VsPeerPlan& plan, const unsigned char* touched,
                                 cudaStream_t stream) {
    PeerSink sink;
    sink.p = plan;
    dim3 grid((xsize + TW - 1) / TW, (ysize + TH - 1) / TH);
    k_grid_finalize_keys<PeerSink><<<grid, kThreads, 0, stream>>>(keygrid, xsize, ysize, dsm_out, simd_cols, nan_count, touched, sink);
    VS_CHECK_LAUNCH(ctx, "k_grid_finalize_keys<peer>");
    return VS_OK;
}
