// Stage B of one tile as a device function (see finalize.cu for the algorithm notes).  Internal header.
#pragma once

#include "finalize_common.cuh"

namespace vsfin {

// all-empty tile: every output is NaN (hole fill and blur of nothing)
template <typename T, typename Sink>
__device__ __forceinline__ void write_nan_tile(float* __restrict__ blur_out, T* __restrict__ filled_out, int ty0, int tx0,
                                               int H, int W, unsigned long long* __restrict__ nan_count,
                                               const Sink& sink) {
    unsigned n = 0;
    for (int i = threadIdx.x; i < TW * TH; i += kThreads) {
        const int r = i / TW, c = i - r * TW;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy < H && gx < W) {
            if (blur_out != nullptr) {
                blur_out[(size_t)gy * W + gx] = CUDART_NAN_F;
                if (!sink.skips_empty_tiles()) sink.store1(gy, gx, W, CUDART_NAN_F);
            }
            if (filled_out != nullptr) filled_out[(size_t)gy * W + gx] = (T)CUDART_NAN;
            ++n;
        }
    }
    if (blur_out != nullptr) block_count_flush(n, nan_count);
}

// Stage B of ONE 64x32 output tile (tile column bx, tile row by), executed by the 256 threads of a CTA: the body of
// k_grid_finalize (finalize.cu) and of the stage-B role of k_stage_ab (stage_ab.cu).  Every thread of the CTA must call
// it (block barriers inside); a caller that loops over tiles puts a __syncthreads() between two calls.
template <typename Key, typename Sink>
__device__ __forceinline__ void grid_finalize_tile(const int bx, const int by, const Key* __restrict__ keygrid, int W, int H,
                                                   typename KeyTraits<Key>::value_t* __restrict__ filled_out,
                                                   float* __restrict__ blur_out, int simd_cols,
                                                   unsigned long long* __restrict__ nan_count, const Sink& sink) {
    typedef typename KeyTraits<Key>::value_t T;
    constexpr bool kSameTile = sizeof(T) == sizeof(float);   // float32 keys: one tile serves fill and blur
    __shared__ __align__(16) T s_raw[TR * TS];              // decoded keys; holes are patched in place after phase 2
    __shared__ __align__(16) float s_fill32[kSameTile ? 4 : TR * TS];  // float32 copy for the f64 path
    __shared__ unsigned short s_hole_pos[MAX_HOLES];
    __shared__ float s_hole_val[kSameTile ? MAX_HOLES : 1];
    __shared__ int s_has_nan;
    float* s_fill = kSameTile ? reinterpret_cast<float*>(s_raw) : s_fill32;
    const int tid = threadIdx.x;
    const int tx0 = bx * TW, ty0 = by * TH;
    if (tid == 0) s_has_nan = 0;
    sink.prepare(ty0, H, W);                                 // PeerSink: per-row destination table of this tile
    __syncthreads();

    // 1. decode keys (+2 halo).  Key 0 (= empty, also used outside the grid) decodes to NaN, and the fill only
    //    uses in-range neighbours (:73).  Holes = empty cells inside the grid within the 1-cell halo; they are
    //    appended to a list with one shared atomic per warp-row (ballot-aggregated).
    //    Warp w loads tile rows w, w+8, ...; lane l takes columns l, l+32 and (l < 4) l+64: no div/mod, and the
    //    row address is computed once per row.
    const int warp = tid >> 5, lane = tid & 31;
    int wcnt = 0;   // holes listed by this warp (warp-uniform)
    // Interior tiles of a grid whose rows are 16-byte multiples: the 36 x 72 box (grid columns tx0-4 .. tx0+67) is
    // fetched with 16-byte loads, 18 per row.  Edge tiles (and odd pitches) take the scalar, bounds-checked loader.
    const bool vec_ok = sizeof(Key) == 4 && (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(keygrid) & 15) == 0) &&
                        tx0 >= 4 && tx0 + TW + 4 <= W && ty0 >= 2 && ty0 + TH + 2 <= H;
    if (vec_ok) {
        constexpr int VPR = TS / 4;                 // 18 vectors per row
        constexpr int NV4 = (TR * VPR + kThreads - 1) / kThreads;   // 3 per thread
        uint4 kv[NV4];
#pragma unroll
        for (int q = 0; q < NV4; ++q) {             // all loads first
            const int i = tid + q * kThreads;
            const int r = i / VPR, v4 = i - r * VPR;
            kv[q] = make_uint4(0u, 0u, 0u, 0u);
            if (i < TR * VPR)
                kv[q] = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(keygrid) +
                                                       (size_t)(ty0 - 2 + r) * W + (tx0 - 4) + 4 * v4);
        }
        {   // a tile whose whole box is empty (large AOIs: most tiles of most views) is all-NaN: skip the work
            unsigned nz = 0;
#pragma unroll
            for (int q = 0; q < NV4; ++q) nz |= kv[q].x | kv[q].y | kv[q].z | kv[q].w;
            if (!__syncthreads_or(nz != 0)) {
                write_nan_tile(blur_out, filled_out, ty0, tx0, H, W, nan_count, sink);
                return;
            }
        }
#pragma unroll
        for (int q = 0; q < NV4; ++q) {
            const int i = tid + q * kThreads;
            const int r = i / VPR, v4 = i - r * VPR;
            const bool act = i < TR * VPR;
            const uint32_t kk[4] = {kv[q].x, kv[q].y, kv[q].z, kv[q].w};
            const int pos0 = r * TS + 4 * v4;       // box column 4*v4 <-> tile column 4*v4 - OFF
            if (act) {
                float4 f = make_float4(vs_unkey32(kk[0]), vs_unkey32(kk[1]), vs_unkey32(kk[2]), vs_unkey32(kk[3]));
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(s_raw) + pos0) = f;
            }
            const bool row_in = act && (unsigned)(r - 1) < (unsigned)(TH + 2);
            bool h4[4];
            bool any_h = false;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * v4 + e - OFF;     // tile column
                h4[e] = row_in && kk[e] == 0 && (unsigned)(c - 1) < (unsigned)(TW + 2);
                any_h |= h4[e];
            }
            if (__any_sync(0xffffffffu, any_h)) {   // warp-uniform
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned m = __ballot_sync(0xffffffffu, h4[e]);
                    if (h4[e]) s_hole_pos[warp * HSEG2 + wcnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(pos0 + e);
                    wcnt += __popc(m);
                }
            }
        }
    } else {
        constexpr int NIT = (TR + 7) / 8;   // rows per warp
        // 1a. issue every global load of this thread before touching the results (memory-level parallelism: the
        //     phase is latency-bound otherwise)
        Key keys[NIT][3];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int r = warp + 8 * it;
            const int gy = ty0 - 2 + r;
            const bool row_ok = r < TR && (unsigned)gy < (unsigned)H;
            const Key* __restrict__ row = keygrid + (size_t)(row_ok ? gy : 0) * W + (tx0 - 2);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int c = lane + 32 * part;
                const int gx = tx0 - 2 + c;
                Key key = 0;
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
                write_nan_tile(blur_out, filled_out, ty0, tx0, H, W, nan_count, sink);
                return;
            }
        }
        // 1b. decode, store, list holes.  Every warp keeps its own hole list (count in a register, no atomics).
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
                        const Key key = keys[it][part];
                        const T v = KeyTraits<Key>::decode(key);
                        s_raw[pos] = v;
                        if (!kSameTile) s_fill[pos] = (float)v;     // produce_dsm.py:58 astype(np.float32)
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
    if (tid == 0 && blur_out != nullptr) sink.mark_tile(bx, by, H);   // this (tile, view) holds data
    __syncthreads();

    // 2. hole fill: NaN cell <- median of its non-NaN 3x3 neighbours in the PRE-fill grid, dense over the list.
    //    The fill must not cascade (lib/proj_to_grid.py:65 reads a copy): with one shared tile the results are
    //    staged and patched in after a barrier; the float64 path reads s_raw and writes the separate float32 tile.
    {
        bool nan_left = false;
        for (int i = lane; i < wcnt; i += 32) {
            const int pos = s_hole_pos[warp * HSEG2 + i];
            const T* c = s_raw + pos;
            T nb[8] = {c[-TS - 1], c[-TS], c[-TS + 1], c[-1], c[1], c[TS - 1], c[TS], c[TS + 1]};
            const T v = vs_median_of_valid8<T>(nb);
            nan_left |= (v != v);
            if (kSameTile) {
                s_hole_val[warp * HSEG2 + i] = (float)v;
            } else {
                s_fill[pos] = (float)v;
                if (filled_out != nullptr) {
                    const int r = pos / TS, cc = pos - r * TS - OFF;
                    const int gy = ty0 - 2 + r, gx = tx0 - 2 + cc;
                    if (r >= 2 && r < TH + 2 && cc >= 2 && cc < TW + 2) filled_out[(size_t)gy * W + gx] = v;
                }
            }
        }
        if (nan_left) s_has_nan = 1;
        __syncthreads();
        if (kSameTile) {
            for (int i = lane; i < wcnt; i += 32) s_fill[s_hole_pos[warp * HSEG2 + i]] = s_hole_val[warp * HSEG2 + i];
            __syncthreads();
        }
    }
    if (filled_out != nullptr) {
        for (int i = tid; i < TW * TH; i += kThreads) {
            const int r = i / TW, c = i - r * TW;
            const int gy = ty0 + r, gx = tx0 + c;
            const T v = s_raw[(r + 2) * TS + OFF + c + 2];
            if (gy < H && gx < W && (kSameTile || v == v)) filled_out[(size_t)gy * W + gx] = v;
        }
    }
    if (blur_out == nullptr) return;

    // 3. cv2.medianBlur(., 3) with replicated borders
    replicate_border(s_fill, ty0, tx0, H, W);
    __syncthreads();
    const unsigned n_nan = blur_tile(s_fill, ty0, tx0, H, W, H, s_has_nan != 0, simd_cols != 0, blur_out, 0, sink);
    block_count_flush(n_nan, nan_count);
}

}  // namespace vsfin
