// Device code shared by the stage-A kernels (rasterize.cu: one kernel per view; stage_ab.cu: stage A of one view
// co-scheduled with stage B of the previous one).  Internal header.
#pragma once

#include <math_constants.h>

#include "geo_chain.cuh"
#include "poly.cuh"
#include "vs_common.cuh"

namespace vsras {

constexpr int kThreads = 256;

struct RasterParams {
    double Mn[3][4];  // rows 0..2 of inv_proj_mat composed with the box normalisation: (M_i - c_i M_3) / h_i
    double M3[4];     // row 3 (homogeneous w)
    double half_z, center_z;  // to recover ENU-up from the normalised third coordinate
    double eps;               // ambiguity threshold (cells)
    int H, W;
    int xsize, ysize;
    int degree;
    int warp_agg;                // VISSAT_K1_WARPAGG=1: warp-aggregated scatter (A/B variant, see the scatter below)
    int tiles_x;                 // sparse mode: tile columns of the touched map
    unsigned char* touched;      // sparse mode (large AOIs): one byte per 64x32 tile, set when a key lands in the tile, so
                                 // that stage B and the key-grid clear visit only those tiles (nullptr: off)
    unsigned long long w_magic;  // ceil(2^48 / W)
};

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async16_s(unsigned smem_addr, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gmem));
}
__device__ __forceinline__ float4 ld_shared_f4(unsigned smem_addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_addr));
    return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// block-level reduction of the per-thread counters, one atomic per counter per block
__device__ __forceinline__ void flush_stats(unsigned long long* stats, unsigned v0, unsigned v1, unsigned v2, unsigned v3) {
    if (stats == nullptr) return;
    __shared__ unsigned s_acc[4];
    if (threadIdx.x < 4) s_acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned vals[4] = {v0, v1, v2, v3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned r = __reduce_add_sync(0xffffffffu, vals[i]);
        if ((threadIdx.x & 31) == 0 && r) atomicAdd(&s_acc[i], r);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_acc[threadIdx.x]) atomicAdd(&stats[threadIdx.x], (unsigned long long)s_acc[threadIdx.x]);
}

// Sparse mode: remember that a key landed in this tile: an unconditional byte store of a constant (fire and forget).
// Testing the mark first was measured and is much slower: the load puts an L1/L2 round trip (174 / 194 us per C3 view
// against 126 without marking) in front of every chunk of a kernel that is latency-bound already.
__device__ __forceinline__ void mark_touched(unsigned char* __restrict__ touched, int tiles_x, int ri, int ci) {
    touched[(ri >> 5) * tiles_x + (ci >> 6)] = 1;
}

// Rare per-point slow path of K1 (point outside the fitted altitude range): exact chain, kept out of line and
// fed from one device-memory parameter block so that it costs the streaming path neither registers nor a
// stack copy of the kernel parameters.  Returns 1 if the point landed in the grid, 2 if it is also within
// eps of a cell edge.
static __device__ __noinline__ int scatter_exact_point(const VsExactParams* __restrict__ ex, double u, double v, double w,
                                                uint32_t* __restrict__ keygrid, unsigned char* __restrict__ touched,
                                                int tiles_x) {
    double E, N, A;
    const VsGeoParams& g = ex->g;
    vs_enu_to_utm_exact(ex->c, g, fma(u, ex->half[0], ex->center[0]), fma(v, ex->half[1], ex->center[1]),
                        fma(w, ex->half[2], ex->center[2]), E, N, A);
    const double colf = (E - g.ul_e) / g.col_res, rowf = (g.ul_n - N) / g.row_res;
    const double cfl = floor(colf), rfl = floor(rowf);
    if (cfl >= 0.0 && rfl >= 0.0 && cfl < (double)g.xsize && rfl < (double)g.ysize && A == A) {
        atomicMax(keygrid + ((int)rfl * g.xsize + (int)cfl), vs_key32((float)A));
        if (touched != nullptr) mark_touched(touched, tiles_x, (int)rfl, (int)cfl);
        const double fcx = colf - cfl, frx = rowf - rfl;
        const double eps = ex->eps;
        return (fcx < eps || fcx > 1.0 - eps || frx < eps || frx > 1.0 - eps) ? 2 : 1;
    }
    return 0;
}

// Coefficients of the per-AOI polynomial as the kernel consumes them (kernel parameter, constant bank).
struct PolyCoefs {
    double c64[3][VS_MAX_TERMS];  // D64 > 0: the degree-D64 part (4 or 10 terms, degree-D64 index order); else all terms
    float c32[3][VS_MAX_TERMS];   // D64 > 0: terms of degree D64+1..D (degree-D index order, low terms zero)
};

constexpr int PX = 4;  // pixels per thread per step (one 16-byte load), evaluated in lockstep

// fast reciprocal: hardware seed (~2^-20) + two Newton steps; |error| ~ 1 ulp, no slow-path branch.  The
// divisor is the homogeneous coordinate of a finite camera ray (never denormal/zero for a valid pixel; a zero
// gives inf/NaN, which the finite test below rejects like the reference's division would).
__device__ __forceinline__ double fast_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}

// floor(v) as an int for |v| < 2^31 without the XU pipe: adding 1.5 * 2^52 with round-toward-minus-infinity leaves
// floor(v) in the low word of the sum (one FP64-pipe DADD.RM instead of an F2I.F64.FLOOR conversion; the conversions are
// the most loaded pipe of stage A).  Exactly __double2int_rd(v) on that range; NaN / out-of-range inputs give garbage,
// which the callers mask (such a point is never "in the box").  VISSAT_NO_MAGIC_FLOOR: A/B switch at compile time.
__device__ __forceinline__ int floor_to_int_fast(double v) {
#ifdef VISSAT_NO_MAGIC_FLOOR
    return __double2int_rd(v);
#else
    return __double2loint(__dadd_rd(v, 6755399441055744.0));
#endif
}

// One 4-pixel chunk that does not sit inside a single image row, or the ragged tail of the image: evaluated
// pixel by pixel through the same device functions (rare: only when W is not a multiple of 4 / at the very end).
template <int D, int D64>
__device__ __forceinline__ void scatter_chunk_generic(const RasterParams& p, const PolyCoefs& pc,
                                                   const VsExactParams* __restrict__ ex, const float* __restrict__ depth,
                                                   int64_t base, int64_t n_pix, uint32_t* __restrict__ keygrid,
                                                   float* __restrict__ height_map, bool audit, unsigned (&cnt)[4]) {
    for (int i = 0; i < PX; ++i) {
        const int64_t idx = base + i;
        if (idx >= n_pix) break;
        const float df = depth[idx];
        float hm = CUDART_NAN_F;
        if (df > 0.0f) {
            const int row = (int)(idx / p.W);
            const int col = (int)(idx - (int64_t)row * p.W);
            const double fc = (double)col, fr = (double)row, d = (double)df;
            const double rw = fast_rcp(fma(p.M3[0], fc, fma(p.M3[3], d, fma(p.M3[1], fr, p.M3[2]))));
            double u[1], v[1], w[1];
            u[0] = fma(p.Mn[0][0], fc, fma(p.Mn[0][3], d, fma(p.Mn[0][1], fr, p.Mn[0][2]))) * rw;
            v[0] = fma(p.Mn[1][0], fc, fma(p.Mn[1][3], d, fma(p.Mn[1][1], fr, p.Mn[1][2]))) * rw;
            w[0] = fma(p.Mn[2][0], fc, fma(p.Mn[2][3], d, fma(p.Mn[2][1], fr, p.Mn[2][2]))) * rw;
            float uf[1] = {(float)u[0]}, vf[1] = {(float)v[0]}, wf[1] = {(float)w[0]};
            if (fabsf(uf[0]) < CUDART_INF_F && fabsf(vf[0]) < CUDART_INF_F && fabsf(wf[0]) < CUDART_INF_F) {
                ++cnt[VS_STAT_VALID];
                hm = (float)fma(w[0], p.half_z, p.center_z);
                if (fabsf(uf[0]) <= 1.0f && fabsf(vf[0]) <= 1.0f) {
                    if (fabsf(wf[0]) <= 1.0f) {
                        double val[3][1];
#pragma unroll
                        for (int o = 0; o < 3; ++o) {
                            if (D64 > 0) {
                                vs_poly_eval_n<(D64 > 0 ? D64 : 1), 1, double>(pc.c64[o], u, v, w, val[o]);
                                float hi[1];
                                vs_poly_eval_n<D, 1, float>(pc.c32[o], uf, vf, wf, hi);
                                val[o][0] += (double)hi[0];
                            } else {
                                vs_poly_eval_n<D, 1, double>(pc.c64[o], u, v, w, val[o]);
                            }
                        }
                        const int ci = __double2int_rd(val[0][0]), ri = __double2int_rd(val[1][0]);
                        if (audit && (fabs(val[0][0] - rint(val[0][0])) < p.eps || fabs(val[1][0] - rint(val[1][0])) < p.eps)) {
                            ++cnt[VS_STAT_AMBIGUOUS];      // re-evaluated with the exact chain (see the main path)
                            const int r = scatter_exact_point(ex, u[0], v[0], w[0], keygrid, p.touched, p.tiles_x);
                            cnt[VS_STAT_INGRID] += (r != 0);
                        } else if ((unsigned)ci < (unsigned)p.xsize && (unsigned)ri < (unsigned)p.ysize) {
                            ++cnt[VS_STAT_INGRID];
                            atomicMax(keygrid + (ri * p.xsize + ci), vs_key32((float)val[2][0]));
                            if (p.touched != nullptr) mark_touched(p.touched, p.tiles_x, ri, ci);
                        }
                    } else {
                        ++cnt[VS_STAT_EXACT];
                        const int r = scatter_exact_point(ex, u[0], v[0], w[0], keygrid, p.touched, p.tiles_x);
                        cnt[VS_STAT_INGRID] += (r != 0);
                        cnt[VS_STAT_AMBIGUOUS] += (r == 2);
                    }
                }
            }
        }
        if (height_map != nullptr) height_map[idx] = hm;
    }
}


// One full 4-pixel chunk inside a single image row, throughput form (no counters, no height map, no tile marking): the
// same arithmetic, in the same order, as the streaming loop of k_unproject_scatter (rasterize.cu) -- the two are
// compared bit for bit by the tests.  Used by the co-scheduled stage A+B kernel (stage_ab.cu).
//   aggregate_2p5d_util.py:76 (depth <= 0 / NaN invalid), :86-90 (X = M.[col,row,1,depth], divide), :96-97 + produce_dsm /
//   lib/proj_to_grid.py:42-61 (ENU -> UTM -> row/col floor, bounds mask, per-cell max).
template <int D, int D64>
__device__ __forceinline__ void scatter_chunk_lean(const RasterParams& p, const PolyCoefs& pc,
                                                   const VsExactParams* __restrict__ ex, const float4 t4, unsigned row0,
                                                   unsigned col0, uint32_t* __restrict__ keygrid) {
    const float d4[PX] = {t4.x, t4.y, t4.z, t4.w};
    if (!(d4[0] > 0.0f || d4[1] > 0.0f || d4[2] > 0.0f || d4[3] > 0.0f)) return;
    double u[PX], v[PX], w[PX];
    {
        const double fr = (double)row0;
        const double fc0 = (double)col0;
        const double r0 = fma(p.Mn[0][1], fr, p.Mn[0][2]), r1 = fma(p.Mn[1][1], fr, p.Mn[1][2]);
        const double r2 = fma(p.Mn[2][1], fr, p.Mn[2][2]), r3 = fma(p.M3[1], fr, p.M3[2]);
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            const double fc = fc0 + (double)i;   // exact
            const double d = (double)d4[i];
            const double rw = fast_rcp(fma(p.M3[0], fc, fma(p.M3[3], d, r3)));
            u[i] = fma(p.Mn[0][0], fc, fma(p.Mn[0][3], d, r0)) * rw;
            v[i] = fma(p.Mn[1][0], fc, fma(p.Mn[1][3], d, r1)) * rw;
            w[i] = fma(p.Mn[2][0], fc, fma(p.Mn[2][3], d, r2)) * rw;
        }
    }
    float uf[PX], vf[PX], wf[PX];
    bool fast[PX];
    bool any_slow = false;
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        uf[i] = (float)u[i];
        vf[i] = (float)v[i];
        wf[i] = (float)w[i];
        const bool in_box = d4[i] > 0.0f && fabsf(uf[i]) <= 1.0f && fabsf(vf[i]) <= 1.0f;
        const bool in_alt = fabsf(wf[i]) <= 1.0f;
        fast[i] = in_box && in_alt;
        any_slow |= in_box && !in_alt;
    }
    int ci[PX], ri[PX];
    uint32_t key[PX];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        double val[PX];
        if (D64 > 0) {
            vs_poly_eval_n<(D64 > 0 ? D64 : 1), PX, double>(pc.c64[o], u, v, w, val);
            float hi[PX];
            vs_poly_eval_n_f32x2<D, PX>(pc.c32[o], uf, vf, wf, hi);
#pragma unroll
            for (int i = 0; i < PX; ++i) val[i] += (double)hi[i];
        } else {
            vs_poly_eval_n<D, PX, double>(pc.c64[o], u, v, w, val);
        }
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            if (o == 2) {
                key[i] = vs_key32((float)val[i]);
            } else {
                const int q = floor_to_int_fast(val[i]);  // floor, lib/proj_to_grid.py:42-43
                if (o == 0) ci[i] = q; else ri[i] = q;
            }
        }
    }
    int cell[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        fast[i] = fast[i] && (unsigned)ci[i] < (unsigned)p.xsize && (unsigned)ri[i] < (unsigned)p.ysize;
        cell[i] = ri[i] * p.xsize + ci[i];
    }
#pragma unroll
    for (int i = 0; i + 1 < PX; ++i) {   // same-cell neighbours merge in registers (forward chain)
        const bool same = fast[i] && fast[i + 1] && cell[i] == cell[i + 1];
        key[i + 1] = same ? max(key[i + 1], key[i]) : key[i + 1];
        fast[i] = fast[i] && !same;
    }
#pragma unroll
    for (int i = 0; i < PX; ++i)
        if (fast[i]) atomicMax(keygrid + cell[i], key[i]);
    if (any_slow) {  // points outside the fitted altitude range: exact chain, out of line (rare)
#pragma unroll
        for (int i = 0; i < PX; ++i) {
            if (d4[i] > 0.0f && fabsf(uf[i]) <= 1.0f && fabsf(vf[i]) <= 1.0f && !(fabsf(wf[i]) <= 1.0f) &&
                fabsf(wf[i]) < CUDART_INF_F)
                scatter_exact_point(ex, u[i], v[i], w[i], keygrid, nullptr, 0);
        }
    }
}

}  // namespace vsras

// host side (rasterize.cu): kernel parameters of stage A for one view
struct vs_ctx;
void vs_make_raster_params(const vs_ctx* ctx, int H, int W, const double* M, vsras::RasterParams* p, vsras::PolyCoefs* pc);
