// Trivariate polynomial (total degree D) in box-normalised ENU coordinates; evaluation order shared by
// host (fit validation) and device (rasteriser) so that both produce the same bits for the same inputs.
#pragma once

#include <math.h>

// flat index of the coefficient of u^i v^j w^k, terms ordered i-major, then j, then k, with i+j+k <= D
__host__ __device__ constexpr int vs_poly_index(int D, int i, int j, int k) {
    int idx = 0;
    for (int ii = 0; ii < i; ++ii) {
        int m = D - ii;  // j + k <= m  -> (m+1)(m+2)/2 terms
        idx += (m + 1) * (m + 2) / 2;
    }
    for (int jj = 0; jj < j; ++jj) idx += (D - i - jj) + 1;
    return idx + k;
}
__host__ __device__ constexpr int vs_poly_terms(int D) { return (D + 1) * (D + 2) * (D + 3) / 6; }

// Nested Horner: sum_i u^i ( sum_j v^j ( sum_k w^k c_ijk ) ), n_terms - 1 fused multiply-adds.
__host__ __device__ __forceinline__ double vs_fma_t(double a, double b, double c) { return fma(a, b, c); }
__host__ __device__ __forceinline__ float vs_fma_t(float a, float b, float c) { return fmaf(a, b, c); }

template <int D, typename T = double>
__host__ __device__ __forceinline__ T vs_poly_eval(const T* __restrict__ c, T u, T v, T w) {
    T r = 0;
#pragma unroll
    for (int i = D; i >= 0; --i) {
        T a = 0;
#pragma unroll
        for (int j = D - i; j >= 0; --j) {
            const int kmax = D - i - j;
            T b = c[vs_poly_index(D, i, j, kmax)];
#pragma unroll
            for (int k = kmax - 1; k >= 0; --k) b = vs_fma_t(b, w, c[vs_poly_index(D, i, j, k)]);
            a = (j == D - i) ? b : vs_fma_t(a, v, b);
        }
        r = (i == D) ? a : vs_fma_t(r, u, a);
    }
    return r;
}

#ifdef __CUDACC__
// float32, L points (L even) in lockstep on sm_100's packed FFMA2: one instruction does the Horner step of two points
// (the coefficient is a broadcast scalar operand), which halves the issue slots of the float32 part.  Each lane is an
// IEEE fma, so the result is bit-identical to vs_poly_eval<D, float>.
template <int D, int L>
__device__ __forceinline__ void vs_poly_eval_n_f32x2(const float* __restrict__ c, const float (&u)[L], const float (&v)[L],
                                                     const float (&w)[L], float (&out)[L]) {
    static_assert(L % 2 == 0, "pairs of points");
    constexpr int H = L / 2;
    float2 U[H], V[H], W[H], r[H], a[H], b[H];
#pragma unroll
    for (int p = 0; p < H; ++p) {
        U[p] = make_float2(u[2 * p], u[2 * p + 1]);
        V[p] = make_float2(v[2 * p], v[2 * p + 1]);
        W[p] = make_float2(w[2 * p], w[2 * p + 1]);
    }
#pragma unroll
    for (int i = D; i >= 0; --i) {
#pragma unroll
        for (int j = D - i; j >= 0; --j) {
            const int kmax = D - i - j;
            const float ck = c[vs_poly_index(D, i, j, kmax)];
#pragma unroll
            for (int p = 0; p < H; ++p) b[p] = make_float2(ck, ck);
#pragma unroll
            for (int k = kmax - 1; k >= 0; --k) {
                const float cc = c[vs_poly_index(D, i, j, k)];
#pragma unroll
                for (int p = 0; p < H; ++p) b[p] = __ffma2_rn(b[p], W[p], make_float2(cc, cc));
            }
#pragma unroll
            for (int p = 0; p < H; ++p) a[p] = (j == D - i) ? b[p] : __ffma2_rn(a[p], V[p], b[p]);
        }
#pragma unroll
        for (int p = 0; p < H; ++p) r[p] = (i == D) ? a[p] : __ffma2_rn(r[p], U[p], a[p]);
    }
#pragma unroll
    for (int p = 0; p < H; ++p) {
        out[2 * p] = r[p].x;
        out[2 * p + 1] = r[p].y;
    }
}
#endif

// The same evaluation for L points in lockstep (terms outer, points inner): every coefficient is fetched once
// and used L times, which matters on sm_100a where a DFMA cannot take a constant-bank operand.
// Bit-identical to L calls of vs_poly_eval.
template <int D, int L, typename T>
__device__ __forceinline__ void vs_poly_eval_n(const T* __restrict__ c, const T (&u)[L], const T (&v)[L], const T (&w)[L],
                                               T (&out)[L]) {
    T r[L], a[L], b[L];
#pragma unroll
    for (int i = D; i >= 0; --i) {
#pragma unroll
        for (int j = D - i; j >= 0; --j) {
            const int kmax = D - i - j;
            const T ck = c[vs_poly_index(D, i, j, kmax)];
#pragma unroll
            for (int p = 0; p < L; ++p) b[p] = ck;
#pragma unroll
            for (int k = kmax - 1; k >= 0; --k) {
                const T cc = c[vs_poly_index(D, i, j, k)];
#pragma unroll
                for (int p = 0; p < L; ++p) b[p] = vs_fma_t(b[p], w[p], cc);
            }
#pragma unroll
            for (int p = 0; p < L; ++p) a[p] = (j == D - i) ? b[p] : vs_fma_t(a[p], v[p], b[p]);
        }
#pragma unroll
        for (int p = 0; p < L; ++p) r[p] = (i == D) ? a[p] : vs_fma_t(r[p], u[p], a[p]);
    }
#pragma unroll
    for (int p = 0; p < L; ++p) out[p] = r[p];
}
