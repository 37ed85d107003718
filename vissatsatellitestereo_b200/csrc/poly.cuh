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
template <int D>
__host__ __device__ __forceinline__ double vs_poly_eval(const double* __restrict__ c, double u, double v, double w) {
    double r = 0.0;
#pragma unroll
    for (int i = D; i >= 0; --i) {
        double a = 0.0;
#pragma unroll
        for (int j = D - i; j >= 0; --j) {
            const int kmax = D - i - j;
            double b = c[vs_poly_index(D, i, j, kmax)];
#pragma unroll
            for (int k = kmax - 1; k >= 0; --k) b = fma(b, w, c[vs_poly_index(D, i, j, k)]);
            a = (j == D - i) ? b : fma(a, v, b);
        }
        r = (i == D) ? a : fma(r, u, a);
    }
    return r;
}
