// cv2.medianBlur(float32, 3) semantics (produce_dsm.py:58, aggregate_2p5d.py:81), including NaNs.
//
// OpenCV sorts the 3x3 window with a fixed 19-exchange network op(a,b): t=a; a=min(a,b); b=max(b,t) and
// returns p4.  Its vectorised body (columns 1..W-2 when W >= lanes+2) uses SIMD min/max, for which
// min(a,b) = (a<b)?a:b and max(b,t) = (b>t)?b:t: the pair is exchanged unless a<b, so a pair holding a NaN
// IS exchanged.  Its scalar body (first/last column, narrow images) uses std::min/std::max:
// min(a,b) = (b<a)?b:a, max(b,t) = (b<t)?t:b: the pair is exchanged only if b<a, so a pair holding a NaN is
// NOT exchanged.  Both reduce to one predicate per exchange.  (verified against cv2 4.13.)
#pragma once

template <bool SIMD>
__device__ __forceinline__ void vs_cex(float& a, float& b) {
    const bool swap = SIMD ? !(a < b) : (b < a);
    const float lo = swap ? b : a;
    const float hi = swap ? a : b;
    a = lo;
    b = hi;
}

template <bool SIMD>
__device__ __forceinline__ float vs_median9_net(float p0, float p1, float p2, float p3, float p4, float p5, float p6,
                                                float p7, float p8) {
    vs_cex<SIMD>(p1, p2); vs_cex<SIMD>(p4, p5); vs_cex<SIMD>(p7, p8); vs_cex<SIMD>(p0, p1);
    vs_cex<SIMD>(p3, p4); vs_cex<SIMD>(p6, p7); vs_cex<SIMD>(p1, p2); vs_cex<SIMD>(p4, p5);
    vs_cex<SIMD>(p7, p8); vs_cex<SIMD>(p0, p3); vs_cex<SIMD>(p5, p8); vs_cex<SIMD>(p4, p7);
    vs_cex<SIMD>(p3, p6); vs_cex<SIMD>(p1, p4); vs_cex<SIMD>(p2, p5); vs_cex<SIMD>(p4, p7);
    vs_cex<SIMD>(p4, p2); vs_cex<SIMD>(p6, p4); vs_cex<SIMD>(p4, p2);
    return p4;
}

// OpenCV's special case for single-row / single-column images: 1-D 3-tap median, scalar semantics.
__device__ __forceinline__ float vs_median3_line(float p0, float p1, float p2) {
    vs_cex<false>(p0, p1);
    vs_cex<false>(p1, p2);
    vs_cex<false>(p0, p1);
    return p1;
}

__device__ __forceinline__ float vs_min_t(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float vs_max_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double vs_min_t(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double vs_max_t(double a, double b) { return fmax(a, b); }

// (lo + hi) / 2 as numpy computes it for a float64 list, rounded to T.  For float the double detour is not needed:
// halving is an exponent shift, so RN_f32((lo + hi) / 2) == RN_f32(lo + hi) / 2 (no overflow or subnormals for heights).
__device__ __forceinline__ float vs_mean2(float lo, float hi) { return __fmul_rn(__fadd_rn(lo, hi), 0.5f); }
__device__ __forceinline__ double vs_mean2(double lo, double hi) { return (lo + hi) * 0.5; }

// Median of the k non-NaN values among 8 candidates (lib/proj_to_grid.py:70-79 -> np.median of a list):
// odd k -> middle element, even k -> mean of the two middle elements (computed in double, as numpy does
// for a float64 list), k == 0 -> NaN.  T is float (32-bit key path) or double (64-bit key path).
template <typename T>
__device__ __forceinline__ T vs_median_of_valid8(T (&v)[8]) {
    int k = 0;
    const T big = (T)CUDART_INF;
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // branch-free: NaN -> +inf sorts last
        const bool ok = (v[i] == v[i]);
        k += ok ? 1 : 0;
        v[i] = ok ? v[i] : big;
    }
    // Batcher odd-even merge sort for 8 inputs (19 exchanges); values are NaN-free here
#define VS_CE(i, j) { const T lo = vs_min_t(v[i], v[j]); const T hi = vs_max_t(v[i], v[j]); v[i] = lo; v[j] = hi; }
    VS_CE(0, 1) VS_CE(2, 3) VS_CE(4, 5) VS_CE(6, 7)
    VS_CE(0, 2) VS_CE(1, 3) VS_CE(4, 6) VS_CE(5, 7)
    VS_CE(1, 2) VS_CE(5, 6)
    VS_CE(0, 4) VS_CE(1, 5) VS_CE(2, 6) VS_CE(3, 7)
    VS_CE(2, 4) VS_CE(3, 5)
    VS_CE(1, 2) VS_CE(3, 4) VS_CE(5, 6)
#undef VS_CE
    // middle elements of the k valid ones: positions (k-1)/2 (0..3) and k/2 (0..4), so v[5..7] are never read.
    // Two-level select on the index bits instead of a compare per candidate.
    const int ilo = (k - 1) >> 1, ihi = k >> 1;
    const bool lo_b0 = (ilo & 1) != 0, lo_b1 = (ilo & 2) != 0;
    const bool hi_b0 = (ihi & 1) != 0, hi_b1 = (ihi & 2) != 0;
    const T lo = lo_b1 ? (lo_b0 ? v[3] : v[2]) : (lo_b0 ? v[1] : v[0]);      // k == 0: ilo = -1 -> v[3]; unused
    const T hi4 = hi_b1 ? (hi_b0 ? v[3] : v[2]) : (hi_b0 ? v[1] : v[0]);
    const T hi = (ihi == 4) ? v[4] : hi4;
    const T avg = vs_mean2(lo, hi);
    const T med = (ilo == ihi) ? hi : avg;
    return k == 0 ? (T)CUDART_NAN : med;
}
