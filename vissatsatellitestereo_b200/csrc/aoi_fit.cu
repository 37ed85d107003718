// vs_set_aoi: per-AOI setup of the fused rasteriser.
//
// The reference converts every unprojected point through pymap3d.enu2geodetic and PROJ's utm
// (aggregate_2p5d_util.py:96-97): ~25 float64 transcendentals per point, two orders of magnitude more FP64
// work than a kernel that streams 4-byte depths can afford.  Over one AOI the composite map
// ENU -> (fractional col, fractional row, altitude) is smooth, so it is replaced by a trivariate
// polynomial fitted HERE against the exact device chain (geo_chain.cuh) and validated on held-out points
// down to the chain's own float64 noise (~1e-9 m).  All chain evaluations run on the GPU; the host only
// solves the small least-squares problem (long double Householder QR).
#include <math.h>
#include <string.h>

#include <vector>

#include "geo_chain.cuh"
#include "poly.cuh"
#include "vs_common.cuh"

VsGeoParams vs_make_geo_params(const VsEllipsoidConsts& c, const vs_aoi& aoi);

namespace {

// (E, N, alt) -> ENU, exact chain (utm inverse, geodetic2ecef, uvw2enu)
__global__ void k_utm_to_enu(VsEllipsoidConsts c, VsGeoParams g, const double* __restrict__ in, int n,
                             double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lat, lon, x, y, z, e, nn, u;
    vs_utm_inverse(c, g.lam0, g.north_off, in[3 * i], in[3 * i + 1], lat, lon);
    vs_geodetic2ecef(c, lat, lon, in[3 * i + 2], x, y, z);
    vs_uvw2enu(g, x - g.x0, y - g.y0, z - g.z0, e, nn, u);
    out[3 * i] = e;
    out[3 * i + 1] = nn;
    out[3 * i + 2] = u;
}

// ENU -> (E, N, alt), exact chain
__global__ void k_enu_to_utm3(VsEllipsoidConsts c, VsGeoParams g, const double* __restrict__ in, int n,
                              double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double E, N, A;
    vs_enu_to_utm_exact(c, g, in[3 * i], in[3 * i + 1], in[3 * i + 2], E, N, A);
    out[3 * i] = E;
    out[3 * i + 1] = N;
    out[3 * i + 2] = A;
}

int run_map(vs_ctx* ctx, bool inverse, const VsEllipsoidConsts& c, const VsGeoParams& g, const std::vector<double>& in,
            std::vector<double>& out) {
    int n = (int)(in.size() / 3);
    out.resize(in.size());
    int rc = vs_ensure_scratch(ctx, 2 * in.size());
    if (rc) return rc;
    double* d_in = ctx->d_scratch;
    double* d_out = ctx->d_scratch + in.size();
    VS_CUDA(cudaMemcpy(d_in, in.data(), in.size() * sizeof(double), cudaMemcpyHostToDevice));
    int blocks = (n + 127) / 128;
    if (inverse)
        k_utm_to_enu<<<blocks, 128>>>(c, g, d_in, n, d_out);
    else
        k_enu_to_utm3<<<blocks, 128>>>(c, g, d_in, n, d_out);
    VS_CHECK_LAUNCH(ctx, inverse ? "k_utm_to_enu" : "k_enu_to_utm3");
    VS_CUDA(cudaMemcpy(out.data(), d_out, in.size() * sizeof(double), cudaMemcpyDeviceToHost));
    return VS_OK;
}

// Least squares min ||A x - b|| for several right-hand sides, Householder QR in long double.
// A is m x n row-major (destroyed), B is m x nrhs row-major (destroyed); X is n x nrhs.
bool lstsq_qr(std::vector<long double>& A, std::vector<long double>& B, int m, int n, int nrhs,
              std::vector<long double>& X) {
    for (int k = 0; k < n; ++k) {
        long double norm = 0;
        for (int i = k; i < m; ++i) norm += A[i * n + k] * A[i * n + k];
        norm = sqrtl(norm);
        if (norm == 0) return false;
        long double alpha = A[k * n + k] > 0 ? -norm : norm;
        std::vector<long double> v(m - k);
        for (int i = k; i < m; ++i) v[i - k] = A[i * n + k];
        v[0] -= alpha;
        long double vnorm2 = 0;
        for (long double t : v) vnorm2 += t * t;
        if (vnorm2 == 0) continue;
        for (int j = k; j < n; ++j) {
            long double dot = 0;
            for (int i = k; i < m; ++i) dot += v[i - k] * A[i * n + j];
            long double f = 2 * dot / vnorm2;
            for (int i = k; i < m; ++i) A[i * n + j] -= f * v[i - k];
        }
        for (int j = 0; j < nrhs; ++j) {
            long double dot = 0;
            for (int i = k; i < m; ++i) dot += v[i - k] * B[i * nrhs + j];
            long double f = 2 * dot / vnorm2;
            for (int i = k; i < m; ++i) B[i * nrhs + j] -= f * v[i - k];
        }
    }
    X.assign((size_t)n * nrhs, 0);
    for (int j = 0; j < nrhs; ++j) {
        for (int k = n - 1; k >= 0; --k) {
            long double s = B[k * nrhs + j];
            for (int l = k + 1; l < n; ++l) s -= A[k * n + l] * X[l * nrhs + j];
            if (A[k * n + k] == 0) return false;
            X[k * nrhs + j] = s / A[k * n + k];
        }
    }
    return true;
}

double eval_poly_host(int D, const double* c, double u, double v, double w) {
    switch (D) {
        case 3: return vs_poly_eval<3>(c, u, v, w);
        case 4: return vs_poly_eval<4>(c, u, v, w);
        case 5: return vs_poly_eval<5>(c, u, v, w);
    }
    return NAN;
}

float eval_poly_host_f32(int D, const float* c, float u, float v, float w) {
    switch (D) {
        case 3: return vs_poly_eval<3, float>(c, u, v, w);
        case 4: return vs_poly_eval<4, float>(c, u, v, w);
        case 5: return vs_poly_eval<5, float>(c, u, v, w);
    }
    return NAN;
}

// deterministic quasi-random numbers in [-1, 1] for the held-out validation points
double halton(int index, int base) {
    double f = 1, r = 0;
    int i = index;
    while (i > 0) {
        f /= base;
        r += f * (i % base);
        i /= base;
    }
    return 2 * r - 1;
}

}  // namespace

int vs_upload_exact_params(vs_ctx* ctx) {
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");
    VsExactParams ex;
    memset(&ex, 0, sizeof(ex));
    ex.c = vs_make_ellipsoid_consts();
    ex.g = ctx->geo;
    for (int d = 0; d < 3; ++d) {
        ex.center[d] = ctx->poly.center[d];
        ex.half[d] = 1.0 / ctx->poly.inv_half[d];
    }
    ex.eps = ctx->ambiguity_eps;
    if (!ctx->d_exact) VS_CUDA(cudaMalloc(&ctx->d_exact, sizeof(VsExactParams)));
    VS_CUDA(cudaMemcpy(ctx->d_exact, &ex, sizeof(ex), cudaMemcpyHostToDevice));
    return VS_OK;
}

extern "C" int vs_set_aoi(vs_ctx* ctx, const vs_aoi* aoi, int max_degree, vs_fit_info* info) {
    VS_REQUIRE(ctx != nullptr && aoi != nullptr, "vs_set_aoi: NULL argument");
    VS_REQUIRE(aoi->xsize > 0 && aoi->ysize > 0, "vs_set_aoi: grid size must be positive");
    VS_REQUIRE((int64_t)aoi->xsize * aoi->ysize < (int64_t)1 << 31, "vs_set_aoi: grid has more than 2^31 cells");
    VS_REQUIRE(aoi->row_res > 0 && aoi->col_res > 0, "vs_set_aoi: resolution must be positive");
    VS_REQUIRE(aoi->zone >= 1 && aoi->zone <= 60, "vs_set_aoi: zone must be 1..60");
    VS_REQUIRE(aoi->alt_hi > aoi->alt_lo, "vs_set_aoi: alt_hi must exceed alt_lo");
    VS_REQUIRE(max_degree == 0 || (max_degree >= 3 && max_degree <= VS_MAX_DEGREE), "vs_set_aoi: max_degree must be 0 or 3..5");
    VS_REQUIRE(fabs(aoi->lat0) < 89.0, "vs_set_aoi: |lat0| must be < 89 deg");
    VsDeviceGuard guard(ctx->device);
    if (!guard.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice");

    VsEllipsoidConsts c = vs_make_ellipsoid_consts();
    VsGeoParams g = vs_make_geo_params(c, *aoi);
    ctx->aoi = *aoi;
    ctx->geo = g;
    memset(&ctx->poly, 0, sizeof(ctx->poly));
    memset(&ctx->fit, 0, sizeof(ctx->fit));
    ctx->aoi_set = true;

    // ---- ENU box that contains the grid over [alt_lo, alt_hi]
    const double ext_e = aoi->xsize * aoi->col_res, ext_n = aoi->ysize * aoi->row_res;
    std::vector<double> corners, enu;
    for (int a = 0; a < 2; ++a)
        for (int iy = 0; iy <= 2; ++iy)
            for (int ix = 0; ix <= 2; ++ix) {
                corners.push_back(aoi->ul_e + 0.5 * ix * ext_e);
                corners.push_back(aoi->ul_n - 0.5 * iy * ext_n);
                corners.push_back(a ? aoi->alt_hi : aoi->alt_lo);
            }
    int rc = run_map(ctx, true, c, g, corners, enu);
    if (rc) return rc;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < enu.size() / 3; ++i)
        for (int d = 0; d < 3; ++d) {
            double v = enu[3 * i + d];
            if (!isfinite(v)) {
                vs_set_error("vs_set_aoi: AOI corner does not map to a finite ENU position (zone/hemisphere wrong?)");
                return VS_ERR_INVALID;
            }
            lo[d] = fmin(lo[d], v);
            hi[d] = fmax(hi[d], v);
        }
    VsPoly& P = ctx->poly;
    for (int d = 0; d < 3; ++d) {
        P.center[d] = 0.5 * (lo[d] + hi[d]);
        double half = 0.5 * (hi[d] - lo[d]);
        half = (d < 2) ? half * 1.02 + 4.0 * fmax(aoi->row_res, aoi->col_res) + 5.0 : half * 1.02 + 1.0;
        P.inv_half[d] = 1.0 / half;
        ctx->fit.box_center[d] = P.center[d];
        ctx->fit.box_half[d] = half;
    }
    if (info) *info = ctx->fit;
    rc = vs_upload_exact_params(ctx);
    if (rc) return rc;
    if (max_degree == 0) return VS_OK;  // exact chain for every point

    // ---- held-out validation points (Halton), shared by all candidate degrees
    const int n_test = 4096;
    std::vector<double> test_uvw(3 * n_test), test_in(3 * n_test), test_out;
    for (int i = 0; i < n_test; ++i) {
        double u = halton(i + 1, 2), v = halton(i + 1, 3), w = halton(i + 1, 5);
        if (i < 64) {  // box faces and corners
            u = (i & 1) ? 1.0 : -1.0;
            if (i & 8) v = (i & 2) ? 1.0 : -1.0;
            if (i & 16) w = (i & 4) ? 1.0 : -1.0;
        }
        test_uvw[3 * i] = u;
        test_uvw[3 * i + 1] = v;
        test_uvw[3 * i + 2] = w;
        for (int d = 0; d < 3; ++d) test_in[3 * i + d] = P.center[d] + test_uvw[3 * i + d] / P.inv_half[d];
    }
    rc = run_map(ctx, false, c, g, test_in, test_out);
    if (rc) return rc;
    // the polynomial sees (p - center) * inv_half computed in double on the device; use the same here
    for (int i = 0; i < n_test; ++i)
        for (int d = 0; d < 3; ++d) test_uvw[3 * i + d] = (test_in[3 * i + d] - P.center[d]) * P.inv_half[d];

    const double tol_m = 2.5e-8;  // chain noise is ~1e-9 (E), ~4e-9 (N at 6e6 m), ~2e-9 (alt)
    for (int D = 3; D <= max_degree; ++D) {
        const int nt = vs_poly_terms(D);
        const int nu = D + 4, nw = D + 3;
        const int m = nu * nu * nw;
        std::vector<double> nodes(3 * m), node_uvw(3 * m), vals;
        int q = 0;
        for (int a = 0; a < nu; ++a)
            for (int b = 0; b < nu; ++b)
                for (int d = 0; d < nw; ++d, ++q) {
                    double uvw[3] = {cos(M_PI * (a + 0.5) / nu), cos(M_PI * (b + 0.5) / nu), cos(M_PI * (d + 0.5) / nw)};
                    for (int t = 0; t < 3; ++t) nodes[3 * q + t] = P.center[t] + uvw[t] / P.inv_half[t];
                }
        rc = run_map(ctx, false, c, g, nodes, vals);
        if (rc) return rc;
        std::vector<long double> A((size_t)m * nt), B((size_t)m * 3), X;
        for (int r = 0; r < m; ++r) {
            long double u = (long double)((nodes[3 * r] - P.center[0]) * P.inv_half[0]);
            long double v = (long double)((nodes[3 * r + 1] - P.center[1]) * P.inv_half[1]);
            long double w = (long double)((nodes[3 * r + 2] - P.center[2]) * P.inv_half[2]);
            long double pu = 1;
            for (int i = 0; i <= D; ++i, pu *= u) {
                long double pv = 1;
                for (int j = 0; j <= D - i; ++j, pv *= v) {
                    long double pw = 1;
                    for (int k = 0; k <= D - i - j; ++k, pw *= w) A[(size_t)r * nt + vs_poly_index(D, i, j, k)] = pu * pv * pw;
                }
            }
            // lib/proj_to_grid.py:42-43 (fractional, before floor)
            B[r * 3 + 0] = (long double)((vals[3 * r] - aoi->ul_e) / aoi->col_res);
            B[r * 3 + 1] = (long double)((aoi->ul_n - vals[3 * r + 1]) / aoi->row_res);
            B[r * 3 + 2] = (long double)vals[3 * r + 2];
        }
        if (!lstsq_qr(A, B, m, nt, 3, X)) continue;
        bool accepted = false;
        // cheapest evaluation first: float64 only for the terms of degree <= d64, float32 above; 0 = all float64
        const int d64_order[3] = {1, 2, 0};
        for (int q = 0; q < 3 && !accepted; ++q) {
            const int d64 = d64_order[q];
            VsPoly cand = P;
            cand.degree = D;
            cand.n_terms = nt;
            cand.d64 = d64;
            for (int t = 0; t < nt; ++t)
                for (int o = 0; o < 3; ++o) cand.coef[o][t] = (double)X[(size_t)t * 3 + o];
            if (d64 > 0)
                for (int i = 0; i <= D; ++i)
                    for (int j = 0; j <= D - i; ++j)
                        for (int k = 0; k <= D - i - j; ++k)
                            for (int o = 0; o < 3; ++o) {
                                const double cf = cand.coef[o][vs_poly_index(D, i, j, k)];
                                if (i + j + k <= d64) {
                                    cand.coefLo[o][vs_poly_index(d64, i, j, k)] = cf;
                                    cand.coefR[o][vs_poly_index(D, i, j, k)] = 0.0f;
                                } else {
                                    cand.coefR[o][vs_poly_index(D, i, j, k)] = (float)cf;
                                }
                            }
            double err_cells = 0, err_alt = 0, err_m = 0;
            for (int i = 0; i < n_test; ++i) {
                const double u = test_uvw[3 * i], v = test_uvw[3 * i + 1], w = test_uvw[3 * i + 2];
                const double want[3] = {(test_out[3 * i] - aoi->ul_e) / aoi->col_res,
                                        (aoi->ul_n - test_out[3 * i + 1]) / aoi->row_res, test_out[3 * i + 2]};
                double got[3];
                for (int o = 0; o < 3; ++o) {
                    // exactly what the device evaluates
                    const double hi = d64 ? (double)eval_poly_host_f32(D, cand.coefR[o], (float)u, (float)v, (float)w) : 0.0;
                    if (d64 == 1) got[o] = vs_poly_eval<1>(cand.coefLo[o], u, v, w) + hi;
                    else if (d64 == 2) got[o] = vs_poly_eval<2>(cand.coefLo[o], u, v, w) + hi;
                    else got[o] = eval_poly_host(D, cand.coef[o], u, v, w);
                }
                const double dc = fabs(got[0] - want[0]), dr = fabs(got[1] - want[1]), da = fabs(got[2] - want[2]);
                err_cells = fmax(err_cells, fmax(dc, dr));
                err_m = fmax(err_m, fmax(dc * aoi->col_res, dr * aoi->row_res));
                err_alt = fmax(err_alt, da);
            }
            const bool ok = err_m <= tol_m && err_alt <= tol_m;
            if (ok || (D == max_degree && d64 == 0)) {
                // keep the last attempt's errors for diagnostics; only a validated fit is activated
                ctx->fit.degree = ok ? D : 0;
                ctx->fit.n_terms = ok ? nt : 0;
                ctx->fit.mixed = ok ? d64 : 0;
                ctx->fit.max_err_cells = err_cells;
                ctx->fit.max_err_alt_m = err_alt;
                if (ok) P = cand;
                accepted = true;
            }
        }
        if (accepted) break;
    }
    if (info) *info = ctx->fit;
    return VS_OK;
}

// Evaluate the activated polynomial on explicit ENU points (host arrays) exactly as K1 evaluates it (same Horner
// order, same float64/float32 split).  Diagnostics: lets a test compare the fit with an independent exact map.
extern "C" int vs_fit_eval(vs_ctx* ctx, const double* enu, int64_t n, double* colf, double* rowf, double* alt) {
    VS_REQUIRE(ctx != nullptr, "vs_fit_eval: NULL context");
    if (!ctx->aoi_set || ctx->poly.degree == 0) {
        vs_set_error("vs_fit_eval: no validated polynomial (vs_set_aoi not called, or exact mode)");
        return VS_ERR_STATE;
    }
    VS_REQUIRE(n >= 0 && (n == 0 || (enu && colf && rowf && alt)), "vs_fit_eval: NULL argument");
    const VsPoly& P = ctx->poly;
    double* outs[3] = {colf, rowf, alt};
    for (int64_t i = 0; i < n; ++i) {
        const double u = (enu[3 * i] - P.center[0]) * P.inv_half[0];
        const double v = (enu[3 * i + 1] - P.center[1]) * P.inv_half[1];
        const double w = (enu[3 * i + 2] - P.center[2]) * P.inv_half[2];
        for (int o = 0; o < 3; ++o) {
            const double hi = P.d64 ? (double)eval_poly_host_f32(P.degree, P.coefR[o], (float)u, (float)v, (float)w) : 0.0;
            if (P.d64 == 1) outs[o][i] = vs_poly_eval<1>(P.coefLo[o], u, v, w) + hi;
            else if (P.d64 == 2) outs[o][i] = vs_poly_eval<2>(P.coefLo[o], u, v, w) + hi;
            else outs[o][i] = eval_poly_host(P.degree, P.coef[o], u, v, w);
        }
    }
    return VS_OK;
}
