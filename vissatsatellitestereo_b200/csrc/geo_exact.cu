// K5: exact float64 converter kernels behind the reference's public converter functions
// (lib/latlonalt_enu_converter.py:36-45, lib/latlon_utm_converter.py:39-63, coordinate_system.py:41-64).
// One thread per point, grid-stride; FP64-pipe bound (~25 double transcendentals per point), so there is
// nothing to stage: loads/stores are coalesced 8-byte accesses.  Compiled with -fmad=false so that the
// operation order of the upstream formulas is what executes.
#include "geo_chain.cuh"
#include "vs_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(const vs_ctx* ctx, int64_t count) {
    int64_t blocks = (count + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

VsGeoParams make_origin(const VsEllipsoidConsts& c, double lat0, double lon0, double alt0) {
    // host mirror of geodetic2ecef for the single origin point (same formulas, host libm)
    VsGeoParams g;
    memset(&g, 0, sizeof(g));
    g.lat0 = lat0;
    g.lon0 = lon0;
    g.alt0 = alt0;
    double lat = lat0 * c.dg2rad, lon = lon0 * c.dg2rad;
    g.sin_lat0 = sin(lat);
    g.cos_lat0 = cos(lat);
    g.sin_lon0 = sin(lon);
    g.cos_lon0 = cos(lon);
    double N = c.a2 / sqrt(c.a2 * (g.cos_lat0 * g.cos_lat0) + c.b2 * (g.sin_lat0 * g.sin_lat0));
    g.x0 = (N + alt0) * g.cos_lat0 * g.cos_lon0;
    g.y0 = (N + alt0) * g.cos_lat0 * g.sin_lon0;
    g.z0 = (N * c.b_over_a_sq + alt0) * g.sin_lat0;
    return g;
}

double utm_lam0(int zone) {
    // PROJ utm setup: lam0 = (zone - 1 + .5) * M_PI / 30. - M_PI
    return ((zone - 1) + .5) * M_PI / 30. - M_PI;
}

__global__ void __launch_bounds__(kThreads) k_enu_to_geodetic(VsEllipsoidConsts c, VsGeoParams g,
                                                              const double* __restrict__ e,
                                                              const double* __restrict__ n,
                                                              const double* __restrict__ u, int64_t count,
                                                              double* __restrict__ lat, double* __restrict__ lon,
                                                              double* __restrict__ alt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z, la, lo, al;
        vs_enu2ecef(g, e[i], n[i], u[i], x, y, z);
        vs_ecef2geodetic(c, x, y, z, la, lo, al);
        lat[i] = la;
        lon[i] = lo;
        alt[i] = al;
    }
}

__global__ void __launch_bounds__(kThreads) k_geodetic_to_enu(VsEllipsoidConsts c, VsGeoParams g,
                                                              const double* __restrict__ lat,
                                                              const double* __restrict__ lon,
                                                              const double* __restrict__ alt, int64_t count,
                                                              double* __restrict__ e, double* __restrict__ n,
                                                              double* __restrict__ u) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z, ee, nn, uu;
        vs_geodetic2ecef(c, lat[i], lon[i], alt[i], x, y, z);
        vs_uvw2enu(g, x - g.x0, y - g.y0, z - g.z0, ee, nn, uu);
        e[i] = ee;
        n[i] = nn;
        u[i] = uu;
    }
}

__global__ void __launch_bounds__(kThreads) k_geodetic_to_utm(VsEllipsoidConsts c, double lam0, double north_off,
                                                              const double* __restrict__ lat,
                                                              const double* __restrict__ lon, int64_t count,
                                                              double* __restrict__ east, double* __restrict__ north) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double E, N;
        vs_utm_forward(c, lam0, north_off, lat[i], lon[i], E, N);
        east[i] = E;
        north[i] = N;
    }
}

__global__ void __launch_bounds__(kThreads) k_utm_to_geodetic(VsEllipsoidConsts c, double lam0, double north_off,
                                                              const double* __restrict__ east,
                                                              const double* __restrict__ north, int64_t count,
                                                              double* __restrict__ lat, double* __restrict__ lon) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double la, lo;
        vs_utm_inverse(c, lam0, north_off, east[i], north[i], la, lo);
        lat[i] = la;
        lon[i] = lo;
    }
}

__global__ void __launch_bounds__(kThreads) k_enu_to_utm(VsEllipsoidConsts c, VsGeoParams g,
                                                         const double* __restrict__ e, const double* __restrict__ n,
                                                         const double* __restrict__ u, int64_t count,
                                                         double* __restrict__ east, double* __restrict__ north,
                                                         double* __restrict__ alt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double E, N, A;
        vs_enu_to_utm_exact(c, g, e[i], n[i], u[i], E, N, A);
        east[i] = E;
        north[i] = N;
        alt[i] = A;
    }
}

}  // namespace

// exported to aoi_fit.cu
VsGeoParams vs_make_geo_params(const VsEllipsoidConsts& c, const vs_aoi& aoi) {
    VsGeoParams g = make_origin(c, aoi.lat0, aoi.lon0, aoi.alt0);
    g.lam0 = utm_lam0(aoi.zone);
    g.north_off = aoi.south ? 10000000.0 : 0.0;
    g.ul_e = aoi.ul_e;
    g.ul_n = aoi.ul_n;
    g.row_res = aoi.row_res;
    g.col_res = aoi.col_res;
    g.xsize = aoi.xsize;
    g.ysize = aoi.ysize;
    return g;
}

#define VS_ENTRY(ctx)                                                       \
    VS_REQUIRE((ctx) != nullptr, "NULL context");                           \
    VsDeviceGuard guard__((ctx)->device);                                   \
    if (!guard__.ok) return vs_cuda_fail(cudaGetLastError(), "cudaSetDevice")

extern "C" {

int vs_enu_to_geodetic(vs_ctx* ctx, const double* e, const double* n, const double* u, int64_t count, double lat0,
                       double lon0, double alt0, double* lat, double* lon, double* alt, void* stream) {
    VS_ENTRY(ctx);
    VS_REQUIRE(count >= 0, "vs_enu_to_geodetic: negative count");
    if (count == 0) return VS_OK;
    VS_REQUIRE(e && n && u && lat && lon && alt, "vs_enu_to_geodetic: NULL array");
    VsEllipsoidConsts c = vs_make_ellipsoid_consts();
    VsGeoParams g = make_origin(c, lat0, lon0, alt0);
    k_enu_to_geodetic<<<grid_for(ctx, count), kThreads, 0, (cudaStream_t)stream>>>(c, g, e, n, u, count, lat, lon, alt);
    VS_CHECK_LAUNCH(ctx, "k_enu_to_geodetic");
    return VS_OK;
}

int vs_geodetic_to_enu(vs_ctx* ctx, const double* lat, const double* lon, const double* alt, int64_t count,
                       double lat0, double lon0, double alt0, double* e, double* n, double* u, void* stream) {
    VS_ENTRY(ctx);
    VS_REQUIRE(count >= 0, "vs_geodetic_to_enu: negative count");
    if (count == 0) return VS_OK;
    VS_REQUIRE(e && n && u && lat && lon && alt, "vs_geodetic_to_enu: NULL array");
    VsEllipsoidConsts c = vs_make_ellipsoid_consts();
    VsGeoParams g = make_origin(c, lat0, lon0, alt0);
    k_geodetic_to_enu<<<grid_for(ctx, count), kThreads, 0, (cudaStream_t)stream>>>(c, g, lat, lon, alt, count, e, n, u);
    VS_CHECK_LAUNCH(ctx, "k_geodetic_to_enu");
    return VS_OK;
}

int vs_geodetic_to_utm(vs_ctx* ctx, const double* lat, const double* lon, int64_t count, int32_t zone, int32_t south,
                       double* east, double* north, void* stream) {
    VS_ENTRY(ctx);
    VS_REQUIRE(count >= 0, "vs_geodetic_to_utm: negative count");
    VS_REQUIRE(zone >= 1 && zone <= 60, "vs_geodetic_to_utm: zone must be 1..60");
    if (count == 0) return VS_OK;
    VS_REQUIRE(lat && lon && east && north, "vs_geodetic_to_utm: NULL array");
    VsEllipsoidConsts c = vs_make_ellipsoid_consts();
    k_geodetic_to_utm<<<grid_for(ctx, count), kThreads, 0, (cudaStream_t)stream>>>(
        c, utm_lam0(zone), south ? 10000000.0 : 0.0, lat, lon, count, east, north);
    VS_CHECK_LAUNCH(ctx, "k_geodetic_to_utm");
    return VS_OK;
}

int vs_utm_to_geodetic(vs_ctx* ctx, const double* east, const double* north, int64_t count, int32_t zone,
                       int32_t south, double* lat, double* lon, void* stream) {
    VS_ENTRY(ctx);
    VS_REQUIRE(count >= 0, "vs_utm_to_geodetic: negative count");
    VS_REQUIRE(zone >= 1 && zone <= 60, "vs_utm_to_geodetic: zone must be 1..60");
    if (count == 0) return VS_OK;
    VS_REQUIRE(lat && lon && east && north, "vs_utm_to_geodetic: NULL array");
    VsEllipsoidConsts c = vs_make_ellipsoid_consts();
    k_utm_to_geodetic<<<grid_for(ctx, count), kThreads, 0, (cudaStream_t)stream>>>(
        c, utm_lam0(zone), south ? 10000000.0 : 0.0, east, north, count, lat, lon);
    VS_CHECK_LAUNCH(ctx, "k_utm_to_geodetic");
    return VS_OK;
}

int vs_enu_to_utm(vs_ctx* ctx, const double* e, const double* n, const double* u, int64_t count, double lat0,
                  double lon0, double alt0, int32_t zone, int32_t south, double* east, double* north, double* alt,
                  void* stream) {
    VS_ENTRY(ctx);
    VS_REQUIRE(count >= 0, "vs_enu_to_utm: negative count");
    VS_REQUIRE(zone >= 1 && zone <= 60, "vs_enu_to_utm: zone must be 1..60");
    if (count == 0) return VS_OK;
    VS_REQUIRE(e && n && u && east && north && alt, "vs_enu_to_utm: NULL array");
    VsEllipsoidConsts c = vs_make_ellipsoid_consts();
    VsGeoParams g = make_origin(c, lat0, lon0, alt0);
    g.lam0 = utm_lam0(zone);
    g.north_off = south ? 10000000.0 : 0.0;
    k_enu_to_utm<<<grid_for(ctx, count), kThreads, 0, (cudaStream_t)stream>>>(c, g, e, n, u, count, east, north, alt);
    VS_CHECK_LAUNCH(ctx, "k_enu_to_utm");
    return VS_OK;
}

}  // extern "C"
