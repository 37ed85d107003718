// Exact float64 geodetic chain (device): the arithmetic the reference delegates to pymap3d 1.7.15 and
// PROJ 6.2 (through pyproj 2.4.0), restated formula by formula in the upstream operation order.
//   lib/latlonalt_enu_converter.py:36-45  -> pymap3d geodetic2enu / enu2geodetic (You 2000 closed form)
//   lib/latlon_utm_converter.py:50-51,61-62 -> PROJ +proj=utm = Poder/Engsager extended TM, 6th order
// Used by the public converter kernels (geo_exact.cu), by vs_set_aoi to sample the map it fits, and as the
// per-point slow path of the fused rasteriser.
#pragma once

#include <math_constants.h>

#include "vs_common.cuh"

#define VS_ETMERC_ORDER 6

struct VsEllipsoidConsts {
    // pymap3d Ellipsoid('wgs84')
    double a, b, E2, E;       // E2 = a^2 - b^2, E = sqrt(E2)
    double a2, b2, b_over_a_sq, a_over_b;
    // PROJ etmerc for WGS84
    double proj_a, Qn, Zb;
    double cgb[VS_ETMERC_ORDER], cbg[VS_ETMERC_ORDER], utg[VS_ETMERC_ORDER], gtu[VS_ETMERC_ORDER];
    double dg2rad, rad2dg;
};

// Filled once on the host (api.cu) and passed to kernels by value.
VsEllipsoidConsts vs_make_ellipsoid_consts();

// Everything the per-point exact slow path of the fused rasteriser needs; lives in device memory
// (vs_ctx::d_exact, uploaded by vs_set_aoi) so that the out-of-line slow path takes one pointer.
struct VsExactParams {
    VsEllipsoidConsts c;
    VsGeoParams g;
    double center[3], half[3];  // ENU box of the polynomial
    double eps;                 // ambiguity threshold (cells)
};

// ---------------------------------------------------------------------------------------------------------
// pymap3d
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vs_geodetic2ecef(const VsEllipsoidConsts& c, double lat_deg, double lon_deg, double alt,
                                                 double& x, double& y, double& z) {
    double lat = lat_deg * c.dg2rad;  // numpy.radians: x * (pi/180)
    double lon = lon_deg * c.dg2rad;
    double sl, cl, so, co;
    sincos(lat, &sl, &cl);
    sincos(lon, &so, &co);
    // get_radius_normal: a**2 / sqrt(a**2 * cos(lat)**2 + b**2 * sin(lat)**2)
    double N = c.a2 / sqrt(c.a2 * (cl * cl) + c.b2 * (sl * sl));
    x = (N + alt) * cl * co;
    y = (N + alt) * cl * so;
    z = (N * c.b_over_a_sq + alt) * sl;
}

__device__ __forceinline__ void vs_ecef2geodetic(const VsEllipsoidConsts& c, double x, double y, double z,
                                                 double& lat_deg, double& lon_deg, double& alt) {
    double r2 = x * x + y * y + z * z;  // r = sqrt(.) ; r**2 is re-squared upstream; difference is < 1 ulp
    double r = sqrt(r2);
    double rr = r * r;
    double t = rr - c.E2;
    double u = sqrt(0.5 * t + 0.5 * sqrt(t * t + 4.0 * c.E2 * (z * z)));
    double Q = hypot(x, y);
    double huE = hypot(u, c.E);
    double Beta = atan(huE / u * z / Q);
    double sB, cB;
    sincos(Beta, &sB, &cB);
    double eps = ((c.b * u - c.a * huE + c.E2) * sB) / (c.a * huE * 1.0 / cB - c.E2 * cB);
    Beta += eps;
    lat_deg = atan(c.a_over_b * tan(Beta)) * c.rad2dg;
    lon_deg = atan2(y, x) * c.rad2dg;
    sincos(Beta, &sB, &cB);
    alt = hypot(z - c.b * sB, Q - c.a * cB);
    bool inside = (x * x) / c.a2 + (y * y) / c.a2 + (z * z) / c.b2 < 1.0;
    if (inside) alt = -alt;
}

// enu2uvw + origin ECEF (enu2ecef)
__device__ __forceinline__ void vs_enu2ecef(const VsGeoParams& g, double e, double n, double up, double& x, double& y,
                                            double& z) {
    double t = g.cos_lat0 * up - g.sin_lat0 * n;
    double w = g.sin_lat0 * up + g.cos_lat0 * n;
    double u = g.cos_lon0 * t - g.sin_lon0 * e;
    double v = g.sin_lon0 * t + g.cos_lon0 * e;
    x = g.x0 + u;
    y = g.y0 + v;
    z = g.z0 + w;
}

// uvw2enu
__device__ __forceinline__ void vs_uvw2enu(const VsGeoParams& g, double u, double v, double w, double& e, double& n,
                                           double& up) {
    double t = g.cos_lon0 * u + g.sin_lon0 * v;
    e = -g.sin_lon0 * u + g.cos_lon0 * v;
    up = g.cos_lat0 * t + g.sin_lat0 * w;
    n = -g.sin_lat0 * t + g.cos_lat0 * w;
}

// ---------------------------------------------------------------------------------------------------------
// PROJ etmerc
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double vs_gatg(const double* p, double B) {
    double h = 0.0, h1, h2 = 0.0;
    double cos_2B = 2.0 * cos(2.0 * B);
    h1 = p[VS_ETMERC_ORDER - 1];
#pragma unroll
    for (int k = VS_ETMERC_ORDER - 2; k >= 0; --k) {
        h = -h2 + cos_2B * h1 + p[k];
        h2 = h1;
        h1 = h;
    }
    return B + h * sin(2.0 * B);
}

__device__ __forceinline__ void vs_clenS(const double* a, double arg_r, double arg_i, double& R, double& I) {
    double sin_arg_r, cos_arg_r;
    sincos(arg_r, &sin_arg_r, &cos_arg_r);
    double sinh_arg_i = sinh(arg_i);
    double cosh_arg_i = cosh(arg_i);
    double r = 2.0 * cos_arg_r * cosh_arg_i;
    double i = -2.0 * sin_arg_r * sinh_arg_i;
    double hi1 = 0.0, hr1 = 0.0, hi = 0.0, hr = a[VS_ETMERC_ORDER - 1], hr2, hi2;
#pragma unroll
    for (int k = VS_ETMERC_ORDER - 2; k >= 0; --k) {
        hr2 = hr1;
        hi2 = hi1;
        hr1 = hr;
        hi1 = hi;
        hr = -hr2 + r * hr1 - i * hi1 + a[k];
        hi = -hi2 + i * hr1 + r * hi1;
    }
    r = sin_arg_r * cosh_arg_i;
    i = cos_arg_r * sinh_arg_i;
    R = r * hr - i * hi;
    I = r * hi + i * hr;
}

__device__ __forceinline__ double vs_asinhy(double x) {
    double y = fabs(x);
    y = log1p(y * (1.0 + y / (hypot(1.0, y) + 1.0)));
    return x < 0.0 ? -y : y;
}

__device__ __forceinline__ void vs_utm_forward(const VsEllipsoidConsts& c, double lam0, double north_off, double lat_deg,
                                               double lon_deg, double& east, double& north) {
    double Cn = c.dg2rad * lat_deg;
    double Ce = c.dg2rad * lon_deg - lam0;
    Cn = vs_gatg(c.cbg, Cn);
    double sin_Cn, cos_Cn, sin_Ce, cos_Ce;
    sincos(Cn, &sin_Cn, &cos_Cn);
    sincos(Ce, &sin_Ce, &cos_Ce);
    Cn = atan2(sin_Cn, cos_Ce * cos_Cn);
    Ce = atan2(sin_Ce * cos_Cn, hypot(sin_Cn, cos_Cn * cos_Ce));
    Ce = vs_asinhy(tan(Ce));
    double dCn, dCe;
    vs_clenS(c.gtu, 2.0 * Cn, 2.0 * Ce, dCn, dCe);
    Cn += dCn;
    Ce += dCe;
    if (fabs(Ce) <= 2.623395162778) {
        double y = c.Qn * Cn + c.Zb;
        double x = c.Qn * Ce;
        east = c.proj_a * x + 500000.0;
        north = c.proj_a * y + north_off;
    } else {
        east = north = CUDART_INF;
    }
}

__device__ __forceinline__ void vs_utm_inverse(const VsEllipsoidConsts& c, double lam0, double north_off, double east,
                                               double north, double& lat_deg, double& lon_deg) {
    double ra = 1.0 / c.proj_a;
    double Ce = (east - 500000.0) * ra;
    double Cn = (north - north_off) * ra;
    Cn = (Cn - c.Zb) / c.Qn;
    Ce = Ce / c.Qn;
    if (fabs(Ce) <= 2.623395162778) {
        double dCn, dCe;
        vs_clenS(c.utg, 2.0 * Cn, 2.0 * Ce, dCn, dCe);
        Cn += dCn;
        Ce += dCe;
        Ce = atan(sinh(Ce));
        double sin_Cn, cos_Cn, sin_Ce, cos_Ce;
        sincos(Cn, &sin_Cn, &cos_Cn);
        sincos(Ce, &sin_Ce, &cos_Ce);
        Ce = atan2(sin_Ce, cos_Ce * cos_Cn);
        Cn = atan2(sin_Cn * cos_Ce, hypot(sin_Ce, cos_Ce * cos_Cn));
        lat_deg = vs_gatg(c.cgb, Cn) * c.rad2dg;
        lon_deg = (Ce + lam0) * c.rad2dg;
    } else {
        lat_deg = lon_deg = CUDART_INF;
    }
}

// aggregate_2p5d_util.py:96-98 for one point
__device__ __forceinline__ void vs_enu_to_utm_exact(const VsEllipsoidConsts& c, const VsGeoParams& g, double e, double n,
                                                    double up, double& east, double& north, double& alt) {
    double x, y, z, lat, lon;
    vs_enu2ecef(g, e, n, up, x, y, z);
    vs_ecef2geodetic(c, x, y, z, lat, lon, alt);
    vs_utm_forward(c, g.lam0, g.north_off, lat, lon, east, north);
}
