#!/usr/bin/env python
"""Independent 50-digit pin for oracle/geodesy.py (and, through it, for the CUDA chain) -> geodesy_mp_golden.npz.

The reference calls pymap3d.enu2geodetic / geodetic2enu (lib/latlonalt_enu_converter.py:36-45) and
pyproj.Proj(proj='utm', ...) (lib/latlon_utm_converter.py:37-52, 61-62).  Neither package is installable here, so
`oracle/geodesy.py` restates their published algorithms (You's closed form, Poder/Engsager's 6th-order series).
This script computes the SAME MAPS from their mathematical definitions, with nothing shared with those algorithms,
in mpmath at 50 significant digits:

  * transverse Mercator (Gauss-Krueger): the conformal map that is true to scale on the central meridian, i.e. the
    analytic continuation of "meridian arc length as a function of isometric latitude" to complex isometric
    coordinates:   psi(phi) = asinh(tan phi) - e atanh(e sin phi)          (isometric latitude)
                   M(phi)   = a (1 - e^2) int_0^phi (1 - e^2 sin^2 t)^(-3/2) dt   (meridian arc)
                   N + iE   = k0 * M(phi_c),  phi_c complex with psi(phi_c) = psi(phi) + i (lambda - lambda0)
    phi_c by Newton iteration in complex arithmetic, M(phi_c) by tanh-sinh quadrature along the straight path
    (the integrand is analytic there).  No series in the third flattening, no Clenshaw sums.
  * ECEF -> geodetic: Newton iteration on the latitude equation to 50 digits (no closed form).
  * ENU <-> ECEF: the rotation by (lat0, lon0) about the exactly computed ECEF origin.

Outputs are rounded once to float64 (plus a float64 residual), so the frozen vectors are exact to ~1e-25 m.
Inverse directions are pinned through the exact forward map applied to the implementation's output
(tests/test_geodesy_pin.py): x -> inverse_impl(x) -> exact forward == x.

Usage:  python tests/golden/make_geodesy_mp.py          (about two minutes on 8 cores)
"""
import multiprocessing as mp_pool
import os
import sys

import numpy as np
from mpmath import mp, mpf, mpc

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

mp.dps = 50

A = mpf(6378137)
RF = mpf('298.257223563')
F = 1 / RF
E2 = F * (2 - F)
E1 = mp.sqrt(E2)
B = A * (1 - F)
K0 = mpf('0.9996')


def _mpf(x):
    """exact binary value of a float64"""
    return mpf(float(x))


# ---- transverse Mercator from its definition --------------------------------------------------------------
def _psi(phi):
    s = mp.sin(phi)
    return mp.asinh(mp.tan(phi)) - E1 * mp.atanh(E1 * s)


def _dpsi(phi):
    s = mp.sin(phi)
    return (1 - E2) / ((1 - E2 * s * s) * mp.cos(phi))


def _arc(phi_c):
    g = lambda t: (1 - E2 * mp.sin(t) ** 2) ** mpf('-1.5')     # noqa: E731
    return A * (1 - E2) * mp.quad(g, [0, phi_c])


def tm_forward_exact(lat_deg, lon_deg, lam0_rad, south):
    """(E, N) as mpf for one point; lam0_rad an mpf."""
    phi = mp.radians(_mpf(lat_deg))
    dlam = mp.radians(_mpf(lon_deg)) - lam0_rad
    zeta = mpc(_psi(phi), dlam)
    # spherical start: gd(zeta) = 2 atan(tanh(zeta / 2)); then Newton on psi(phi_c) = zeta
    pc = 2 * mp.atan(mp.tanh(zeta / 2))
    for _ in range(60):
        step = (_psi(pc) - zeta) / _dpsi(pc)
        pc = pc - step
        if abs(step) < mpf(10) ** (-(mp.dps - 4)):
            break
    else:
        raise RuntimeError('Newton did not converge')
    w = K0 * _arc(pc)
    north = w.real + (mpf(10000000) if south else 0)
    east = w.imag + 500000
    return east, north


def utm_lam0(zone):
    return mp.radians(mpf((zone - 1) * 6 - 180 + 3))


# ---- ellipsoid <-> ECEF <-> ENU ------------------------------------------------------------------------------
def geodetic2ecef_exact(lat_deg, lon_deg, h):
    phi, lam = mp.radians(_mpf(lat_deg)), mp.radians(_mpf(lon_deg))
    s, c = mp.sin(phi), mp.cos(phi)
    n = A / mp.sqrt(1 - E2 * s * s)
    hh = _mpf(h)
    return (n + hh) * c * mp.cos(lam), (n + hh) * c * mp.sin(lam), (n * (1 - E2) + hh) * s


def ecef2geodetic_exact(x, y, z):
    p = mp.sqrt(x * x + y * y)
    lam = mp.atan2(y, x)
    phi = mp.atan2(z, p * (1 - E2))
    for _ in range(200):      # f(phi) = p tan(phi) - z - e^2 N(phi) sin(phi) = 0, Newton
        s, c = mp.sin(phi), mp.cos(phi)
        w2 = 1 - E2 * s * s
        n = A / mp.sqrt(w2)
        fval = p * s / c - z - E2 * n * s
        dn = A * E2 * s * c / (w2 * mp.sqrt(w2))
        dval = p / (c * c) - E2 * (dn * s + n * c)
        step = fval / dval
        phi -= step
        if abs(step) < mpf(10) ** (-(mp.dps - 4)):
            break
    else:
        raise RuntimeError('latitude iteration did not converge')
    s, c = mp.sin(phi), mp.cos(phi)
    n = A / mp.sqrt(1 - E2 * s * s)
    # numerically safe altitude: p cos + z sin - N (1 - e^2 sin^2)
    h = p * c + z * s - n * (1 - E2 * s * s)
    return mp.degrees(phi), mp.degrees(lam), h


def enu2ecef_exact(e, n, u, lat0, lon0, h0):
    x0, y0, z0 = geodetic2ecef_exact(lat0, lon0, h0)
    p0, l0 = mp.radians(_mpf(lat0)), mp.radians(_mpf(lon0))
    sp, cp, sl, cl = mp.sin(p0), mp.cos(p0), mp.sin(l0), mp.cos(l0)
    e, n, u = _mpf(e), _mpf(n), _mpf(u)
    dx = -sl * e - sp * cl * n + cp * cl * u
    dy = cl * e - sp * sl * n + cp * sl * u
    dz = cp * n + sp * u
    return x0 + dx, y0 + dy, z0 + dz


def ecef2enu_exact(x, y, z, lat0, lon0, h0):
    x0, y0, z0 = geodetic2ecef_exact(lat0, lon0, h0)
    p0, l0 = mp.radians(_mpf(lat0)), mp.radians(_mpf(lon0))
    sp, cp, sl, cl = mp.sin(p0), mp.cos(p0), mp.sin(l0), mp.cos(l0)
    dx, dy, dz = x - x0, y - y0, z - z0
    e = -sl * dx + cl * dy
    n = -sp * cl * dx - sp * sl * dy + cp * dz
    u = cp * cl * dx + cp * sl * dy + sp * dz
    return e, n, u


def split(v):
    """mpf -> (float64 hi, float64 lo)"""
    hi = float(v)
    return hi, float(v - mpf(hi))


# ---- work items (module-level functions: picklable for the pool) ------------------------------------------------
def _job_enu_to_utm(args):
    """aggregate_2p5d_util.py:96-98 for one point, exactly: ENU -> (lat, lon, alt) -> (E, N)."""
    mp.dps = 50
    e, n, u, lat0, lon0, h0, zone, south = args
    x, y, z = enu2ecef_exact(e, n, u, lat0, lon0, h0)
    lat, lon, alt = ecef2geodetic_exact(x, y, z)
    # the exact composite map (the float64 rounding of lat/lon between the two library calls is part of the
    # implementations' noise, not of the map): continue in 50 digits
    phi = mp.radians(lat)
    dlam = mp.radians(lon) - utm_lam0(zone)
    zeta = mpc(_psi(phi), dlam)
    pc = 2 * mp.atan(mp.tanh(zeta / 2))
    for _ in range(60):
        step = (_psi(pc) - zeta) / _dpsi(pc)
        pc -= step
        if abs(step) < mpf(10) ** (-(mp.dps - 4)):
            break
    w = K0 * _arc(pc)
    north = w.real + (mpf(10000000) if south else 0)
    east = w.imag + 500000
    return split(lat) + split(lon) + split(alt) + split(east) + split(north)


def _job_tm_forward(args):
    mp.dps = 50
    lat, lon, zone, south = args
    e, n = tm_forward_exact(lat, lon, utm_lam0(zone), south)
    return split(e) + split(n)


def _job_geodetic_to_enu(args):
    mp.dps = 50
    lat, lon, h, lat0, lon0, h0 = args
    x, y, z = geodetic2ecef_exact(lat, lon, h)
    e, n, u = ecef2enu_exact(x, y, z, lat0, lon0, h0)
    return split(e) + split(n) + split(u)


def main():
    from oracle import geodesy
    from vissatsatellitestereo_b200 import synthetic as S

    rng = np.random.default_rng(20261017)
    out = {}
    with mp_pool.get_context('fork').Pool(os.cpu_count()) as pool:
        # ---- (1) the five benchmark AOIs: ENU points over the grid footprint and the altitude range -> (lat, lon,
        #          alt, E, N).  2 048 points each (+ the 8 corners of the box), 10 280 in total.
        aoi_rows, jobs = [], []
        for name in ('C1', 'C2', 'C3', 'C4', 'C5'):
            cfg = S.CONFIGS[name]
            aoi = S.make_aoi(cfg, geodesy)
            lat0, lon0, h0 = geodesy.enu_origin_from_aoi(aoi)
            half_e = 0.5 * cfg.e_size * cfg.res + 150.0
            half_n = 0.5 * cfg.n_size * cfg.res + 150.0
            n_pts = 2048
            e = rng.uniform(-half_e, half_e, n_pts)
            n = rng.uniform(-half_n, half_n, n_pts)
            u = rng.uniform(-60.0, 260.0, n_pts)           # alt0 = alt_min = -30: altitudes -90 .. 230 m
            corners = np.array([[sx * half_e, sy * half_n, uz] for sx in (-1, 1) for sy in (-1, 1) for uz in (-60.0, 260.0)])
            e = np.concatenate([e, corners[:, 0]])
            n = np.concatenate([n, corners[:, 1]])
            u = np.concatenate([u, corners[:, 2]])
            south = aoi['hemisphere'] != 'N'
            for i in range(e.size):
                jobs.append((e[i], n[i], u[i], lat0, lon0, h0, aoi['zone_number'], south))
                aoi_rows.append((e[i], n[i], u[i], lat0, lon0, h0, aoi['zone_number'], 1.0 if south else 0.0))
        res = pool.map(_job_enu_to_utm, jobs, chunksize=16)
        out['aoi_in'] = np.array(aoi_rows, dtype=np.float64)            # e n u lat0 lon0 h0 zone south
        out['aoi_out'] = np.array(res, dtype=np.float64)                # (lat, lon, alt, E, N) x (hi, lo)
        print('AOI points:', out['aoi_in'].shape[0], flush=True)

        # ---- (2) world-wide UTM forward: all zones' geometry, both hemispheres, |lon - CM| <= 3.5 deg
        n_w = 1536
        lat = np.concatenate([rng.uniform(-80.0, 84.0, n_w - 6), [0.0, 1e-9, -1e-9, 84.0, -80.0, 45.0]])
        zone = rng.integers(1, 61, n_w)
        dlon = np.concatenate([rng.uniform(-3.5, 3.5, n_w - 6), [0.0, 3.0, -3.0, 0.0, 0.0, 3.5]])
        lon = (zone - 1) * 6.0 - 180.0 + 3.0 + dlon
        south = lat < 0
        jobs = [(lat[i], lon[i], int(zone[i]), bool(south[i])) for i in range(n_w)]
        res = pool.map(_job_tm_forward, jobs, chunksize=16)
        out['utm_in'] = np.stack([lat, lon, zone.astype(np.float64), south.astype(np.float64)], axis=1)
        out['utm_out'] = np.array(res, dtype=np.float64)                # (E, N) x (hi, lo)
        print('UTM points:', n_w, flush=True)

        # ---- (3) world-wide geodetic -> ENU about random origins, points within ~5 km and +-500 m of the origin
        n_g = 1536
        lat0 = rng.uniform(-85.0, 85.0, n_g)
        lon0 = rng.uniform(-180.0, 180.0, n_g)
        h0 = rng.uniform(-100.0, 3000.0, n_g)
        lat = lat0 + rng.uniform(-0.04, 0.04, n_g)
        lon = lon0 + rng.uniform(-0.04, 0.04, n_g)
        h = h0 + rng.uniform(-500.0, 500.0, n_g)
        jobs = [(lat[i], lon[i], h[i], lat0[i], lon0[i], h0[i]) for i in range(n_g)]
        res = pool.map(_job_geodetic_to_enu, jobs, chunksize=32)
        out['enu_in'] = np.stack([lat, lon, h, lat0, lon0, h0], axis=1)
        out['enu_out'] = np.array(res, dtype=np.float64)                # (e, n, u) x (hi, lo)
        print('ENU points:', n_g, flush=True)

    # published known answers the map must reproduce (sanity of THIS script, asserted here and in the test):
    # PROJ docs: echo 12 56 | proj +proj=utm +zone=32  ->  687071.44  6210141.33
    e, n = tm_forward_exact(56.0, 12.0, utm_lam0(32), False)
    assert abs(e - mpf('687071.44')) < mpf('0.006') and abs(n - mpf('6210141.33')) < mpf('0.006'), (e, n)
    # meridian quadrant of WGS84: 10 001 965.729 m (Karney 2011); k0 * quadrant at the pole on the central meridian
    q = _arc(mp.pi / 2)
    assert abs(q - mpf('10001965.729')) < mpf('0.001'), q
    out['quadrant'] = np.array(split(q))
    np.savez_compressed(os.path.join(HERE, 'geodesy_mp_golden.npz'), **out)
    print('wrote', os.path.join(HERE, 'geodesy_mp_golden.npz'))


if __name__ == '__main__':
    main()
