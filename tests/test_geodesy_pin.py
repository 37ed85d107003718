"""Independent pin of the third-party geodesy (pymap3d / PROJ utm) behind the reference's converters.

tests/golden/geodesy_mp_golden.npz holds the maps evaluated from their mathematical definitions in mpmath at 50
digits (tests/golden/make_geodesy_mp.py; nothing shared with the series the oracle and the kernels use).
Both `oracle/geodesy.py` (CPU, here) and the CUDA chain (`-m gpu`, below) must agree with it to float64 noise:

    AOI set   (10 280 points over the five benchmark AOIs):   <= 5e-9 m in E, N and altitude
    world set (1 536 points, all zones, both hemispheres, |lon - CM| <= 3.5 deg; 1 536 ENU origins): <= 1e-8 m

Reference call sites: lib/latlonalt_enu_converter.py:36-45, lib/latlon_utm_converter.py:37-52,61-62.
"""
import os

import numpy as np
import pytest

from oracle import geodesy

HERE = os.path.dirname(os.path.abspath(__file__))
TOL_AOI = 5e-9       # metres
TOL_WORLD = 1e-8     # metres
M_PER_DEG = 111.32e3


@pytest.fixture(scope='module')
def mp_golden():
    return np.load(os.path.join(HERE, 'golden', 'geodesy_mp_golden.npz'))


def _hi_lo(a, k):
    return a[:, 2 * k], a[:, 2 * k + 1]


def _err(got, gold, k):
    hi, lo = _hi_lo(gold, k)
    return np.abs((np.asarray(got).reshape(-1) - hi) - lo)


def _aoi_groups(z):
    a_in = z['aoi_in']
    for key in np.unique(a_in[:, 3:8], axis=0):
        yield key, np.all(a_in[:, 3:8] == key, axis=1)


def _check_aoi(z, enu_to_latlonalt, utm_forward, utm_inverse, latlonalt_to_enu, tol):
    a_in, a_out = z['aoi_in'], z['aoi_out']
    worst = {}
    n_groups = 0
    for key, m in _aoi_groups(z):
        n_groups += 1
        lat0, lon0, h0, zone, south = key[0], key[1], key[2], int(key[3]), bool(key[4])
        e, n, u = a_in[m, 0:1], a_in[m, 1:2], a_in[m, 2:3]
        lat, lon, alt = enu_to_latlonalt(e, n, u, lat0, lon0, h0)
        E, N = utm_forward(lat, lon, zone, south)
        g = a_out[m]
        coslat = np.cos(np.radians(g[:, 0]))
        errs = {'lat_m': _err(lat, g, 0) * M_PER_DEG, 'lon_m': _err(lon, g, 1) * M_PER_DEG * coslat,
                'alt_m': _err(alt, g, 2), 'E_m': _err(E, g, 3), 'N_m': _err(N, g, 4)}
        # inverse directions through the exact forward map: exact (E, N) -> lat/lon must give the exact lat/lon
        la2, lo2 = utm_inverse(g[:, 6:7], g[:, 8:9], zone, south)
        errs['inv_lat_m'] = _err(la2, g, 0) * M_PER_DEG
        errs['inv_lon_m'] = _err(lo2, g, 1) * M_PER_DEG * coslat
        e2, n2, u2 = latlonalt_to_enu(g[:, 0:1], g[:, 2:3], g[:, 4:5], lat0, lon0, h0)
        errs['inv_e_m'] = np.abs(np.asarray(e2).reshape(-1) - e[:, 0])
        errs['inv_n_m'] = np.abs(np.asarray(n2).reshape(-1) - n[:, 0])
        errs['inv_u_m'] = np.abs(np.asarray(u2).reshape(-1) - u[:, 0])
        for k, v in errs.items():
            worst[k] = max(worst.get(k, 0.0), float(v.max()))
    assert n_groups == 5
    print('AOI set, max |impl - exact| (m):', {k: '%.2e' % v for k, v in worst.items()})
    for k, v in worst.items():
        assert v <= tol, (k, v)
    return worst


def _check_world(z, utm_forward, utm_inverse, latlonalt_to_enu, enu_to_latlonalt, tol):
    u_in, u_out = z['utm_in'], z['utm_out']
    worst = {'E_m': 0.0, 'N_m': 0.0, 'inv_lat_m': 0.0, 'inv_lon_m': 0.0}
    for zone in np.unique(u_in[:, 2]):
        for south in (0.0, 1.0):
            m = (u_in[:, 2] == zone) & (u_in[:, 3] == south)
            if not m.any():
                continue
            lat, lon = u_in[m, 0:1], u_in[m, 1:2]
            E, N = utm_forward(lat, lon, int(zone), bool(south))
            worst['E_m'] = max(worst['E_m'], float(_err(E, u_out[m], 0).max()))
            worst['N_m'] = max(worst['N_m'], float(_err(N, u_out[m], 1).max()))
            la2, lo2 = utm_inverse(u_out[m, 0:1], u_out[m, 2:3], int(zone), bool(south))
            worst['inv_lat_m'] = max(worst['inv_lat_m'], float(np.abs(la2 - lat).max() * M_PER_DEG))
            worst['inv_lon_m'] = max(worst['inv_lon_m'],
                                     float((np.abs(lo2 - lon) * np.cos(np.radians(lat))).max() * M_PER_DEG))
    g_in, g_out = z['enu_in'], z['enu_out']
    we = np.zeros(3)
    wi = np.zeros(3)
    for i in range(g_in.shape[0]):
        lat, lon, h, lat0, lon0, h0 = g_in[i]
        one = lambda x: np.array([[x]])                                   # noqa: E731
        r = latlonalt_to_enu(one(lat), one(lon), one(h), lat0, lon0, h0)
        we = np.maximum(we, [abs(float(np.asarray(r[j]).reshape(-1)[0]) - g_out[i, 2 * j] - g_out[i, 2 * j + 1]) for j in range(3)])
        q = enu_to_latlonalt(one(g_out[i, 0]), one(g_out[i, 2]), one(g_out[i, 4]), lat0, lon0, h0)
        wi = np.maximum(wi, [abs(float(np.asarray(q[0]).reshape(-1)[0]) - lat) * M_PER_DEG,
                             abs(float(np.asarray(q[1]).reshape(-1)[0]) - lon) * M_PER_DEG * np.cos(np.radians(lat)),
                             abs(float(np.asarray(q[2]).reshape(-1)[0]) - h)])
    worst.update({'enu_e_m': we[0], 'enu_n_m': we[1], 'enu_u_m': we[2],
                  'inv_enu_lat_m': wi[0], 'inv_enu_lon_m': wi[1], 'inv_enu_alt_m': wi[2]})
    print('world set, max |impl - exact| (m):', {k: '%.2e' % v for k, v in worst.items()})
    for k, v in worst.items():
        assert v <= tol, (k, v)
    return worst


# ------------------------------------------------------------------------------------------------ the generator
def test_exact_map_known_answers(mp_golden):
    """The 50-digit map itself against published numbers (PROJ documentation example, WGS84 meridian quadrant)."""
    from mpmath import mp, mpf
    import importlib.util
    spec = importlib.util.spec_from_file_location('make_geodesy_mp', os.path.join(HERE, 'golden', 'make_geodesy_mp.py'))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    mp.dps = 50
    e, n = G.tm_forward_exact(56.0, 12.0, G.utm_lam0(32), False)          # echo 12 56 | proj +proj=utm +zone=32
    assert abs(e - mpf('687071.44')) < mpf('0.006') and abs(n - mpf('6210141.33')) < mpf('0.006')
    assert abs(G._arc(mp.pi / 2) - mpf('10001965.729')) < mpf('0.001')    # Karney 2011, WGS84 quadrant
    assert abs(float(mp_golden['quadrant'][0]) - 10001965.729) < 1e-3
    # equator / central meridian, and the frozen file agrees with a fresh evaluation of a few of its rows
    e, n = G.tm_forward_exact(0.0, -57.0, G.utm_lam0(21), True)
    assert abs(e - 500000) < mpf('1e-40') and abs(n - 10000000) < mpf('1e-40')
    a_in, a_out = mp_golden['aoi_in'], mp_golden['aoi_out']
    for i in (0, 777, a_in.shape[0] - 1):
        r = G._job_enu_to_utm(tuple(a_in[i, :6]) + (int(a_in[i, 6]), bool(a_in[i, 7])))
        assert np.array_equal(np.array(r), a_out[i])


# ------------------------------------------------------------------------------------------------ the oracle (CPU)
def test_oracle_geodesy_vs_exact_aoi(mp_golden):
    _check_aoi(mp_golden, geodesy.enu_to_latlonalt, geodesy.utm_forward, geodesy.utm_inverse,
               geodesy.latlonalt_to_enu, TOL_AOI)


def test_oracle_geodesy_vs_exact_world(mp_golden):
    _check_world(mp_golden, geodesy.utm_forward, geodesy.utm_inverse, geodesy.latlonalt_to_enu,
                 geodesy.enu_to_latlonalt, TOL_WORLD)


# ------------------------------------------------------------------------------------------------ the CUDA chain
def _gpu_fns():
    from vissatsatellitestereo_b200 import engine
    from vissatsatellitestereo_b200.lib import latlonalt_enu_converter as C1
    from vissatsatellitestereo_b200.lib._geo_common import run2
    engine.require_cuda()
    fwd = lambda lat, lon, zone, south: run2('vs_geodetic_to_utm', lat, lon, int(zone), 1 if south else 0)   # noqa: E731
    inv = lambda e, n, zone, south: run2('vs_utm_to_geodetic', e, n, int(zone), 1 if south else 0)            # noqa: E731
    return C1.enu_to_latlonalt, fwd, inv, C1.latlonalt_to_enu


@pytest.mark.gpu
def test_cuda_chain_vs_exact_aoi(mp_golden):
    e2g, fwd, inv, g2e = _gpu_fns()
    _check_aoi(mp_golden, e2g, fwd, inv, g2e, TOL_AOI)


@pytest.mark.gpu
def test_cuda_chain_vs_exact_world(mp_golden):
    e2g, fwd, inv, g2e = _gpu_fns()
    _check_world(mp_golden, fwd, inv, g2e, e2g, TOL_WORLD)


@pytest.mark.gpu
def test_cuda_enu_to_utm_vs_exact(mp_golden):
    """vs_enu_to_utm (aggregate_2p5d_util.py:96-98 in one pass, the chain K1's polynomial is fitted to and
    validated against) within 5e-9 m of the exact composite map on every AOI point."""
    import ctypes as C
    import torch
    from vissatsatellitestereo_b200 import engine, _native
    engine.require_cuda()
    ctx, dev = engine.default_context()
    a_in, a_out = mp_golden['aoi_in'], mp_golden['aoi_out']
    worst = np.zeros(3)
    for key, m in _aoi_groups(mp_golden):
        t = [torch.from_numpy(np.ascontiguousarray(a_in[m, k])).to(dev) for k in range(3)]
        o = [torch.empty_like(t[0]) for _ in range(3)]
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        p = lambda x: C.c_void_p(x.data_ptr())                             # noqa: E731
        _native.check(_native.lib.vs_enu_to_utm(ctx.handle, p(t[0]), p(t[1]), p(t[2]), t[0].numel(), float(key[0]),
                                                float(key[1]), float(key[2]), int(key[3]), int(key[4]),
                                                p(o[0]), p(o[1]), p(o[2]), st), 'vs_enu_to_utm')
        g = a_out[m]
        E, N, A = (x.cpu().numpy() for x in o)
        worst = np.maximum(worst, [_err(E, g, 3).max(), _err(N, g, 4).max(), _err(A, g, 2).max()])
    print('vs_enu_to_utm, max |GPU - exact| (m): E %.2e  N %.2e  alt %.2e' % tuple(worst))
    assert worst.max() <= TOL_AOI


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['C1', 'C2', 'C3', 'C4', 'C5'])
def test_k1_polynomial_vs_exact(mp_golden, name):
    """The per-AOI polynomial K1 evaluates (fitted to the device chain) against the EXACT map: fractional row/col
    within 1e-7 cell (the ambiguity threshold the cell-index audit uses) and altitude within 3e-8 m."""
    import torch
    from vissatsatellitestereo_b200 import engine, synthetic as S
    engine.require_cuda()
    cfg = S.CONFIGS[name]
    aoi = S.make_aoi(cfg, geodesy)
    eng = engine.DsmEngine(aoi, cfg.res, cfg.res, device=0)
    assert eng.fit['degree'] >= 3
    lat0, lon0, h0 = geodesy.enu_origin_from_aoi(aoi)
    a_in, a_out = mp_golden['aoi_in'], mp_golden['aoi_out']
    m = np.all(a_in[:, 3:6] == np.array([lat0, lon0, h0]), axis=1)
    assert m.sum() >= 2048
    c, h = np.array(eng.fit['box_center']), np.array(eng.fit['box_half'])
    pts = a_in[m, :3]
    inside = np.all(np.abs((pts - c) / h) <= 1.0, axis=1)
    assert inside.sum() >= 300, inside.sum()
    colf, rowf, alt = eng.eval_poly(pts[inside])
    g = a_out[m][inside]
    E = g[:, 6] + g[:, 7]
    N = g[:, 8] + g[:, 9]
    want_col = (E - aoi['ul_easting']) / cfg.res
    want_row = (aoi['ul_northing'] - N) / cfg.res
    ec = np.abs(colf - want_col).max()
    er = np.abs(rowf - want_row).max()
    ea = np.abs(alt - g[:, 4]).max()
    print('{}: polynomial (degree {}, mixed {}) vs exact: col {:.2e} cell, row {:.2e} cell, alt {:.2e} m'.format(
        name, eng.fit['degree'], eng.fit['mixed'], ec, er, ea))
    assert ec <= 1e-7 and er <= 1e-7 and ea <= 3e-8
    eng.close()
