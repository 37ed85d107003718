"""GPU parity: the CUDA path (through the C ABI) against the oracle and the reference-run golden vectors.

Bars (BASELINE.json north_star): cell indices, occupancy masks and median selection bit-exact; heights within
1e-3 m.  Where float64 rounding noise of the CPU chain itself (~1e-9 m) can legitimately move a point across a
cell edge or flip the strict `abs(x-med) > mad` test, the tests audit it instead of hiding it.
"""
import json
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import geodesy, pipeline as op

pytestmark = pytest.mark.gpu

HEIGHT_TOL = 1e-3   # metres, north_star


def _eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope='module')
def lanes():
    return op.detect_cv2_simd_lanes()


@pytest.fixture(scope='module')
def eng_mod():
    from vissatsatellitestereo_b200 import engine
    engine.require_cuda()
    return engine


# ------------------------------------------------------------------------------------------------ converters
def test_exact_converters_vs_oracle(eng_mod):
    from vissatsatellitestereo_b200.lib import latlonalt_enu_converter as C1, latlon_utm_converter as C2
    rng = np.random.default_rng(1)
    n = 20000
    lat0, lon0, alt0 = -34.4899, -58.5856, -30.0
    e = rng.uniform(-3000, 3000, (n, 1))
    nn = rng.uniform(-3000, 3000, (n, 1))
    u = rng.uniform(-100, 400, (n, 1))
    lat, lon, alt = C1.enu_to_latlonalt(e, nn, u, lat0, lon0, alt0)
    olat, olon, oalt = geodesy.enu_to_latlonalt(e, nn, u, lat0, lon0, alt0)
    assert lat.shape == (n, 1)
    # 1e-8 m == 9e-14 deg; float64 chain noise is ~1e-9 m
    assert np.abs(lat - olat).max() * 111e3 < 2e-8 and np.abs(lon - olon).max() * 111e3 < 2e-8
    assert np.abs(alt - oalt).max() < 2e-8
    east, north = C2.latlon_to_eastnorh(lat, lon)
    oe, on = geodesy.latlon_to_eastnorh(olat, olon)
    assert np.abs(east - oe).max() < 3e-8 and np.abs(north - on).max() < 3e-8
    la2, lo2 = C2.eastnorth_to_latlon(east, north, 21, 'S')
    ola2, olo2 = geodesy.eastnorth_to_latlon(oe, on, 21, 'S')
    assert np.abs(la2 - ola2).max() * 111e3 < 3e-8 and np.abs(lo2 - olo2).max() * 111e3 < 3e-8
    e2, n2, u2 = C1.latlonalt_to_enu(lat, lon, alt, lat0, lon0, alt0)
    assert np.abs(e2 - e).max() < 3e-8 and np.abs(n2 - nn).max() < 3e-8 and np.abs(u2 - u).max() < 3e-8
    # scalars in -> scalars out (lib/latlonalt_enu_converter.py:49-58 sample)
    es, ns, us = C1.latlonalt_to_enu(-34.450, -58.579, 20.31, -34.448, -58.577, -30.0)
    assert isinstance(es, float)
    assert abs(es + 183.79014029) < 1e-6 and abs(ns + 221.86356704) < 1e-6 and abs(us - 50.30348255) < 1e-6
    # published known answers
    # (PROJ docs: `echo 12 56 | proj +proj=utm +zone=32`; utm's zone rule would pick 33 for lon = 12)
    from vissatsatellitestereo_b200.lib._geo_common import run2
    E, N = run2('vs_geodetic_to_utm', np.array([[56.0]]), np.array([[12.0]]), 32, 0)
    assert abs(E[0, 0] - 687071.44) < 0.006 and abs(N[0, 0] - 6210141.33) < 0.006
    assert C2.latlon_to_zone_number(56.0, 12.0) == 33 and C2.latlon_to_zone_number(60.0, 5.0) == 32
    # northern hemisphere + both of the reference's __main__ samples
    for s in (1, -1):
        la = np.array([[s * 47.9941214]])
        lo = np.array([[7.8509671]])
        E, N = C2.latlon_to_eastnorh(la, lo)
        oE, oN = geodesy.latlon_to_eastnorh(la, lo)
        assert abs(E[0, 0] - oE[0, 0]) < 1e-8 and abs(N[0, 0] - oN[0, 0]) < 1e-8


# ------------------------------------------------------------------------------------------------ proj_to_grid
@pytest.mark.parametrize('k', range(5))
def test_proj_to_grid_bit_exact_vs_reference_golden(golden, eng_mod, k):
    from vissatsatellitestereo_b200.lib.proj_to_grid import proj_to_grid
    pts = golden['ptg{}_points'.format(k)]
    xoff, yoff, xres, yres, xs, ys = golden['ptg{}_args'.format(k)]
    got = proj_to_grid(pts, xoff, yoff, xres, yres, int(xs), int(ys))
    assert got.dtype == np.float64
    assert _eq(got, golden['ptg{}_dsm'.format(k)])


def test_proj_to_grid_large_random_vs_oracle(eng_mod):
    from vissatsatellitestereo_b200.lib.proj_to_grid import proj_to_grid
    rng = np.random.default_rng(3)
    n, xs, ys = 400000, 301, 257
    pts = np.stack([1000 + rng.uniform(-5, xs * 0.3 + 5, n), 5000 - rng.uniform(-5, ys * 0.3 + 5, n),
                    rng.normal(20, 30, n)], 1)
    pts = pts[(pts[:, 0] < 1030) | (pts[:, 0] > 1045)]     # a NaN band wider than the 3x3 fill
    pts[rng.random(pts.shape[0]) < 0.01, 2] = np.nan
    got = proj_to_grid(pts, 1000.0, 5000.0, 0.3, 0.3, xs, ys)
    want = op.proj_to_grid_fast(pts, 1000.0, 5000.0, 0.3, 0.3, xs, ys)
    assert _eq(got, want)
    # empty input -> all-NaN grid
    assert np.isnan(proj_to_grid(np.zeros((0, 3)), 0.0, 0.0, 1.0, 1.0, 4, 3)).all()


# ------------------------------------------------------------------------------------------------ medianBlur
@pytest.mark.parametrize('shape', [(1, 1), (1, 9), (9, 1), (2, 2), (3, 3), (5, 17), (6, 18), (40, 33), (33, 70), (257, 131)])
def test_median3x3_bit_exact_vs_cv2(eng_mod, lanes, shape):
    from vissatsatellitestereo_b200 import synthetic as S
    eng = _tiny_engine(eng_mod, lanes)
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    for nan_frac in (0.0, 0.3, 0.8):
        img = rng.normal(size=shape).astype(np.float32)
        img[rng.random(shape) < nan_frac] = np.nan
        got = eng.median3x3(torch.from_numpy(img).cuda()).cpu().numpy()
        assert _eq(got, cv2.medianBlur(img, 3)), (shape, nan_frac)
    # row-band form (multi-GPU fusion): rows [r0, r1) from a buffer holding rows r0-1 .. r1
    if shape[0] >= 8 and shape[1] > 1:
        img = rng.normal(size=shape).astype(np.float32)
        img[rng.random(shape) < 0.3] = np.nan
        want = cv2.medianBlur(img, 3)
        r0, r1 = 3, shape[0] - 2
        band = torch.from_numpy(np.ascontiguousarray(img[r0 - 1:r1 + 1])).cuda()
        got = eng.median3x3(band, row_begin=r0, row_end=r1, in_row0=r0 - 1, h_total=shape[0]).cpu().numpy()
        assert _eq(got, want[r0:r1])


_TINY = {}


def _tiny_engine(eng_mod, lanes):
    if 'e' not in _TINY:
        from vissatsatellitestereo_b200 import synthetic as S
        cfg = S.scaled(S.CONFIGS['C1'], views=1, depth=64, grid=32)
        _TINY['e'] = eng_mod.DsmEngine(S.make_aoi(cfg, geodesy), cfg.res, cfg.res, simd_lanes=lanes)
    return _TINY['e']


# ------------------------------------------------------------------------------------------------ fusion alone
@pytest.mark.parametrize('V', [1, 2, 3, 4, 7, 8, 9, 16, 23, 33, 36, 37, 43, 44, 50, 52, 56, 57, 60, 61, 64, 65, 81, 88, 100, 104, 113, 120, 128, 129, 136, 200,
                               208, 257, 300, 400, 416, 512, 513, 700, 1025, 2048])
def test_fusion_bit_exact_vs_numpy(eng_mod, lanes, V):
    eng = _tiny_engine(eng_mod, lanes)
    rng = np.random.default_rng(V)
    H, W = (37, 53) if V <= 300 else (9, 41)
    cube = (30 + 5 * rng.normal(size=(V, H, W))).astype(np.float32)
    cube[rng.random(cube.shape) < 0.35] = np.nan
    cube[:, 0, 0] = np.nan
    cube[2:, 0, 1] = np.nan
    cube[:, 0, 2] = 7.25
    cube[:, 1, :] = np.round(cube[:, 1, :])          # many ties
    cube[:, 2, :] = np.where(rng.random((V, W)) < 0.5, 3.0, 4.0).astype(np.float32)
    # inputs that are NOT alike across the lanes of the multi-lane kernels (lane = view mod 4 / 8 / 32): views alternating
    # between two levels, levels by view mod 8, a ramp over the view index, one lane far away -- the split search of
    # k_fuse_large must take its bounded number of exchanges and finish with the bisection
    vi = np.arange(V, dtype=np.float32)[:, None]
    noise = rng.normal(size=(V, W)).astype(np.float32)
    cube[:, 3, :] = np.where(vi % 2 == 0, 10.0, 50.0) + noise
    cube[:, 4, :] = 5.0 * (vi % 8) + 0.1 * noise
    cube[:, 5, :] = 0.37 * vi + 0.01 * noise
    cube[:, 6, :] = np.where(vi % 8 == 3, -500.0, 20.0) + noise
    cube[:, 7, :] = np.where(vi % 32 < 16, 1.0, 2.0) + 0.001 * noise
    cube[rng.random(cube.shape) < 0.02] = np.nan
    want = op.fuse_dsms([cube[v].copy() for v in range(V)], blur=False)
    got = eng.fuse(torch.from_numpy(cube).cuda()).cpu().numpy()
    assert _eq(got, want)


@pytest.mark.parametrize('case', ['c1', 'c5', 'c3'])
def test_fusion_of_reference_per_view_dsms_bit_exact(golden, eng_mod, lanes, case):
    """Stage C + final blur fed with the reference's own per-view DSMs == the reference's fused DSM."""
    aoi = json.loads(str(golden[case + '_aoi']))
    res = float(golden[case + '_res'])
    eng = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)
    pv = torch.from_numpy(golden[case + '_per_view']).cuda()
    got = eng.fuse_and_blur(pv).cpu().numpy()
    assert _eq(got, golden[case + '_fused'])
    assert eng.last_nan_count() == int(np.isnan(golden[case + '_fused']).sum())


# ------------------------------------------------------------------------------------------------ stage A + B
def _ambiguous_cells(depth, M, aoi, res, eps_cells):
    """Cells that a point within eps of a cell edge could move into / out of (oracle-side audit)."""
    _, pts = op.unproject_depth(depth, M)
    utm = op.enu_points_to_utm(pts, aoi)
    e_size, n_size = op.grid_shape(aoi, res, res)
    rowf = (aoi['ul_northing'] - utm[:, 1]) / res
    colf = (utm[:, 0] - aoi['ul_easting']) / res
    fr = rowf - np.floor(rowf)
    fc = colf - np.floor(colf)
    amb = (fr < eps_cells) | (fr > 1 - eps_cells) | (fc < eps_cells) | (fc > 1 - eps_cells)
    mask = np.zeros((n_size, e_size), dtype=bool)
    r = np.floor(rowf[amb]).astype(int)
    c = np.floor(colf[amb]).astype(int)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            rr, cc = r + dr, c + dc
            ok = (rr >= 0) & (rr < n_size) & (cc >= 0) & (cc < e_size)
            mask[rr[ok], cc[ok]] = True
    return mask, int(amb.sum())


@pytest.mark.parametrize('case', ['c1', 'c5', 'c3'])
def test_per_view_dsm_vs_reference_golden(golden, eng_mod, lanes, case):
    aoi = json.loads(str(golden[case + '_aoi']))
    res = float(golden[case + '_res'])
    eng = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)
    assert eng.fit['degree'] >= 3, eng.fit
    depths, mats, want = golden[case + '_depths'], golden[case + '_mats'], golden[case + '_per_view']
    n_exact_bits = n_cells = 0
    for v in range(depths.shape[0]):
        got = eng.view_dsm(torch.from_numpy(depths[v]).cuda(), mats[v]).cpu().numpy()
        st = eng.stats()
        _, pts = op.unproject_depth(depths[v], mats[v])
        assert st['valid'] == pts.shape[0]
        assert st['exact'] == 0
        allow = np.zeros(want[v].shape, dtype=bool)
        if st['ambiguous']:
            amb_mask, _ = _ambiguous_cells(depths[v], mats[v], aoi, res, 1e-7)
            allow = cv2.dilate(amb_mask.astype(np.uint8), np.ones((5, 5), np.uint8)).astype(bool)
        # occupancy mask bit-exact
        assert np.array_equal(np.isnan(got)[~allow], np.isnan(want[v])[~allow]), 'view {}'.format(v)
        diff = np.abs(got.astype(np.float64) - want[v].astype(np.float64))
        diff[np.isnan(diff)] = 0
        assert diff[~allow].max() <= HEIGHT_TOL, 'view {} max diff {}'.format(v, diff[~allow].max())
        # Bit-identical float32 heights are expected everywhere except (a) where the ~1e-9 m noise of either
        # float64 chain crosses a float32 rounding boundary (~1e-4 of cells) and (b) hole cells filled from an
        # EVEN number of neighbours: the reference averages the two middle float64 altitudes and then rounds,
        # the 32-bit-key path averages the two rounded values (<= 1 ulp apart); (b) also reaches the 3x3 blur
        # neighbourhood of such a cell.
        _, pts = op.unproject_depth(depths[v], mats[v])
        raw = op._scatter_nanmax(op.enu_points_to_utm(pts, aoi), aoi['ul_easting'], aoi['ul_northing'], res, res,
                                 eng.e_size, eng.n_size)
        valid = (~np.isnan(raw)).astype(np.float32)
        nb = cv2.filter2D(valid, -1, np.ones((3, 3), np.float32), borderType=cv2.BORDER_CONSTANT)
        even_hole = np.isnan(raw) & (nb > 0) & (np.round(nb).astype(int) % 2 == 0)
        even_zone = cv2.dilate(even_hole.astype(np.uint8), np.ones((3, 3), np.uint8)).astype(bool)
        same = (got == want[v]) | (np.isnan(got) & np.isnan(want[v]))
        n_exact_bits += int(np.sum(same[~even_zone]))
        n_cells += int(np.sum(~even_zone))
    assert n_exact_bits / n_cells > 0.998, n_exact_bits / n_cells


@pytest.mark.parametrize('case', ['c1', 'c5', 'c3'])
def test_end_to_end_vs_reference_golden(golden, eng_mod, lanes, case):
    aoi = json.loads(str(golden[case + '_aoi']))
    res = float(golden[case + '_res'])
    eng = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)
    depths, mats = golden[case + '_depths'], golden[case + '_mats']
    V = depths.shape[0]
    stack = torch.empty((V, eng.n_size, eng.e_size), dtype=torch.float32, device='cuda')
    n_amb = 0
    for v in range(V):
        eng.view_dsm(torch.from_numpy(depths[v]).cuda(), mats[v], out=stack[v])
        n_amb += eng.stats()['ambiguous']
    got = eng.fuse_and_blur(stack).cpu().numpy()
    want = golden[case + '_fused']
    fragile = op.fusion_fragility([golden[case + '_per_view'][v] for v in range(V)])
    fragile = cv2.dilate(fragile.astype(np.uint8), np.ones((3, 3), np.uint8)).astype(bool)
    allow = np.zeros(want.shape, dtype=bool)
    if n_amb == 0:
        assert np.array_equal(np.isnan(got), np.isnan(want))
    else:
        # a point within 1e-7 cell of a cell edge may land on either side (the float64 chains differ by ~1e-9 m): the
        # cells it can reach, through the per-view 3x3 fill + 3x3 blur and the final blur, are set aside -- and counted
        for v in range(V):
            m, _ = _ambiguous_cells(depths[v], mats[v], aoi, res, 1e-7)
            allow |= m
        allow = cv2.dilate(allow.astype(np.uint8), np.ones((7, 7), np.uint8)).astype(bool)
        assert allow.sum() <= 121 * n_amb
        assert np.array_equal(np.isnan(got)[~allow], np.isnan(want)[~allow])
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    diff[np.isnan(diff)] = 0
    diff[allow] = 0
    bad = diff > HEIGHT_TOL
    # a >1 mm difference is only acceptable where the reference's own strict MAD test sits on a float32 tie
    print('[end-to-end {}] fused cells > {} m: {} of {} (fragile mask: {} cells); ambiguous points: {}'.format(
        case, HEIGHT_TOL, int(bad.sum()), bad.size, int(fragile.sum()), n_amb))
    assert not np.any(bad & ~fragile), 'unexplained cells: {}'.format(np.argwhere(bad & ~fragile)[:10])
    assert bad.sum() <= fragile.sum()
    assert bad.sum() <= max(2, 1e-4 * bad.size), int(bad.sum())      # absolute cap on the escape hatch


def test_captured_step_replays_bit_identically(golden, eng_mod, lanes):
    """DsmEngine.capture_step: stages A-C as one CUDA graph (internal streams included) == the eager calls, and a replay
    picks up new depth values written into the same device buffers."""
    case = 'c1'
    aoi = json.loads(str(golden[case + '_aoi']))
    res = float(golden[case + '_res'])
    eng = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)
    depths = [torch.from_numpy(d).cuda() for d in golden[case + '_depths']]
    mats = list(golden[case + '_mats'])
    V = len(depths)
    stack = torch.empty((V, eng.n_size, eng.e_size), dtype=torch.float32, device='cuda')
    eng.views_to_dsm(depths, mats, stack)
    want_stack = stack.clone()
    want = eng.fuse_and_blur(stack).clone()
    g = eng.capture_step(depths, mats, stack, fuse=True)
    assert g.launches_per_replay == 2 * V + 2      # K1 + K2 per view (the key-grid clear is a memset), fusion, blur
    for _ in range(3):
        stack.fill_(7.0)
        got = g.replay()
        torch.cuda.synchronize()
        assert _eq(stack.cpu().numpy(), want_stack.cpu().numpy())
        assert _eq(got.cpu().numpy(), want.cpu().numpy())
    # same buffers, new contents
    first = depths[0].clone()
    depths[0].copy_(depths[1])
    got2 = g.replay().clone()
    eng.views_to_dsm(depths, mats, stack)
    assert _eq(got2.cpu().numpy(), eng.fuse_and_blur(stack).cpu().numpy())
    depths[0].copy_(first)
    eng.close()


def test_exact_mode_and_altitude_fallback(golden, eng_mod, lanes):
    """max_degree=0 runs the exact chain for every pixel; a narrow fitted altitude range sends points through
    the per-point slow path.  Both must agree with the polynomial path."""
    case = 'c1'
    aoi = json.loads(str(golden[case + '_aoi']))
    res = float(golden[case + '_res'])
    depth = torch.from_numpy(golden[case + '_depths'][0]).cuda()
    M = golden[case + '_mats'][0]
    ref = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)
    a = ref.view_dsm(depth, M).cpu().numpy()
    exact = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes, max_degree=0)
    assert exact.fit['degree'] == 0
    hm = torch.empty_like(depth)
    b = exact.view_dsm(depth, M, height_map=hm).cpu().numpy()
    assert exact.stats()['exact'] > 0
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.nanmax(np.abs(a - b)) <= 1e-4
    height_map, _ = op.unproject_depth(golden[case + '_depths'][0], M)
    assert np.array_equal(np.isnan(hm.cpu().numpy()), np.isnan(height_map))
    assert np.nanmax(np.abs(hm.cpu().numpy() - height_map)) < 1e-4
    hm2 = torch.empty_like(depth)
    ref.view_dsm(depth, M, height_map=hm2)
    assert _eq(hm2.cpu().numpy(), hm.cpu().numpy()) or np.nanmax(np.abs(hm2.cpu().numpy() - height_map)) < 1e-4
    # narrow altitude range: alt_max below most of the terrain
    aoi2 = dict(aoi)
    aoi2['alt_max'] = aoi['alt_min'] + 1.0
    import vissatsatellitestereo_b200.engine as E
    narrow = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)
    narrow._aoi = E.aoi_struct(aoi, res, res, alt_margin=(0.0, -100.0))   # alt_hi = alt_max - 100
    narrow.fit = narrow.ctx.set_aoi(narrow._aoi, 5)
    c = narrow.view_dsm(depth, M).cpu().numpy()
    assert narrow.stats()['exact'] > 0
    assert np.array_equal(np.isnan(a), np.isnan(c))
    assert np.nanmax(np.abs(a - c)) <= 1e-4


def test_errors_are_loud(eng_mod):
    from vissatsatellitestereo_b200 import _native
    ctx = _native.Context(0)
    import ctypes as C
    with pytest.raises(_native.VisSatError):
        _native.check(_native.lib.vs_unproject_rasterize(ctx.handle, None, 4, 4, (C.c_double * 16)(), None, 1, None, None, None))
    bad = _native.vs_aoi()
    with pytest.raises(_native.VisSatError):
        ctx.set_aoi(bad)
    with pytest.raises(_native.VisSatError):
        _native.Context(9999)


def test_tma_pipeline_variant_is_bit_identical(golden, eng_mod, lanes, monkeypatch):
    """VISSAT_TMA=1 routes stage B (and the final blur) through the TMA-fed persistent kernels; results must be
    bit-identical to the plain-load kernels (and therefore to the reference-pinned expectations above)."""
    case = 'c3'
    aoi = json.loads(str(golden[case + '_aoi']))
    res = float(golden[case + '_res'])
    depths, mats = golden[case + '_depths'], golden[case + '_mats']
    plain = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)
    monkeypatch.setenv('VISSAT_TMA', '1')
    tma = eng_mod.DsmEngine(aoi, res, res, simd_lanes=lanes)          # the flag is read when the context is created
    monkeypatch.delenv('VISSAT_TMA')
    assert plain.e_size % 4 == 0                                      # TMA path applicable (row pitch multiple of 16 B)
    views_a, views_b = [], []
    for v in range(depths.shape[0]):
        d = torch.from_numpy(depths[v]).cuda()
        views_a.append(plain.view_dsm(d, mats[v]).clone())
        views_b.append(tma.view_dsm(d, mats[v]).clone())
        assert torch.equal(torch.nan_to_num(views_a[-1], nan=-1e9), torch.nan_to_num(views_b[-1], nan=-1e9)), v
    fa = plain.fuse_and_blur(torch.stack(views_a))
    fb = tma.fuse_and_blur(torch.stack(views_b))
    assert torch.equal(torch.nan_to_num(fa, nan=-1e9), torch.nan_to_num(fb, nan=-1e9))
    assert _eq(fb.cpu().numpy(), op.fuse_dsms([v.cpu().numpy() for v in views_a]))
    assert plain.last_nan_count() == tma.last_nan_count()
    # random NaN-heavy images through both blur kernels, incl. row bands
    rng = np.random.default_rng(11)
    for shape in [(64, 64), (100, 132), (257, 260)]:
        img = rng.normal(size=shape).astype(np.float32)
        img[rng.random(shape) < 0.3] = np.nan
        t = torch.from_numpy(img).cuda()
        want = cv2.medianBlur(img, 3)
        assert _eq(tma.median3x3(t).cpu().numpy(), want)
        r0, r1 = 5, shape[0] - 7
        band = t[r0 - 1:r1 + 1].contiguous()
        got = tma.median3x3(band, row_begin=r0, row_end=r1, in_row0=r0 - 1, h_total=shape[0]).cpu().numpy()
        assert _eq(got, want[r0:r1])


def test_key_space_stage_b_bit_identical_to_round1_kernel(eng_mod, lanes, monkeypatch):
    """finalize_keys.cu (hole fill + 3x3 median computed on the order-preserving keys; used by the sparse mode) against the
    round-1 float-space kernel (the dense default, itself pinned to cv2 / the reference goldens) on random key grids: all hole
    densities, NaN regions wider than the fill, odd row pitch, edge tiles, 1-row / 1-column grids."""
    from vissatsatellitestereo_b200 import synthetic as S
    import ctypes as C
    from vissatsatellitestereo_b200._native import lib, check
    cfg = S.scaled(S.CONFIGS['C1'], views=1, depth=64, grid=32)
    aoi = S.make_aoi(cfg, geodesy)
    monkeypatch.setenv('VISSAT_K2_KEYS', '1')
    new = eng_mod.DsmEngine(aoi, cfg.res, cfg.res, simd_lanes=lanes)          # the switches are read at context creation
    monkeypatch.delenv('VISSAT_K2_KEYS')
    monkeypatch.setenv('VISSAT_K2_LEGACY', '1')
    old = eng_mod.DsmEngine(aoi, cfg.res, cfg.res, simd_lanes=lanes)
    monkeypatch.delenv('VISSAT_K2_LEGACY')
    rng = np.random.default_rng(5)

    def run(e, keys, W, H):
        out = torch.empty((H, W), dtype=torch.float32, device='cuda')
        cnt = torch.zeros(1, dtype=torch.int64, device='cuda')
        check(lib.vs_grid_finalize(e.ctx.handle, C.c_void_p(keys.data_ptr()), W, H, C.c_void_p(out.data_ptr()), lanes,
                                   C.c_void_p(cnt.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out.cpu().numpy(), int(cnt.item())

    shapes = [(1, 1), (1, 37), (41, 1), (2, 2), (5, 17), (36, 72), (64, 128), (97, 131), (200, 260), (257, 516), (300, 64)]
    for (H, W) in shapes:
        for p_hole in (0.0, 0.1, 0.4, 0.8, 0.97, 1.0):
            alt = (30 + 20 * rng.normal(size=(H, W))).astype(np.float32)
            alt[rng.random((H, W)) < 0.1] = np.round(alt[rng.random((H, W)) < 0.1].mean())         # ties
            b = alt.view(np.uint32)
            keys = np.where(b >> 31 != 0, ~b, b | np.uint32(0x80000000)).astype(np.uint32)
            keys[rng.random((H, W)) < p_hole] = 0
            if H > 20 and W > 20:
                keys[H // 3:H // 3 + 9, W // 4:W // 4 + 11] = 0                                     # NaNs survive the fill
            kt = torch.from_numpy(keys.view(np.int32)).cuda()
            a, na = run(new, kt, W, H)
            o, no = run(old, kt, W, H)
            assert np.array_equal(a, o, equal_nan=True), (H, W, p_hole)
            assert na == no == int(np.isnan(a).sum())
    new.close()
    old.close()


@pytest.mark.parametrize('case', [dict(views=7, depth=256, grid=160), dict(views=1, depth=128, grid=64),
                                  dict(views=5, depth=(130, 203), grid=97), dict(views=9, depth=512, grid=300)])
def test_coscheduled_stage_ab_bit_identical_to_separate_kernels(eng_mod, lanes, case):
    """stage_ab.cu (one kernel per view step: stage B of the previous view + stage A of the next + the key-grid clear, work
    taken from two device queues) against the separate stage-A / stage-B kernels (themselves pinned to the reference goldens):
    per-view planes, NaN counters and the fused DSM, bit for bit; odd view counts (the views alternate between internal
    streams), a single view, a depth map whose row pitch is not a multiple of 4 pixels, repeated calls (the three rotating
    key grids of a stream must come back empty), 1..4 internal streams."""
    from vissatsatellitestereo_b200 import synthetic as S
    d = case['depth']
    cfg = S.scaled(S.CONFIGS['C1'], views=case['views'], depth=d if isinstance(d, int) else d[0], grid=case['grid'])
    if not isinstance(d, int):
        cfg.height, cfg.width = d
    scene = S.make_scene(cfg, geodesy, device='cuda')
    eng = eng_mod.DsmEngine(scene.aoi, cfg.res, cfg.res, simd_lanes=lanes)
    assert eng.fit['degree'] == 3
    V = cfg.n_views
    depths = list(scene.depths)
    depths[0] = depths[0].clone()
    depths[0][::3, 1::4] = float('nan')          # invalid pixels
    depths[0][5:9, :] = -1.0
    want = torch.empty((V, eng.n_size, eng.e_size), dtype=torch.float32, device='cuda')
    want_nan = torch.zeros(V, dtype=torch.int64, device='cuda')
    eng.set_coschedule(False)
    eng.views_to_dsm(depths, scene.mats, want, count_nan=want_nan)
    want_fused = eng.fuse_and_blur(want) if V >= 3 else None
    eng.set_coschedule(True)
    n0 = eng.launch_count()
    for streams in (4, 1, 2, 3):
        eng.set_streams(streams)
        for rep in range(2):
            got = torch.full_like(want, -7.0)
            got_nan = torch.full_like(want_nan, -1)
            eng.views_to_dsm(depths, scene.mats, got, count_nan=got_nan)
            assert torch.equal(torch.nan_to_num(got, nan=-1e9), torch.nan_to_num(want, nan=-1e9)), (streams, rep)
            assert torch.equal(got_nan, want_nan)
    # the co-scheduled path was really taken: V + (streams in use) launches per call instead of 2 V (+ memsets)
    assert eng.launch_count() - n0 == sum(2 * (V + min(s, 2, V)) for s in (4, 1, 2, 3))
    if want_fused is not None:
        assert torch.equal(torch.nan_to_num(eng.fuse_and_blur(got), nan=-1e9), torch.nan_to_num(want_fused, nan=-1e9))
    # a captured step replays it
    eng.set_streams(4)
    stack = torch.empty_like(want)
    g = eng.capture_step(depths, scene.mats, stack, fuse=V >= 3)
    stack.fill_(3.0)
    g.replay()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(stack, nan=-1e9), torch.nan_to_num(want, nan=-1e9))
    eng.close()


def test_fusion_split_search_variant_bit_exact(lanes):
    """VISSAT_FUSE_SPLIT=1 (order statistics of the multi-lane fusion kernels by a split search with a bracketed bisection as
    fallback; opt-in, read once per process -> a child process): bit-exact against numpy on random, tied and lane-structured
    inputs, V = 257 (32 lanes) and 400 (8 lanes)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from oracle import geodesy, pipeline as op
from vissatsatellitestereo_b200 import engine as E, synthetic as S
cfg = S.scaled(S.CONFIGS['C1'], views=1, depth=64, grid=32)
eng = E.DsmEngine(S.make_aoi(cfg, geodesy), cfg.res, cfg.res, simd_lanes=%d)
for V in (257, 400):
    rng = np.random.default_rng(V)
    H, W = 9, 41
    cube = (30 + 5 * rng.normal(size=(V, H, W))).astype(np.float32)
    cube[rng.random(cube.shape) < 0.3] = np.nan
    vi = np.arange(V, dtype=np.float32)[:, None]
    noise = rng.normal(size=(V, W)).astype(np.float32)
    cube[:, 1, :] = np.round(cube[:, 1, :])
    cube[:, 3, :] = np.where(vi %% 2 == 0, 10.0, 50.0) + noise
    cube[:, 4, :] = 5.0 * (vi %% 8) + 0.1 * noise
    cube[:, 5, :] = 0.37 * vi + 0.01 * noise
    cube[:, 6, :] = np.where(vi %% 8 == 3, -500.0, 20.0) + noise
    want = op.fuse_dsms([cube[v].copy() for v in range(V)], blur=False)
    got = eng.fuse(torch.from_numpy(cube).cuda()).cpu().numpy()
    assert np.array_equal(got, want, equal_nan=True), V
print('SPLIT_OK')
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), lanes)
    env = dict(os.environ, VISSAT_FUSE_SPLIT='1')
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'SPLIT_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
