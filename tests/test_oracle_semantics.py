"""Semantics the CUDA kernels must reproduce, established against the real third-party code on CPU:
cv2.medianBlur NaN behaviour, numpy's float32 pairwise sum, numpy nanmedian/nanmean fusion."""
import numpy as np
import cv2
import pytest

from oracle import geodesy, pipeline as op


def _eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('shape', [(1, 1), (1, 7), (7, 1), (3, 3), (5, 17), (6, 18), (9, 19), (33, 40), (64, 129)])
@pytest.mark.parametrize('nan_frac', [0.0, 0.3, 0.9])
def test_median3x3_emulation_equals_cv2(shape, nan_frac):
    lanes = op.detect_cv2_simd_lanes()
    assert lanes in (4, 8, 16, 32, 64)
    rng = np.random.default_rng(hash((shape, nan_frac)) % 2 ** 32)
    for _ in range(3):
        img = rng.normal(size=shape).astype(np.float32)
        img[rng.random(shape) < nan_frac] = np.nan
        assert _eq(op.median3x3_emul(img, lanes), cv2.medianBlur(img, 3))


@pytest.mark.parametrize('n', [1, 2, 5, 7, 8, 9, 13, 16, 50, 64, 100, 127, 128, 129, 136, 200, 257, 300])
def test_pairwise_sum_emulation_equals_numpy(n):
    rng = np.random.default_rng(n)
    a = (rng.normal(size=(64, n)) * 100).astype(np.float32)
    want = np.sum(a, axis=1)
    got = np.array([op.pairwise_sum_f32(a[i]) for i in range(a.shape[0])], dtype=np.float32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize('V', [3, 4, 8, 9, 50, 64, 130])
def test_fusion_bruteforce_equals_numpy(V):
    rng = np.random.default_rng(V)
    H, W = 12, 20
    cube = (30 + 5 * rng.normal(size=(H, W, V))).astype(np.float32)
    cube[rng.random(cube.shape) < 0.4] = np.nan
    cube[0, 0, :] = np.nan
    cube[0, 1, 2:] = np.nan
    cube[0, 2, :] = 7.25       # all equal -> mad 0, nothing rejected
    want = op.fuse_dsms([cube[:, :, v].copy() for v in range(V)], blur=False)
    got = np.array([[op.fuse_cell_bruteforce(cube[i, j]) for j in range(W)] for i in range(H)], dtype=np.float32)
    assert _eq(got, want)


def test_hole_fill_fast_equals_loop():
    rng = np.random.default_rng(5)
    for shape, frac in [((1, 1), 1.0), ((4, 9), 0.5), ((20, 31), 0.8), ((16, 16), 0.1)]:
        pts_n = int(shape[0] * shape[1] * (1 - frac)) + 1
        pts = np.stack([rng.uniform(0, shape[1], pts_n), -rng.uniform(0, shape[0], pts_n), rng.normal(size=pts_n)], 1)
        a = op.proj_to_grid(pts, 0.0, 0.0, 1.0, 1.0, shape[1], shape[0])
        b = op.proj_to_grid_fast(pts, 0.0, 0.0, 1.0, 1.0, shape[1], shape[0])
        assert _eq(a, b)


# ---- geodesy known answers (published by the third-party projects themselves)
def test_proj_documented_utm_example():
    # PROJ docs, operations/projections/utm: `echo 12 56 | proj +proj=utm +zone=32` -> 687071.44 6210141.33
    e, n = geodesy.utm_forward(56.0, 12.0, 32, False)
    assert abs(e - 687071.44) < 0.006 and abs(n - 6210141.33) < 0.006


def test_pymap3d_test_triples():
    # pymap3d tests: lla0 = (42, -82, 200) <-> xyz0 ; aer0 = (33, 70, 1000) -> enu0 -> lla1
    x, y, z = geodesy.geodetic2ecef(42.0, -82.0, 200.0)
    assert np.allclose([x, y, z], [660675.2518247, -4700948.68316, 4245737.66222], rtol=0, atol=1e-4)
    lat, lon, alt = geodesy.enu2geodetic(186.277521, 286.842228, 939.692621, 42.0, -82.0, 200.0)
    assert np.allclose([lat, lon, alt], [42.002581974253744, -81.997751960067460, 1139.7], rtol=0, atol=[1e-9, 1e-9, 0.01])


def test_meridian_arc_quadrature():
    from scipy.integrate import quad
    a = geodesy.PROJ_A
    f = 1 / geodesy.PROJ_RF
    e2 = 2 * f - f * f
    for lat in (-34.45, 10.0, 47.9941214, 80.0):
        M = quad(lambda p: a * (1 - e2) / (1 - e2 * np.sin(p) ** 2) ** 1.5, 0, np.radians(lat), epsabs=1e-9, epsrel=1e-14)[0]
        E, N = geodesy.utm_forward(lat, 9.0, 32, lat < 0)
        assert abs(E - 500000.0) < 1e-9
        assert abs(N - (0.9996 * M + (1e7 if lat < 0 else 0))) < 5e-8


def test_round_trips_and_reference_main_samples():
    rng = np.random.default_rng(0)
    lat = rng.uniform(-80, 84, 20000)
    lon = 9 + rng.uniform(-3, 3, 20000)
    E, N = geodesy.utm_forward(lat, lon, 32, False)
    la, lo = geodesy.utm_inverse(E, N, 32, False)
    assert np.abs(la - lat).max() * 111e3 < 1e-8 and np.abs(lo - lon).max() * 111e3 < 1e-8
    # lib/latlonalt_enu_converter.py:49-58 sample, round trip
    e, n, u = geodesy.latlonalt_to_enu(-34.450, -58.579, 20.31, -34.448, -58.577, -30.0)
    la, lo, al = geodesy.enu_to_latlonalt(e, n, u, -34.448, -58.577, -30.0)
    assert abs(la + 34.450) < 1e-12 and abs(lo + 58.579) < 1e-12 and abs(al - 20.31) < 1e-8
    # lib/latlon_utm_converter.py:68-69,82-83 samples (zone 32, both hemispheres)
    for s in (1, -1):
        lat = np.array([[s * 47.9941214]])
        lon = np.array([[7.8509671]])
        e, n = geodesy.latlon_to_eastnorh(lat, lon)
        la, lo = geodesy.eastnorth_to_latlon(e, n, 32, 'N' if s > 0 else 'S')
        assert abs(la[0, 0] - lat[0, 0]) < 1e-12 and abs(lo[0, 0] - lon[0, 0]) < 1e-12
    assert geodesy.utm_zone_number(-34.45, -58.58) == 21
    assert geodesy.utm_zone_number(60.0, 5.0) == 32 and geodesy.utm_zone_number(75.0, 10.0) == 33
