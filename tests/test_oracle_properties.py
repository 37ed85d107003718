"""Property tests (hypothesis) of the oracle: the size-independent invariants the CUDA path is also held to in
tests/test_gpu_fullsize.py (SURVEY.md §4 test plan (ii))."""
import cv2
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pipeline as op

SET = settings(max_examples=40, deadline=None)


def _eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@SET
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 9), st.integers(1, 9), st.integers(0, 200))
def test_scatter_nanmax_is_order_invariant_and_idempotent(seed, xs, ys, n):
    """lib/proj_to_grid.py:42-61: per-cell nanmax does not depend on the point order, ignores NaN values and points
    outside the grid, and re-rasterising the occupied cells' own (centre, value) points reproduces the grid."""
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-2, xs + 2, n), -rng.uniform(-2, ys + 2, n), rng.normal(size=n)], 1)
    pts[rng.random(n) < 0.1, 2] = np.nan
    a = op._scatter_nanmax(pts, 0.0, 0.0, 1.0, 1.0, xs, ys)
    b = op._scatter_nanmax(pts[rng.permutation(n)], 0.0, 0.0, 1.0, 1.0, xs, ys)
    assert _eq(a, b)
    inside = (pts[:, 0] >= 0) & (pts[:, 0] < xs) & (-pts[:, 1] >= 0) & (-pts[:, 1] < ys) & ~np.isnan(pts[:, 2])
    assert _eq(op._scatter_nanmax(pts[inside], 0.0, 0.0, 1.0, 1.0, xs, ys), a)
    rr, cc = np.nonzero(~np.isnan(a))
    again = np.stack([cc + 0.5, -(rr + 0.5), a[rr, cc]], 1) if rr.size else np.empty((0, 3))
    assert _eq(op._scatter_nanmax(again, 0.0, 0.0, 1.0, 1.0, xs, ys), a)


@SET
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 12), st.integers(1, 12), st.floats(0.0, 1.0))
def test_hole_fill_reads_the_prefill_grid(seed, h, w, frac):
    """lib/proj_to_grid.py:65-79: occupied cells are unchanged, a hole is filled iff it has an occupied neighbour in
    the PRE-fill grid (no cascade into a second ring of holes), and the fill lies within the range of the data."""
    rng = np.random.default_rng(seed)
    raw = rng.normal(size=(h, w))
    raw[rng.random((h, w)) < frac] = np.nan
    filled = op.fill_holes_fast(raw)
    occ = ~np.isnan(raw)
    assert np.array_equal(filled[occ], raw[occ])
    pad = np.pad(occ, 1)
    has_nb = np.zeros((h, w), bool)
    for dy in (0, 1, 2):
        for dx in (0, 1, 2):
            if (dy, dx) != (1, 1):
                has_nb |= pad[dy:dy + h, dx:dx + w]
    assert np.array_equal(~np.isnan(filled), occ | has_nb)
    if occ.any():
        assert np.nanmin(filled) >= np.nanmin(raw) and np.nanmax(filled) <= np.nanmax(raw)


@SET
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 20), st.integers(1, 40), st.floats(0.0, 0.95))
def test_median3x3_emulation_equals_cv2_everywhere(seed, h, w, nan_frac):
    lanes = op.detect_cv2_simd_lanes()
    rng = np.random.default_rng(seed)
    img = rng.normal(size=(h, w)).astype(np.float32)
    img[rng.random((h, w)) < nan_frac] = np.nan
    assert _eq(op.median3x3_emul(img, lanes), cv2.medianBlur(img, 3))


@SET
@given(st.integers(0, 2 ** 32 - 1), st.integers(1, 40), st.floats(0.0, 0.9))
def test_fusion_cell_properties(seed, V, nan_frac):
    """aggregate_2p5d.py:65-78 per cell: <= 2 measurements -> NaN; otherwise the result is a mean of at least
    ceil(k/2) of the measurements, lies within their range, equals the value when all measurements agree, and only the
    float32 summation order depends on the view order (|difference| stays within a few ulps)."""
    rng = np.random.default_rng(seed)
    x = (50 + 10 * rng.normal(size=V)).astype(np.float32)
    x[rng.random(V) < nan_frac] = np.nan
    k = int(np.sum(~np.isnan(x)))
    got = op.fuse_cell_bruteforce(x)
    want = op.fuse_dsms([np.full((1, 1), v, np.float32) for v in x], blur=False)[0, 0]
    assert _eq(np.float32(got), np.float32(want))
    if k <= 2:
        assert np.isnan(got)
        return
    assert np.nanmin(x) <= got <= np.nanmax(x)
    perm = op.fuse_cell_bruteforce(x[rng.permutation(V)])
    assert abs(float(perm) - float(got)) <= 8 * np.spacing(np.float32(abs(got)))
    same = np.where(np.isnan(x), np.nan, np.float32(7.25)).astype(np.float32)
    assert op.fuse_cell_bruteforce(same) == np.float32(7.25)
