"""The oracle against vectors produced by running the reference itself (tests/golden/make_golden.py)."""
import json

import numpy as np
import pytest

from oracle import pipeline as op


def _eq(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize('k', range(5))
def test_proj_to_grid_matches_reference(golden, k):
    pts = golden['ptg{}_points'.format(k)]
    xoff, yoff, xres, yres, xs, ys = golden['ptg{}_args'.format(k)]
    want = golden['ptg{}_dsm'.format(k)]
    got = op.proj_to_grid(pts, xoff, yoff, xres, yres, int(xs), int(ys))
    assert got.dtype == np.float64 and got.shape == want.shape
    assert _eq(got, want)
    assert _eq(op.proj_to_grid_fast(pts, xoff, yoff, xres, yres, int(xs), int(ys)), want)


def test_read_array_matches_reference(golden, tmp_path):
    p = tmp_path / 'a.bin'
    p.write_bytes(golden['read_array_file'].tobytes())
    got = op.read_array(str(p))
    assert got.dtype == np.float32
    assert _eq(got, golden['read_array_out'])
    p2 = tmp_path / 'b.bin'
    op.write_array(str(p2), golden['read_array_in'])
    assert p2.read_bytes() == p.read_bytes()


@pytest.mark.parametrize('case', ['c1', 'c5', 'c3'])
def test_whole_step_matches_reference(golden, case):
    """Per-view DSMs and the fused DSM, bit-for-bit against reference run_fuse outputs."""
    aoi = json.loads(str(golden[case + '_aoi']))
    res = float(golden[case + '_res'])
    mats = golden[case + '_mats']
    depths = golden[case + '_depths']
    want_pv = golden[case + '_per_view']
    want_fused = golden[case + '_fused']
    per_view = []
    for v in range(depths.shape[0]):
        dsm, _ = op.convert_depth_map(depths[v], mats[v], aoi, res, res, fast=(v % 2 == 0))
        dsm = op.tif_roundtrip(dsm)
        assert dsm.dtype == np.float32
        assert _eq(dsm, want_pv[v]), 'view {}'.format(v)
        per_view.append(dsm)
    fused = op.tif_roundtrip(op.fuse_dsms(per_view))
    assert _eq(fused, want_fused)
    geo = golden[case + '_geo']
    assert geo[0] == aoi['ul_easting'] and geo[3] == aoi['ul_northing'] and geo[1] == res and geo[5] == -res
