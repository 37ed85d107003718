"""the aggregate_3d tail (aggregate_3d.py:54-83).  Golden vectors come from running the
reference's own aggregate_3d.run_fuse (tests/golden/make_golden_3d.py)."""
import json
import os

import numpy as np
import pytest

from oracle import pipeline as op

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ['a3d0', 'a3d1']


@pytest.fixture(scope='module')
def golden3d():
    return np.load(os.path.join(REPO, 'tests', 'golden', 'reference_golden_3d.npz'), allow_pickle=False)


def _fused_ply(golden3d, case, tmp_path):
    path = str(tmp_path / (case + '_fused.ply'))
    with open(path, 'wb') as fp:
        fp.write(golden3d[case + '_fused_ply'].tobytes())
    return path


@pytest.mark.parametrize('case', CASES)
def test_oracle_tail_matches_reference(golden3d, case, tmp_path):
    """ply2np reads the file the reference's plyfile wrote; the oracle tail reproduces the reference's outputs."""
    from vissatsatellitestereo_b200.lib.ply_np_converter import ply2np
    pts, color, comments = ply2np(_fused_ply(golden3d, case, tmp_path))
    assert pts.dtype == np.float64 and color.dtype == np.uint8 and comments is None
    assert np.array_equal(color, golden3d[case + '_utm_color'])
    aoi = json.loads(str(golden3d[case + '_aoi']))
    res = float(golden3d[case + '_res'])
    utm_pts, dsm = op.aggregate_3d_tail(pts, aoi, res, res, fast=False)
    assert np.array_equal(utm_pts, golden3d[case + '_utm'])
    assert np.array_equal(dsm, golden3d[case + '_dsm'], equal_nan=True)
    utm_pts2, dsm2 = op.aggregate_3d_tail(pts, aoi, res, res, fast=True)
    assert np.array_equal(dsm2, dsm, equal_nan=True)


def test_fuse_without_colmap_or_ply_is_an_error(tmp_path, monkeypatch):
    from vissatsatellitestereo_b200 import aggregate_3d
    monkeypatch.setenv('PATH', str(tmp_path))
    os.makedirs(str(tmp_path / 'colmap/mvs'))
    with pytest.raises(FileNotFoundError):
        aggregate_3d.fuse(str(tmp_path / 'colmap'))
    open(str(tmp_path / 'colmap/mvs/fused.ply'), 'wb').close()
    aggregate_3d.fuse(str(tmp_path / 'colmap'))          # existing cloud is accepted


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_run_fuse_3d_file_outputs(golden3d, case, tmp_path):
    from vissatsatellitestereo_b200 import aggregate_3d, produce_dsm
    from vissatsatellitestereo_b200.lib.dsm_util import read_dsm_tif
    from vissatsatellitestereo_b200.lib.ply_np_converter import ply2np
    work_dir = str(tmp_path / 'work')
    os.makedirs(os.path.join(work_dir, 'colmap/mvs'))
    aoi = json.loads(str(golden3d[case + '_aoi']))
    res = float(golden3d[case + '_res'])
    with open(os.path.join(work_dir, 'aoi.json'), 'w') as fp:
        json.dump(aoi, fp, indent=2)
    os.replace(_fused_ply(golden3d, case, tmp_path), os.path.join(work_dir, 'colmap/mvs/fused.ply'))
    produce_dsm.e_resolution = produce_dsm.n_resolution = res
    try:
        aggregate_3d.run_fuse(work_dir)
    finally:
        produce_dsm.e_resolution = produce_dsm.n_resolution = 0.5
    out_dir = os.path.join(work_dir, 'mvs_results/aggregate_3d')
    pts, color, comments = ply2np(os.path.join(out_dir, 'aggregate_3d.ply'))
    assert comments == json.loads(str(golden3d[case + '_comments']))
    assert np.array_equal(color, golden3d[case + '_utm_color'])
    want = golden3d[case + '_utm']
    assert pts.shape == want.shape and pts.dtype == np.float64
    # float64 chain on the device vs numpy/libm on the host: a few ulps of 6e6 m
    assert np.max(np.abs(pts[:, :2] - want[:, :2])) <= 2e-8
    assert np.max(np.abs(pts[:, 2] - want[:, 2])) <= 2e-8
    dsm, meta = read_dsm_tif(os.path.join(out_dir, 'aggregate_3d_dsm.tif'))
    want_dsm = golden3d[case + '_dsm']
    assert np.array_equal(np.isnan(dsm), np.isnan(want_dsm))
    assert np.nanmax(np.abs(dsm - want_dsm)) <= 1e-3          # north_star tolerance: heights within 1e-3 m
    assert np.mean(dsm[~np.isnan(dsm)] == want_dsm[~np.isnan(dsm)]) > 0.999
    assert meta['geo'] == (aoi['ul_easting'], res, 0.0, aoi['ul_northing'], 0.0, -res)
    assert os.path.exists(os.path.join(out_dir, 'aggregate_3d_dsm.jpg'))
