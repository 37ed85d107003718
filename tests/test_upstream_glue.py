"""host glue either side of the path -- write_aoi (stereo_pipeline.py:185-226) and the
inv_proj_mats derivation (reparam_depth.py:117-141).  CPU only."""
import json
import math
import os

import numpy as np
import pytest

from oracle import geodesy
from vissatsatellitestereo_b200 import synthetic as S


def _rotation_to_quaternion(R):
    w = math.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = math.copysign(math.sqrt(max(0.0, 1 + R[0, 0] - R[1, 1] - R[2, 2])) / 2, R[2, 1] - R[1, 2])
    y = math.copysign(math.sqrt(max(0.0, 1 - R[0, 0] + R[1, 1] - R[2, 2])) / 2, R[0, 2] - R[2, 0])
    z = math.copysign(math.sqrt(max(0.0, 1 - R[0, 0] - R[1, 1] + R[2, 2])) / 2, R[1, 0] - R[0, 1])
    return w, x, y, z


def test_write_aoi_matches_explorer_config(tmp_path):
    """aoi_config/MVS3DM_Explorer.json:4-11 bounding box; corners cross-checked against the PROJ inverse restatement."""
    from vissatsatellitestereo_b200.stereo_pipeline import write_aoi, utm_to_latlon
    config = {'work_dir': str(tmp_path), 'alt_min': -30.0, 'alt_max': 120.0,
              'bounding_box': {'zone_number': 21, 'hemisphere': 'S', 'ul_easting': 354052.3651180889,
                               'ul_northing': 6182702.10540914, 'width': 712.0, 'height': 652.0}}
    aoi = write_aoi(config)
    with open(os.path.join(str(tmp_path), 'aoi.json')) as fp:
        assert json.load(fp) == aoi
    assert list(aoi.keys()) == ['zone_number', 'hemisphere', 'ul_easting', 'ul_northing', 'lr_easting', 'lr_northing',
                                'width', 'height', 'lat_min', 'lat_max', 'lon_min', 'lon_max', 'alt_min', 'alt_max']
    assert aoi['lr_easting'] == 354052.3651180889 + 712.0 and aoi['lr_northing'] == 6182702.10540914 - 652.0
    # UL corner of the Explorer AOI through the (independent) PROJ inverse: -34.486961674 / -58.589466115
    lat, lon = utm_to_latlon(354052.3651180889, 6182702.10540914, 21, False)
    assert abs(lat - -34.486961674) < 5e-8 and abs(lon - -58.589466115) < 5e-8
    for e, n in ((aoi['ul_easting'], aoi['ul_northing']), (aoi['lr_easting'], aoi['lr_northing'])):
        plat, plon = geodesy.utm_inverse(e, n, 21, True)
        lat, lon = utm_to_latlon(e, n, 21, False)
        assert abs(lat - plat) < 5e-8 and abs(lon - plon) < 5e-8
    assert aoi['lat_min'] < aoi['lat_max'] < 0 and aoi['lon_min'] < aoi['lon_max'] < 0
    # northern hemisphere / other zone
    lat, lon = utm_to_latlon(395201.3103816, 5673135.2407718, 32, True)
    assert abs(lat - 51.2) < 5e-8 and abs(lon - 7.5) < 5e-8


def test_inv_proj_mats_from_camera_dict(tmp_path):
    """P4 = [K[R|t]; last_row], M = inv(P4): rebuilds the matrices the synthetic scenes (and the adapted COLMAP) use."""
    from vissatsatellitestereo_b200 import reparam_depth as RD
    from vissatsatellitestereo_b200.aggregate_2p5d_util import load_inv_proj_mats
    cfg = S.scaled(S.CONFIGS['C1'], views=4, depth=64, grid=32, name='glue')
    camera_dict, last_rows, want = {}, {}, {}
    for v in range(4):
        M, P4 = S.make_camera(cfg, v, cfg.alt_min)
        K = np.array([[P4[2, :3] @ P4[2, :3]]])            # |r3| = 1 for a rotation row
        assert abs(K[0, 0] - 1) < 1e-12
        R3 = P4[2, :3]
        # decompose P3 = K [R|t] (K upper triangular with K22 = 1, zero skew here)
        f = np.linalg.norm(np.cross(P4[0, :3], R3))
        cx, cy = P4[0, :3] @ R3, P4[1, :3] @ R3
        R = np.vstack(((P4[0, :3] - cx * R3) / f, (P4[1, :3] - cy * R3) / f, R3))
        Kmat = np.array([[f, 0, cx], [0, f, cy], [0, 0, 1.0]])
        t = np.linalg.solve(Kmat, P4[:3, 3])
        name = S.view_name(v)
        camera_dict[name] = (cfg.width, cfg.height, f, f, cx, cy, 0.0) + _rotation_to_quaternion(R) + tuple(t)
        last_rows[name] = P4[3]
        want[name] = M
        assert np.allclose(RD.quaternion_to_rotation(*_rotation_to_quaternion(R)), R, atol=1e-12)
    # non-unit quaternions are normalised (pyquaternion does the same)
    assert np.allclose(RD.quaternion_to_rotation(2, 0, 0, 0), np.eye(3))
    with open(str(tmp_path / 'last_rows.txt'), 'w') as fp:                # reparam_depth.py:171-174 format
        for name in sorted(last_rows):
            vec = last_rows[name]
            fp.write('{} {} {} {} {}\n'.format(name, vec[0], vec[1], vec[2], vec[3]))
    path = RD.ensure_inv_proj_mats(str(tmp_path), camera_dict)
    got = load_inv_proj_mats(str(tmp_path))
    assert sorted(got) == sorted(want)
    pix = np.array([[10.0, 20.0, 1.0, 597990.0]]).T
    for name in want:
        a, b = got[name] @ pix, want[name] @ pix
        assert np.allclose(a[:3] / a[3], b[:3] / b[3], rtol=0, atol=1e-6)       # same ENU point to a micrometre
    # an existing file is left alone
    os.utime(path, (1, 1))
    RD.ensure_inv_proj_mats(str(tmp_path), camera_dict)
    assert os.path.getmtime(path) == 1


def test_pipeline_rejects_out_of_scope_steps(tmp_path):
    from vissatsatellitestereo_b200.stereo_pipeline import StereoPipeline
    cfgfile = str(tmp_path / 'c.json')
    config = {'work_dir': str(tmp_path / 'w'), 'alt_min': -30.0, 'alt_max': 120.0,
              'bounding_box': {'zone_number': 21, 'hemisphere': 'S', 'ul_easting': 354052.3651180889,
                               'ul_northing': 6182702.10540914, 'width': 100.0, 'height': 100.0},
              'steps_to_run': {'clean_data': False, 'colmap_mvs': True, 'aggregate_2p5d': False, 'aggregate_3d': False}}
    with open(cfgfile, 'w') as fp:
        json.dump(config, fp)
    os.makedirs(config['work_dir'])
    with pytest.raises(NotImplementedError):
        StereoPipeline(cfgfile).run()
    config['steps_to_run']['colmap_mvs'] = False
    with open(cfgfile, 'w') as fp:
        json.dump(config, fp)
    StereoPipeline(cfgfile).run()
    assert os.path.exists(os.path.join(config['work_dir'], 'aoi.json'))
    with open(os.path.join(config['work_dir'], 'runtime.txt')) as fp:
        assert 'aggregate_2p5d, skipped' in fp.read()


@pytest.mark.parametrize('model', ['perspective', 'pinhole'])
def test_reparam_depth_matches_reference_files(tmp_path, model):
    """reparam_depth(sparse_dir, save_dir, camera_model) (reparam_depth.py:69-195): the five files it writes, byte for
    byte against what the REFERENCE wrote for the same sparse model (tests/golden/make_golden_reparam.py ran
    /root/reference/reparam_depth.py with the reference's own colmap/read_model.py)."""
    from vissatsatellitestereo_b200.reparam_depth import reparam_depth
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reparam', model)
    reparam_depth(os.path.join(base, 'sparse'), str(tmp_path), camera_model=model)
    names = ['depth_ranges.txt', 'last_rows.txt', 'raw_depth.txt', 'reference_plane.txt', 'reparam_depth.txt']
    assert sorted(os.listdir(str(tmp_path))) == names
    for n in names:
        with open(os.path.join(base, 'want', n)) as fa, open(os.path.join(str(tmp_path), n)) as fb:
            assert fa.read() == fb.read(), n
    # and the matrices derived from last_rows.txt invert back to P4 = [K [R | t]; last_row] (reparam_depth.py:117-141)
    from vissatsatellitestereo_b200.reparam_depth import read_last_rows
    rows = read_last_rows(os.path.join(str(tmp_path), 'last_rows.txt'))
    assert len(rows) == 5 and all(r.shape == (4,) and r[0] == 0 and r[1] == 0 and r[2] > 0 for r in rows.values())


def test_read_model_text(tmp_path):
    from vissatsatellitestereo_b200.colmap.read_model import read_model
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reparam', 'perspective', 'sparse')
    cams, imgs, pts = read_model(base, ext='.txt')
    assert len(cams) == 5 and len(imgs) == 5 and len(pts) == 400
    assert cams[1].model == 'PERSPECTIVE' and cams[1].params.shape == (5,) and imgs[3].name == '0002.png'
    assert imgs[1].qvec.shape == (4,) and imgs[1].tvec.shape == (3,) and imgs[1].xys.shape[1] == 2
    p = pts[7]
    assert p.xyz.shape == (3,) and p.image_ids.size == p.point2D_idxs.size >= 2
    with pytest.raises(NotImplementedError):
        read_model(base, ext='.bin')


def test_utm_to_latlon_known_values_of_the_utm_package():
    """`utm.to_latlon` (utm==0.4.2, requirements.txt; called by stereo_pipeline.py:198-205) is absent from the image.  Its
    own test-suite (test/test_utm.py: KnownValues, compared to 4 decimals there) anchors the restatement in
    stereo_pipeline.utm_to_latlon: eastings / northings rounded to a metre <-> latitude / longitude to 5 decimals.  The same
    vectors are also run through the oracle's exact inverse (pinned to the 50-digit map), which must agree with the
    restated series to < 5e-8 deg away from the pole-ward limit of UTM."""
    from oracle import geodesy
    from vissatsatellitestereo_b200.stereo_pipeline import utm_to_latlon
    known = [((50.77535, 6.08389), (294409, 5628898, 32, True)),        # Aachen
             ((40.71435, -74.00597), (583960, 4507523, 18, True)),      # New York
             ((-41.28646, 174.77624), (313784, 5427057, 60, False)),    # Wellington
             ((-33.92487, 18.42406), (261878, 6243186, 34, False)),     # Capetown
             ((-32.89018, -68.84405), (514586, 6360877, 19, False)),    # Mendoza
             ((64.83778, -147.71639), (466013, 7190568, 6, True)),      # Fairbanks
             ((56.79680, -5.00601), (377486, 6296562, 30, True)),       # Ben Nevis
             ((84.0, -5.00601), (476594, 9328501, 30, True))]           # latitude 84
    for (lat, lon), (e, n, zone, northern) in known:
        got = utm_to_latlon(e, n, zone, northern)
        assert abs(got[0] - lat) < 5e-5 and abs(got[1] - lon) < 5e-5, (lat, lon, got)
        exact = geodesy.utm_inverse(e, n, zone, not northern)
        tol = 5e-8 if abs(lat) < 80 else 5e-7
        assert abs(got[0] - float(exact[0])) < tol and abs(got[1] - float(exact[1])) < tol, (lat, lon, got, exact)
