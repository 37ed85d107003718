"""The step entry point on a real work_dir: aggregate_2p5d.run_fuse writes the reference's file set and its
contents match the reference-run golden vectors."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import pipeline as op

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_work_dir(golden, case, work_dir):
    from vissatsatellitestereo_b200 import synthetic as S
    aoi = json.loads(str(golden[case + '_aoi']))
    scene = S.SynthScene(cfg=None, aoi=aoi)
    for v in range(golden[case + '_depths'].shape[0]):
        scene.names.append(S.view_name(v))
        scene.mats.append(golden[case + '_mats'][v])
        scene.depths.append(torch.from_numpy(golden[case + '_depths'][v]))
    S.write_work_dir(scene, work_dir)
    # a file the reference would skip with "something funny is happening" (aggregate_2p5d_util.py:66-69)
    open(os.path.join(work_dir, 'colmap/mvs/stereo/depth_maps', 'notes.geometric.txt'), 'w').close()
    return aoi


@pytest.mark.parametrize('case,res', [('c1', 0.5), ('c3', 0.3)])
def test_run_fuse_file_outputs(golden, tmp_path, case, res):
    from vissatsatellitestereo_b200 import aggregate_2p5d, produce_dsm
    from vissatsatellitestereo_b200.lib.dsm_util import read_dsm_tif
    from vissatsatellitestereo_b200.lib.ply_np_converter import ply2np
    work_dir = str(tmp_path / 'work')
    os.makedirs(work_dir)
    aoi = _write_work_dir(golden, case, work_dir)
    produce_dsm.e_resolution = produce_dsm.n_resolution = res      # module globals, as in the reference (:41-42)
    try:
        # stale output must be wiped (aggregate_2p5d_util.py:127-128)
        os.makedirs(os.path.join(work_dir, 'colmap/mvs/dsm/dsm_tif'))
        open(os.path.join(work_dir, 'colmap/mvs/dsm/dsm_tif/zzz_stale.tif'), 'w').close()
        aggregate_2p5d.run_fuse(work_dir, max_processes=2)
    finally:
        produce_dsm.e_resolution = produce_dsm.n_resolution = 0.5
    tif_dir = os.path.join(work_dir, 'colmap/mvs/dsm/dsm_tif')
    tifs = sorted(os.listdir(tif_dir))
    V = golden[case + '_depths'].shape[0]
    assert tifs == ['{:04d}.tif'.format(v) for v in range(V)]
    want_pv = golden[case + '_per_view']
    for v, name in enumerate(tifs):
        got, meta = read_dsm_tif(os.path.join(tif_dir, name))
        assert np.array_equal(np.isnan(got), np.isnan(want_pv[v]))
        assert np.nanmax(np.abs(got - want_pv[v])) <= 1e-3
        assert meta['geo'] == (aoi['ul_easting'], res, 0.0, aoi['ul_northing'], 0.0, -res)
        assert (meta['zone_number'], meta['hemisphere']) == (aoi['zone_number'], aoi['hemisphere'])
        assert os.path.exists(os.path.join(work_dir, 'colmap/mvs/dsm/dsm_jpg', name[:-4] + '.jpg'))
        assert os.path.exists(os.path.join(work_dir, 'colmap/mvs/dsm/dsm_img_grid', name[:-4] + '.jpg'))
    out_dir = os.path.join(work_dir, 'mvs_results/aggregate_2p5d')
    fused, _ = read_dsm_tif(os.path.join(out_dir, 'aggregate_2p5d_dsm.tif'))
    want = golden[case + '_fused']
    assert np.array_equal(np.isnan(fused), np.isnan(want))
    fragile = op.fusion_fragility([want_pv[v] for v in range(V)])
    import cv2
    fragile = cv2.dilate(fragile.astype(np.uint8), np.ones((3, 3), np.uint8)).astype(bool)
    diff = np.abs(fused.astype(np.float64) - want.astype(np.float64))
    diff[np.isnan(diff)] = 0
    bad = diff > 1e-3
    print('[run_fuse {}] fused cells > 1e-3 m: {} of {} (fragile mask: {} cells)'.format(case, int(bad.sum()), bad.size,
                                                                                       int(fragile.sum())))
    assert not np.any(bad & ~fragile)
    assert bad.sum() <= max(2, 1e-4 * bad.size), int(bad.sum())      # absolute cap on the escape hatch
    assert os.path.exists(os.path.join(out_dir, 'aggregate_2p5d_dsm.jpg'))
    pts, color, comments = ply2np(os.path.join(out_dir, 'aggregate_2p5d.ply'))
    assert pts.shape == (int((~np.isnan(fused)).sum()), 3) and color.shape == pts.shape
    assert comments == ['projection: UTM {}{}'.format(aoi['zone_number'], aoi['hemisphere'])]
    # vertices sit at the upper-left corner of their cell (aggregate_2p5d.py:92-107)
    ii, jj = np.nonzero(~np.isnan(fused))
    assert np.array_equal(pts[:, 0], aoi['ul_easting'] + jj * res)
    assert np.array_equal(pts[:, 1], aoi['ul_northing'] - ii * res)
    assert np.array_equal(pts[:, 2].astype(np.float32), fused[ii, jj])


def test_fuse_existing_tifs(golden, tmp_path):
    """Stage C from files (the reference's aggregate_2p5d.py:57-81 on its own): bit-exact."""
    from vissatsatellitestereo_b200.aggregate_2p5d import fuse_dsm_tifs
    from vissatsatellitestereo_b200.lib.dsm_util import write_dsm_tif
    case = 'c5'
    aoi = json.loads(str(golden[case + '_aoi']))
    files = []
    for v, dsm in enumerate(golden[case + '_per_view']):
        f = str(tmp_path / '{:04d}.tif'.format(v))
        write_dsm_tif(dsm, f, (aoi['ul_easting'], aoi['ul_northing'], 0.3, 0.3), (21, 'S'), nodata_val=-10000)
        files.append(f)
    got = fuse_dsm_tifs(files)
    assert np.array_equal(got, golden[case + '_fused'], equal_nan=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_result_identical_to_one_gpu():
    """the fused DSM must not depend on the number of ranks (bit-identical)."""
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29533',
                        os.path.join(REPO, 'tests', 'mgpu_check.py')], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'MGPU_OK' in r.stdout and 'MISMATCH' not in r.stdout
    # stage-B peer stores: bit-identical too wherever CUDA IPC peer mapping is available
    assert 'MGPU_PEER_OK' in r.stdout or 'MGPU_PEER_UNAVAILABLE' in r.stdout
