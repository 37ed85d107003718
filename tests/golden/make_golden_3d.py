#!/usr/bin/env python
"""Generate tests/golden/reference_golden_3d.npz by EXECUTING the reference's aggregate_3d.run_fuse tail
(aggregate_3d.py:54-83) from /root/reference on a synthetic fused ENU point cloud.

    python tests/golden/make_golden_3d.py          (build container only)

Runs verbatim: aggregate_3d.run_fuse, lib/ply_np_converter.py + lib/plyfile.py, coordinate_system.local_to_global,
lib/latlon_utm_converter.latlon_to_eastnorh, produce_dsm.produce_dsm_from_points, lib/proj_to_grid.proj_to_grid,
lib/dsm_util.write_dsm_tif/read_dsm_tif.  Stand-ins: those of make_golden.py, plus `aggregate_3d.fuse` (it shells out
to the COLMAP binary, absent here): the input `fused.ply` is written by the reference's own np2ply instead.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import make_golden as MG  # noqa: E402
from oracle import geodesy, pipeline as OP  # noqa: E402
from vissatsatellitestereo_b200 import synthetic as S  # noqa: E402


def main():
    MG.install_shims()
    sys.path.insert(0, REF)
    import aggregate_3d                                  # reference module
    import produce_dsm                                   # reference module
    from lib.ply_np_converter import np2ply, ply2np      # reference functions
    from lib.dsm_util import read_dsm_tif                # reference function

    aggregate_3d.fuse = lambda colmap_dir: None          # COLMAP binary: absent
    out = {}
    rng = np.random.default_rng(3)
    for name, base, views, depth, grid, res, keep in (('a3d0', 'C1', 3, 160, 96, 0.5, 0.2),
                                                       ('a3d1', 'C5', 2, 128, 128, 0.3, 0.5)):
        cfg = S.scaled(S.CONFIGS[base], views=views, depth=depth, grid=grid, name='g_' + name)
        scene = S.make_scene(cfg, geodesy, device='cpu')
        # "fused" cloud = a random subset of the unprojected depth pixels of a few views, in ENU
        pts = []
        for d, M in zip(scene.depths, scene.mats):
            _, p = OP.unproject_depth(d.numpy(), M)
            pts.append(p[rng.random(p.shape[0]) < keep, :3])
        pts = np.concatenate(pts)
        color = rng.integers(0, 256, size=(pts.shape[0], 3)).astype(np.uint8)
        produce_dsm.e_resolution = res
        produce_dsm.n_resolution = res
        work_dir = tempfile.mkdtemp(prefix='vissat_golden3d_')
        try:
            os.makedirs(os.path.join(work_dir, 'colmap/mvs'))
            with open(os.path.join(work_dir, 'aoi.json'), 'w') as fp:
                json.dump(scene.aoi, fp, indent=2)
            np2ply(pts, os.path.join(work_dir, 'colmap/mvs/fused.ply'), color=color, use_double=True)
            with open(os.path.join(work_dir, 'colmap/mvs/fused.ply'), 'rb') as fp:
                ply_bytes = np.frombuffer(fp.read(), dtype=np.uint8)
            aggregate_3d.run_fuse(work_dir)
            utm_pts, utm_color, comments = ply2np(os.path.join(work_dir, 'mvs_results/aggregate_3d/aggregate_3d.ply'))
            dsm, meta = read_dsm_tif(os.path.join(work_dir, 'mvs_results/aggregate_3d/aggregate_3d_dsm.tif'))
        finally:
            shutil.rmtree(work_dir)
        out[name + '_aoi'] = np.array(json.dumps(scene.aoi))
        out[name + '_res'] = np.float64(res)
        out[name + '_fused_ply'] = ply_bytes
        out[name + '_utm'] = utm_pts
        out[name + '_utm_color'] = utm_color
        out[name + '_comments'] = np.array(json.dumps(list(comments)))
        out[name + '_dsm'] = dsm
        print(name, 'points', pts.shape[0], 'grid', dsm.shape, 'nan', np.isnan(dsm).mean(), comments)
    path = os.path.join(HERE, 'reference_golden_3d.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    main()
