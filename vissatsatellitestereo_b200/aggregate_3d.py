"""Drop-in for the tail of the reference's aggregate_3d.py (:54-83): COLMAP's fused ENU point cloud -> UTM point
cloud (`aggregate_3d.ply`) -> DSM (`aggregate_3d_dsm.tif` / `.jpg`).

Same kernels as the 2.5D path, second consumer: the exact float64 ENU -> geodetic -> UTM chain
(`vs_enu_to_geodetic`, `vs_geodetic_to_utm`) and the point rasteriser (`vs_points_rasterize` + `vs_grid_finalize64`).
The points stay on the device between the conversion and the rasterisation; the host copy is only made for the PLY.

`fuse()` (aggregate_3d.py:43-51) shells out to the COLMAP binary; that program is outside this repository.  It is run
when `colmap` is on PATH; otherwise an existing `colmap/mvs/fused.ply` is used, and its absence is an error.
"""
import json
import logging
import os
import shutil
import subprocess

import numpy as np
import torch

from . import engine, produce_dsm
from ._native import lib, check
from .lib.latlon_utm_converter import latlon_to_zone_number
from .lib.ply_np_converter import ply2np, np2ply


# the unit of max_depth_error is now in meter
def fuse(colmap_dir):
    out_ply = os.path.join(colmap_dir, 'mvs/fused.ply')
    exe = shutil.which('colmap')
    if exe is None:
        if os.path.exists(out_ply):
            logging.info('colmap not on PATH; using existing {}'.format(out_ply))
            return
        raise FileNotFoundError('colmap is not on PATH and {} does not exist'.format(out_ply))
    cmd = [exe, 'stereo_fusion', '--workspace_path', os.path.join(colmap_dir, 'mvs'), '--output_path', out_ply,
           '--input_type', 'geometric', '--StereoFusion.min_num_pixels', '4', '--StereoFusion.max_reproj_error', '2',
           '--StereoFusion.max_depth_error', '1.0', '--StereoFusion.max_normal_error', '10']
    subprocess.run(cmd, check=True)


def enu_points_to_utm_device(work_dir, points_enu, device=None):
    """coordinate_system.local_to_global (:41-51) + latlon_to_eastnorh (lib/latlon_utm_converter.py:39-52) for an
    (N, 3) float64 ENU array, on the device.  Returns a device (N, 3) float64 tensor (east, north, alt)."""
    ctx, dev = engine.default_context(device)
    with open(os.path.join(work_dir, 'aoi.json')) as fp:
        bbx = json.load(fp)
    lat0 = (bbx['lat_min'] + bbx['lat_max']) / 2.0
    lon0 = (bbx['lon_min'] + bbx['lon_max']) / 2.0
    alt0 = bbx['alt_min']
    if not torch.is_tensor(points_enu):
        points_enu = torch.from_numpy(np.ascontiguousarray(np.asarray(points_enu, dtype=np.float64)))
    pts = points_enu.to(device=dev, dtype=torch.float64)
    n = pts.shape[0]
    out = torch.empty((n, 3), dtype=torch.float64, device=dev)
    if n == 0:
        return out
    cols = pts.t().contiguous()                       # (3, N): e, n, u planes
    geo = torch.empty((3, n), dtype=torch.float64, device=dev)
    st = engine._stream(dev)
    p = engine._ptr
    check(lib.vs_enu_to_geodetic(ctx.handle, p(cols[0]), p(cols[1]), p(cols[2]), n, lat0, lon0, alt0,
                                 p(geo[0]), p(geo[1]), p(geo[2]), st), 'vs_enu_to_geodetic')
    lat = geo[0]
    # assume all the points are either on north or south hemisphere (lib/latlon_utm_converter.py:41)
    n_north = int((lat >= 0).sum().item())
    assert n_north == n or n_north == 0
    first = geo[:2, 0].cpu().numpy()
    south = not (first[0] >= 0)
    zone_number = latlon_to_zone_number(float(first[0]), float(first[1]))
    en = torch.empty((2, n), dtype=torch.float64, device=dev)
    check(lib.vs_geodetic_to_utm(ctx.handle, p(geo[0]), p(geo[1]), n, int(zone_number), 1 if south else 0,
                                 p(en[0]), p(en[1]), st), 'vs_geodetic_to_utm')
    out[:, 0] = en[0]
    out[:, 1] = en[1]
    out[:, 2] = geo[2]
    return out


def run_fuse(work_dir):
    fuse(os.path.join(work_dir, 'colmap'))

    os.makedirs(os.path.join(work_dir, 'mvs_results'), exist_ok=True)
    out_dir = os.path.join(work_dir, 'mvs_results/aggregate_3d')
    os.makedirs(out_dir, exist_ok=True)

    points, color, comments = ply2np(os.path.join(work_dir, 'colmap/mvs/fused.ply'))

    # ENU -> lat, lon, alt -> UTM on the device; the (E, N, alt) array stays there for the rasteriser
    points_utm = enu_points_to_utm_device(work_dir, points)

    with open(os.path.join(work_dir, 'aoi.json')) as fp:
        aoi_dict = json.load(fp)
    comment_1 = 'projection: UTM {}{}'.format(aoi_dict['zone_number'], aoi_dict['hemisphere'])
    comments = [comment_1, ]
    np2ply(points_utm.cpu().numpy(), os.path.join(out_dir, 'aggregate_3d.ply'), color=color, comments=comments,
           use_double=True)

    # write dsm to tif
    tif_to_write = os.path.join(out_dir, 'aggregate_3d_dsm.tif')
    jpg_to_write = os.path.join(out_dir, 'aggregate_3d_dsm.jpg')
    produce_dsm.produce_dsm_from_points(work_dir, points_utm, tif_to_write, jpg_to_write)
