"""Deterministic synthetic inputs for the aggregate_2p5d path.

Builds what the upstream steps of the reference would leave in a work_dir for this step:
``aoi.json`` (stereo_pipeline.py:185-226), one 4x4 ``inv_proj_mats.txt`` row per view
(the inverse of P4 = [K[R|t]; depth_min*(0,0,1,-min_z)], reparam_depth.py:131-141) and COLMAP
``<name>.png.geometric.bin`` float32 depth maps (colmap/read_dense.py:36-51), invalid pixels = -1e20.

Input synthesis is not part of the measured path.  torch is used so that full-size benchmark inputs
can be generated directly on the GPU; small parity cases are generated on the CPU and the same arrays
are given to both the oracle and the CUDA path.

`geo` is any object exposing the reference's converter functions ``eastnorth_to_latlon(east, north,
zone_number, hemisphere)`` (lib/latlon_utm_converter.py:55); the product's own module or, in tests,
the oracle's.
"""
import json
import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch

EXPLORER_UL_E = 354052.3651180889      # aoi_config/MVS3DM_Explorer.json:7-8
EXPLORER_UL_N = 6182702.10540914
EXPLORER_ZONE = 21
EXPLORER_HEMI = 'S'


@dataclass
class SynthConfig:
    name: str
    n_views: int
    height: int          # depth-map rows
    width: int           # depth-map cols
    e_size: int          # grid cols
    n_size: int          # grid rows
    res: float           # grid resolution (m)
    gsd: float           # ground sample distance of the cameras (m/px)
    max_off_nadir_deg: float = 25.0
    invalid_iid: float = 0.05
    invalid_blocks: float = 0.05
    random_offsets: bool = False     # C3: each view covers part of the grid at a random offset
    fuse: bool = True
    config_id: int = 0
    alt_min: float = -30.0
    alt_max: float = 120.0


# the five BASELINE.json configs (grid for C4 is the SURVEY §8d assumption)
CONFIGS = {
    'C1': SynthConfig('C1', 8, 1024, 1024, 512, 512, 0.5, 0.25, config_id=1),
    'C2': SynthConfig('C2', 50, 2048, 2048, 2048, 2048, 0.3, 0.3, config_id=2),
    'C3': SynthConfig('C3', 200, 4096, 4096, 8192, 8192, 0.3, 0.35, random_offsets=True, config_id=3),
    'C4': SynthConfig('C4', 100, 4096, 4096, 2048, 2048, 1.0, 0.5, fuse=False, config_id=4),
    'C5': SynthConfig('C5', 64, 4096, 4096, 4096, 4096, 0.3, 0.3, max_off_nadir_deg=5.0,
                      invalid_iid=0.5, invalid_blocks=0.0, config_id=5),
}


def scaled(cfg, views=None, depth=None, grid=None, name=None):
    """Down-scaled copy of a config (same geometry recipe) for CI-sized parity tests."""
    c = SynthConfig(**cfg.__dict__)
    if views is not None:
        c.n_views = views
    if depth is not None:
        f = depth / c.height
        c.height = depth
        c.width = int(round(cfg.width * f))
        c.gsd = cfg.gsd  # keep GSD: footprint shrinks with the image
    if grid is not None:
        f = grid / c.e_size
        c.e_size = grid
        c.n_size = int(round(cfg.n_size * f))
    if name is not None:
        c.name = name
    return c


def make_aoi(cfg, geo, ul_e=EXPLORER_UL_E, ul_n=EXPLORER_UL_N, zone=EXPLORER_ZONE, hemi=EXPLORER_HEMI):
    """aoi.json dict as stereo_pipeline.py:210-223 writes it. width/height are (size-0.5)*res so that
    int(width/res)+1 == size robustly (produce_dsm.py:54-55)."""
    width = (cfg.e_size - 0.5) * cfg.res
    height = (cfg.n_size - 0.5) * cfg.res
    lr_e = ul_e + width
    lr_n = ul_n - height
    ce = np.array([ul_e, lr_e, lr_e, ul_e], dtype=np.float64).reshape(-1, 1)
    cn = np.array([ul_n, ul_n, lr_n, lr_n], dtype=np.float64).reshape(-1, 1)
    lat, lon = geo.eastnorth_to_latlon(ce, cn, zone, hemi)
    lat = np.asarray(lat, dtype=np.float64).ravel()
    lon = np.asarray(lon, dtype=np.float64).ravel()
    return {'zone_number': zone, 'hemisphere': hemi,
            'ul_easting': ul_e, 'ul_northing': ul_n, 'lr_easting': lr_e, 'lr_northing': lr_n,
            'width': width, 'height': height,
            'lat_min': float(lat.min()), 'lat_max': float(lat.max()),
            'lon_min': float(lon.min()), 'lon_max': float(lon.max()),
            'alt_min': cfg.alt_min, 'alt_max': cfg.alt_max}


class Terrain:
    """h(x, y) in ENU metres: sum of sinusoids (sigma ~ 8 m) + flat-roofed boxes, as a 0.5 m raster."""

    def __init__(self, cfg, device='cpu', spacing=0.5):
        rng = np.random.default_rng(1000 * cfg.config_id + 999)
        ext_x = cfg.e_size * cfg.res * 0.5 + 200.0
        ext_y = cfg.n_size * cfg.res * 0.5 + 200.0
        self.spacing = spacing
        self.x0 = -ext_x
        self.y0 = -ext_y
        nx = int(2 * ext_x / spacing) + 1
        ny = int(2 * ext_y / spacing) + 1
        xs = torch.arange(nx, device=device, dtype=torch.float32) * spacing + self.x0
        ys = torch.arange(ny, device=device, dtype=torch.float32) * spacing + self.y0
        h = torch.zeros((ny, nx), device=device, dtype=torch.float32)
        for octave in range(6):
            wl = 600.0 / (2 ** octave)
            amp = 9.0 / (1.6 ** octave)
            th = rng.uniform(0, 2 * math.pi)
            ph = rng.uniform(0, 2 * math.pi)
            kx, ky = math.cos(th) * 2 * math.pi / wl, math.sin(th) * 2 * math.pi / wl
            h += amp * torch.sin(kx * xs[None, :] + ky * ys[:, None] + ph)
        mid = 0.5 * (cfg.alt_min + cfg.alt_max) - 20.0
        h += mid
        n_boxes = int(4e-4 * (2 * ext_x) * (2 * ext_y) / 4) + 4
        n_boxes = min(n_boxes, 4000)
        bx = rng.uniform(-ext_x, ext_x, n_boxes)
        by = rng.uniform(-ext_y, ext_y, n_boxes)
        bw = rng.uniform(8, 50, n_boxes)
        bh = rng.uniform(8, 50, n_boxes)
        bz = rng.uniform(5, 40, n_boxes)
        for i in range(n_boxes):
            i0 = max(int((bx[i] - bw[i] / 2 - self.x0) / spacing), 0)
            i1 = min(int((bx[i] + bw[i] / 2 - self.x0) / spacing), nx)
            j0 = max(int((by[i] - bh[i] / 2 - self.y0) / spacing), 0)
            j1 = min(int((by[i] + bh[i] / 2 - self.y0) / spacing), ny)
            if i1 > i0 and j1 > j0:
                base = h[j0:j1, i0:i1].mean()
                h[j0:j1, i0:i1] = base + float(bz[i])
        self.h = torch.clamp(h, cfg.alt_min + 5.0, cfg.alt_max - 5.0)
        self.nx, self.ny = nx, ny

    def __call__(self, x, y):
        i = torch.clamp(((x - self.x0) / self.spacing).round().long(), 0, self.nx - 1)
        j = torch.clamp(((y - self.y0) / self.spacing).round().long(), 0, self.ny - 1)
        return self.h[j, i].to(torch.float64)


def make_camera(cfg, view, alt0):
    """Satellite-like perspective camera in the AOI's ENU frame (z = 0 at alt0 = alt_min).
    Returns (M = inv(P4) float64 4x4, P4)."""
    rng = np.random.default_rng(1000 * cfg.config_id + view)
    dist = 6.0e5
    off = math.radians(rng.uniform(0.0, cfg.max_off_nadir_deg))
    az = rng.uniform(0.0, 2 * math.pi)
    ext_x = cfg.e_size * cfg.res
    ext_y = cfg.n_size * cfg.res
    tx = ty = 0.0
    if cfg.random_offsets:
        fx_ = max(ext_x - cfg.width * cfg.gsd, 0.0) / 2
        fy_ = max(ext_y - cfg.height * cfg.gsd, 0.0) / 2
        tx = rng.uniform(-fx_, fx_)
        ty = rng.uniform(-fy_, fy_)
    z_mid = 0.5 * (cfg.alt_max - cfg.alt_min) - 20.0
    target = np.array([tx, ty, z_mid])
    d = np.array([math.sin(off) * math.cos(az), math.sin(off) * math.sin(az), math.cos(off)])
    C = target + dist * d
    zc = (target - C) / np.linalg.norm(target - C)            # forward
    north = np.array([0.0, 1.0, 0.0])
    yc = -(north - zc * np.dot(north, zc))                     # image 'down' = south-ish
    yc /= np.linalg.norm(yc)
    xc = np.cross(yc, zc)
    R = np.vstack((xc, yc, zc))
    t = -R @ C
    f = dist / cfg.gsd
    K = np.array([[f, 0.0, cfg.width / 2.0], [0.0, f, cfg.height / 2.0], [0.0, 0.0, 1.0]])
    P3 = K @ np.hstack((R, t.reshape(3, 1)))
    min_z = -20.0                                              # reparam_depth.py:98-99 (margin 20 m)
    depth_min = dist - 2000.0
    P4 = np.vstack((P3, depth_min * np.array([[0.0, 0.0, 1.0, -min_z]])))   # reparam_depth.py:104,141
    M = np.linalg.inv(P4)
    return M, P4


def make_depth_map(cfg, view, M, terrain, device='cpu'):
    """float32 (H,W) depth map whose unprojection (aggregate_2p5d_util.py:86-90) lands on the terrain
    (+ N(0,0.3 m) height noise); invalid pixels are -1e20."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 * cfg.config_id + view)
    Mt = torch.as_tensor(M, dtype=torch.float64, device=dev)
    col = torch.arange(cfg.width, dtype=torch.float64, device=dev)[None, :]
    row = torch.arange(cfg.height, dtype=torch.float64, device=dev)[:, None]
    a = [Mt[i, 0] * col + Mt[i, 1] * row + Mt[i, 2] for i in range(4)]
    z_mid = 0.5 * (cfg.alt_max - cfg.alt_min) - 20.0
    h = torch.full((cfg.height, cfg.width), z_mid, dtype=torch.float64, device=dev)
    alt0 = cfg.alt_min
    for _ in range(3):                                       # fixed-point ray/terrain intersection
        d = (h * a[3] - a[2]) / (Mt[2, 3] - h * Mt[3, 3])
        w = a[3] + Mt[3, 3] * d
        x = (a[0] + Mt[0, 3] * d) / w
        y = (a[1] + Mt[1, 3] * d) / w
        h = terrain(x.float(), y.float()) - alt0               # ENU up = altitude - alt_min
    h = h + 0.3 * torch.randn(h.shape, generator=gen, device=dev, dtype=torch.float32).double()
    d = (h * a[3] - a[2]) / (Mt[2, 3] - h * Mt[3, 3])
    depth = d.to(torch.float32)
    invalid = torch.rand(depth.shape, generator=gen, device=dev) < cfg.invalid_iid
    if cfg.invalid_blocks > 0:
        rng = np.random.default_rng(1000 * cfg.config_id + view + 500000)
        area = cfg.invalid_blocks * cfg.height * cfg.width
        bs = max(int(min(cfg.height, cfg.width) / 16), 2)
        for _ in range(max(int(area / (bs * bs)), 1)):
            r0 = int(rng.integers(0, max(cfg.height - bs, 1)))
            c0 = int(rng.integers(0, max(cfg.width - bs, 1)))
            invalid[r0:r0 + bs, c0:c0 + bs] = True
    depth = torch.where(invalid | (depth <= 0), torch.full_like(depth, -1e20), depth)
    return depth


def view_name(view):
    return '{:04d}.png'.format(view)


@dataclass
class SynthScene:
    cfg: SynthConfig
    aoi: dict
    names: list = field(default_factory=list)
    mats: list = field(default_factory=list)       # list of np.float64 (4,4)
    depths: list = field(default_factory=list)     # list of torch.float32 (H,W) on `device`


def make_scene(cfg, geo, device='cpu', views=None):
    aoi = make_aoi(cfg, geo)
    terrain = Terrain(cfg, device=device)
    scene = SynthScene(cfg=cfg, aoi=aoi)
    for v in (range(cfg.n_views) if views is None else views):
        M, _ = make_camera(cfg, v, aoi['alt_min'])
        scene.names.append(view_name(v))
        scene.mats.append(M)
        scene.depths.append(make_depth_map(cfg, v, M, terrain, device=device))
    return scene


def write_colmap_array(path, array):
    """COLMAP dense array file (colmap/read_dense.py:36-51 reads it): 'W&H&C&' + float32 payload."""
    array = np.ascontiguousarray(np.asarray(array, dtype=np.float32))
    assert array.ndim == 2
    h, w = array.shape
    with open(path, 'wb') as fid:
        fid.write('{}&{}&{}&'.format(w, h, 1).encode('ascii'))
        fid.write(array.tobytes())


def write_work_dir(scene, work_dir, depth_type='geometric'):
    """Lay the scene out as the reference expects it under work_dir (aggregate_2p5d_util.py:53-64)."""
    mvs_dir = os.path.join(work_dir, 'colmap', 'mvs')
    depth_dir = os.path.join(mvs_dir, 'stereo', 'depth_maps')
    os.makedirs(depth_dir, exist_ok=True)
    with open(os.path.join(work_dir, 'aoi.json'), 'w') as fp:
        json.dump(scene.aoi, fp, indent=2)
    with open(os.path.join(mvs_dir, 'inv_proj_mats.txt'), 'w') as fp:
        for name, M in zip(scene.names, scene.mats):
            fp.write('{} {}\n'.format(name, ' '.join(repr(float(x)) for x in M.reshape(-1))))
    for name, depth in zip(scene.names, scene.depths):
        write_colmap_array(os.path.join(depth_dir, '{}.{}.bin'.format(name, depth_type)),
                           depth.detach().cpu().numpy())
