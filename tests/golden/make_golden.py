#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What runs verbatim from the reference (imported, not copied):
  * lib/proj_to_grid.py:proj_to_grid               (needs `np.int`; we restore the removed alias)
  * colmap/read_dense.py:read_array
  * aggregate_2p5d_util.py:convert_depth_maps / convert_depth_map_worker   (inline, no fork pool)
  * produce_dsm.py:produce_dsm_from_points / produce_dsm_from_height
  * lib/dsm_util.py:write_dsm_tif / read_dsm_tif    (on a tiny in-memory GDAL stand-in)
  * aggregate_2p5d.py:run_fuse                      (the whole step, incl. the fusion block :65-81)
  * lib/ply_np_converter.py + lib/plyfile.py

Third-party modules that are absent from this image (no network) get stand-ins, installed in
sys.modules before the reference is imported:
  numpy_groupies.aggregate(func='nanmax')  -> independent dict-of-lists implementation below
  pymap3d / pyproj / utm                   -> oracle.geodesy (restated algorithms; parity unpinned)
  osgeo.gdal/gdal_array/osr                -> in-memory dataset objects
  matplotlib / imageio / visualization.*   -> no-op previews (jpg output is out of scope)
The stand-ins only replace the *imports*; every line of the reference files listed above executes.
"""
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)

from oracle import geodesy  # noqa: E402
from vissatsatellitestereo_b200 import synthetic as S  # noqa: E402


# ------------------------------------------------------------------ stand-ins for absent imports
def install_shims():
    if not hasattr(np, 'int'):
        np.int = int          # removed alias used at lib/proj_to_grid.py:42-43,53

    # numpy_groupies: independent, deliberately naive implementation of aggregate(..., 'nanmax')
    npg = types.ModuleType('numpy_groupies')

    def aggregate(group_idx, a, func='sum', fill_value=0, size=None):
        assert func == 'nanmax'
        size = int(group_idx.max()) + 1 if size is None else size
        best = {}
        for g, v in zip(group_idx.tolist(), a.tolist()):
            if v != v:
                continue
            if g not in best or v > best[g]:
                best[g] = v
        out = np.full(size, fill_value, dtype=np.float64)
        for g, v in best.items():
            out[g] = v
        return out

    npg.aggregate = aggregate
    sys.modules['numpy_groupies'] = npg

    pm = types.ModuleType('pymap3d')
    pm.geodetic2enu = geodesy.geodetic2enu
    pm.enu2geodetic = geodesy.enu2geodetic
    sys.modules['pymap3d'] = pm

    utm = types.ModuleType('utm')
    utm.from_latlon = lambda lat, lon: (None, None, geodesy.utm_zone_number(float(lat), float(lon)), None)
    sys.modules['utm'] = utm

    pyproj = types.ModuleType('pyproj')

    class Proj:
        def __init__(self, proj=None, ellps=None, zone=None, south=False):
            assert proj == 'utm' and ellps == 'WGS84'
            self.zone, self.south = zone, south

        def __call__(self, x, y, inverse=False):
            if inverse:
                lat, lon = geodesy.utm_inverse(x, y, self.zone, self.south)
                return lon, lat
            return geodesy.utm_forward(y, x, self.zone, self.south)

    pyproj.Proj = Proj
    sys.modules['pyproj'] = pyproj

    # ---- GDAL stand-in: a dataset is a dict kept in an in-process registry when released
    _FILES = {}
    osgeo = types.ModuleType('osgeo')
    gdal = types.ModuleType('osgeo.gdal')
    gdal_array = types.ModuleType('osgeo.gdal_array')
    osr = types.ModuleType('osgeo.osr')
    GDT_Float32 = 6
    gdal.DCAP_RASTER = 'DCAP_RASTER'
    gdal.DMD_EXTENSIONS = 'DMD_EXTENSIONS'

    class Band:
        def __init__(self, ds):
            self.ds = ds
            self.DataType = GDT_Float32

        def WriteArray(self, arr, xoff=0, yoff=0):
            self.ds.state['image'] = np.array(arr, dtype=np.float32)

        def SetNoDataValue(self, v):
            self.ds.state['nodata'] = float(v)

        def FlushCache(self):
            pass

        def ReadAsArray(self):
            return self.ds.state['image']

        def GetNoDataValue(self):
            return self.ds.state['nodata']

    class Dataset:
        def __init__(self, path, state=None, writable=False):
            self.path, self.writable = path, writable
            self.state = state if state is not None else {}
            self.RasterCount = 1

        RasterXSize = property(lambda self: self.state['image'].shape[1])
        RasterYSize = property(lambda self: self.state['image'].shape[0])

        def GetRasterBand(self, i):
            assert i == 1
            return Band(self)

        def SetGeoTransform(self, g):
            self.state['geo'] = tuple(g)

        def SetProjection(self, p):
            self.state['proj'] = p

        def SetMetadata(self, m):
            self.state['meta'] = dict(m)

        def GetGeoTransform(self):
            return self.state['geo']

        def GetProjection(self):
            return self.state['proj']

        def GetMetadata(self):
            return self.state['meta']

        def __del__(self):
            if self.writable:
                _FILES[self.path] = self.state
                open(self.path, 'wb').close()      # so that os.listdir / os.path.exists see the .tif

    class Driver:
        def GetMetadataItem(self, key):
            return {'DCAP_RASTER': 'YES', 'DMD_EXTENSIONS': 'tif tiff'}.get(key)

        def Create(self, path, w, h, bands, dtype):
            assert bands == 1 and dtype == GDT_Float32
            return Dataset(path, {'image': np.zeros((h, w), np.float32)}, writable=True)

    gdal.GetDriverCount = lambda: 1
    gdal.GetDriver = lambda i: Driver()
    gdal.Open = lambda path: Dataset(path, _FILES[path])
    gdal_array.GDALTypeCodeToNumericTypeCode = lambda code: {GDT_Float32: np.float32}[code]
    gdal_array.NumericTypeCodeToGDALTypeCode = lambda t: {np.float32: GDT_Float32}[t]

    class SpatialReference:
        def SetProjCS(self, name):
            self.name = name

        def SetWellKnownGeogCS(self, name):
            pass

        def SetUTM(self, zone, north):
            pass

        def ExportToWkt(self):
            return 'PROJCS["{}",GEOGCS["WGS 84"]]'.format(self.name)

    osr.SpatialReference = SpatialReference
    osgeo.gdal, osgeo.gdal_array, osgeo.osr = gdal, gdal_array, osr
    for name, mod in (('osgeo', osgeo), ('osgeo.gdal', gdal), ('osgeo.gdal_array', gdal_array), ('osgeo.osr', osr)):
        sys.modules[name] = mod

    # ---- previews: out of scope (lossy jpg); keep the call sites alive
    vis = types.ModuleType('visualization')
    vis.__path__ = []
    phm = types.ModuleType('visualization.plot_height_map')

    def plot_height_map(height_map, out_file, maskout=None, save_cbar=False, force_range=None):
        np.save(out_file + '.shape.npy', np.array(height_map.shape[:2]))

    phm.plot_height_map = plot_height_map
    sys.modules['visualization'] = vis
    sys.modules['visualization.plot_height_map'] = phm
    imageio = types.ModuleType('imageio')
    imageio.imread = lambda path: np.zeros(tuple(np.load(path + '.shape.npy')) + (3,), np.uint8)
    imageio.imwrite = lambda path, im: None
    sys.modules['imageio'] = imageio

    # the reference's pool forks workers whose exceptions are swallowed; run them inline instead so
    # failures surface (aggregate_2p5d_util.py:141-146)
    import multiprocessing

    class InlinePool:
        def __init__(self, n):
            pass

        def apply_async(self, fn, args=()):
            fn(*args)

        def close(self):
            pass

        def join(self):
            pass

    multiprocessing.Pool = InlinePool


def run_reference_case(name, cfg, resolution, out):
    """Build a work_dir, run the reference's run_fuse on it, collect arrays."""
    import produce_dsm                                   # reference module
    import aggregate_2p5d                                # reference module
    from lib.dsm_util import read_dsm_tif                # reference module
    from colmap.read_dense import read_array             # reference module

    produce_dsm.e_resolution = resolution               # module globals read at call time (:41-42)
    produce_dsm.n_resolution = resolution
    scene = S.make_scene(cfg, geodesy, device='cpu')
    work_dir = tempfile.mkdtemp(prefix='vissat_golden_')
    try:
        S.write_work_dir(scene, work_dir)
        aggregate_2p5d.run_fuse(work_dir, max_processes=1)
        tif_dir = os.path.join(work_dir, 'colmap/mvs/dsm/dsm_tif')
        per_view = []
        for item in sorted(os.listdir(tif_dir)):
            if item.endswith('.tif'):
                dsm, meta = read_dsm_tif(os.path.join(tif_dir, item))
                per_view.append(dsm)
        fused, meta = read_dsm_tif(os.path.join(work_dir, 'mvs_results/aggregate_2p5d/aggregate_2p5d_dsm.tif'))
        depth0 = read_array(os.path.join(work_dir, 'colmap/mvs/stereo/depth_maps',
                                         scene.names[0] + '.geometric.bin'))
        assert np.array_equal(depth0, scene.depths[0].numpy())
        out[name + '_aoi'] = np.array(json.dumps(scene.aoi))
        out[name + '_res'] = np.float64(resolution)
        out[name + '_mats'] = np.stack(scene.mats)
        out[name + '_depths'] = np.stack([d.numpy() for d in scene.depths])
        out[name + '_per_view'] = np.stack(per_view)
        out[name + '_fused'] = fused
        out[name + '_geo'] = np.array(meta['geo'])
        print(name, 'views', len(per_view), 'grid', fused.shape, 'per-view nan', np.isnan(per_view[0]).mean(),
              'fused nan', np.isnan(fused).mean())
    finally:
        shutil.rmtree(work_dir)


def main():
    install_shims()
    sys.path.insert(0, REF)
    from lib.proj_to_grid import proj_to_grid            # reference function
    from colmap.read_dense import read_array             # reference function

    out = {}
    # ---- G1: proj_to_grid on adversarial random points (duplicates, OOB, NaN values, big holes)
    rng = np.random.default_rng(20261017)
    cases = []
    for k, (xs, ys, npts) in enumerate([(37, 23, 1500), (64, 64, 900), (5, 3, 40), (1, 1, 5), (18, 2, 60)]):
        xoff, yoff, res = 354052.3651180889, 6182702.10540914, (0.5, 0.3, 1.0, 0.5, 0.25)[k]
        pts = np.empty((npts, 3))
        pts[:, 0] = xoff + rng.uniform(-2 * res, (xs + 2) * res, npts)
        pts[:, 1] = yoff - rng.uniform(-2 * res, (ys + 2) * res, npts)
        pts[:, 2] = rng.normal(30, 10, npts)
        pts[rng.random(npts) < 0.05, 2] = np.nan
        if k == 1:                                        # leave a big empty region
            pts = pts[pts[:, 0] < xoff + 40 * res]
        # exact-on-edge points
        pts[:3, 0] = xoff + np.array([0.0, 1.0, 2.0]) * res
        pts[:3, 1] = yoff - np.array([0.0, 1.0, 2.0]) * res
        dsm = proj_to_grid(pts, xoff, yoff, res, res, xs, ys)
        out['ptg{}_points'.format(k)] = pts
        out['ptg{}_args'.format(k)] = np.array([xoff, yoff, res, res, xs, ys])
        out['ptg{}_dsm'.format(k)] = dsm
        cases.append(dsm.shape)
    print('proj_to_grid cases', cases)

    # ---- G2: read_array on a file with an awkward header
    tmp = tempfile.mkdtemp(prefix='vissat_golden_')
    arr = rng.normal(size=(7, 13)).astype(np.float32)
    path = os.path.join(tmp, 'a.bin')
    S.write_colmap_array(path, arr)
    got = read_array(path)
    assert got.shape == (7, 13)
    out['read_array_in'] = arr
    out['read_array_out'] = got
    with open(path, 'rb') as fp:
        out['read_array_file'] = np.frombuffer(fp.read(), dtype=np.uint8)
    shutil.rmtree(tmp)

    # ---- G3: the whole step through the reference's run_fuse
    c1 = S.scaled(S.CONFIGS['C1'], views=6, depth=192, grid=96, name='g_c1')
    run_reference_case('c1', c1, 0.5, out)
    c5 = S.scaled(S.CONFIGS['C5'], views=9, depth=160, grid=160, name='g_c5')
    run_reference_case('c5', c5, 0.3, out)
    c3 = S.scaled(S.CONFIGS['C3'], views=12, depth=128, grid=256, name='g_c3')
    run_reference_case('c3', c3, 0.3, out)

    np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **out)
    print('wrote', os.path.join(HERE, 'reference_golden.npz'),
          os.path.getsize(os.path.join(HERE, 'reference_golden.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
