"""Drop-in for the reference's produce_dsm.py (:41-88): UTM points or a height grid -> GeoTIFF DSM (+ preview).

Module globals e_resolution / n_resolution are read at call time, as in the reference (rebinding them changes
the grid resolution).  The rasterisation (per-cell nanmax, 3x3 hole fill, float32 cast, 3x3 median) runs on the
GPU; file writing is host I/O."""
import json
import os

import numpy as np

from . import engine
from .lib.dsm_util import write_dsm_tif
from .visualization.plot_height_map import plot_height_map

e_resolution = 0.5  # 0.5 meters per pixel
n_resolution = 0.5

write_previews = True   # the reference always writes the .jpg previews; set False to skip them


def _load_aoi(work_dir):
    with open(os.path.join(work_dir, 'aoi.json')) as fp:
        return json.load(fp)


# points is in UTM
def produce_dsm_from_points(work_dir, points, tif_to_write, jpg_to_write=None):
    aoi_dict = _load_aoi(work_dir)
    ul_e = aoi_dict['ul_easting']
    ul_n = aoi_dict['ul_northing']
    e_size = int(aoi_dict['width'] / e_resolution) + 1
    n_size = int(aoi_dict['height'] / n_resolution) + 1
    # lib/proj_to_grid.py:41-81 + cv2.medianBlur(dsm.astype(np.float32), 3), both on the device
    _, dsm = engine.proj_to_grid_device(points, ul_e, ul_n, e_resolution, n_resolution, e_size, n_size, blur=True)
    dsm = dsm.cpu().numpy()
    write_dsm_tif(dsm, tif_to_write, (ul_e, ul_n, e_resolution, n_resolution),
                  (aoi_dict['zone_number'], aoi_dict['hemisphere']), nodata_val=-10000)
    if jpg_to_write is not None and write_previews:
        plot_height_map(np.clip(dsm, aoi_dict['alt_min'], aoi_dict['alt_max']), jpg_to_write, save_cbar=True)
    return (ul_e, ul_n, e_size, n_size, e_resolution, n_resolution)


def produce_dsm_from_height(work_dir, height, tif_to_write, jpg_to_write=None):
    aoi_dict = _load_aoi(work_dir)
    ul_e = aoi_dict['ul_easting']
    ul_n = aoi_dict['ul_northing']
    n_size, e_size = height.shape[:2]
    write_dsm_tif(height, tif_to_write, (ul_e, ul_n, e_resolution, n_resolution),
                  (aoi_dict['zone_number'], aoi_dict['hemisphere']), nodata_val=-10000)
    if jpg_to_write is not None and write_previews:
        plot_height_map(np.clip(height, aoi_dict['alt_min'], aoi_dict['alt_max']), jpg_to_write, save_cbar=True)
    return (ul_e, ul_n, e_size, n_size, e_resolution, n_resolution)
