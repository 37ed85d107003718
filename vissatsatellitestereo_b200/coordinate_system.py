"""Drop-in for the reference's coordinate_system.py (:41-64): ENU <-> (lat, lon, alt) about the AOI origin
(bbox centre, alt_min) read from <work_dir>/aoi.json."""
import json
import os

from .lib.latlonalt_enu_converter import latlonalt_to_enu, enu_to_latlonalt


def _origin(work_dir):
    with open(os.path.join(work_dir, 'aoi.json')) as fp:
        bbx = json.load(fp)
    lat0 = (bbx['lat_min'] + bbx['lat_max']) / 2.0
    lon0 = (bbx['lon_min'] + bbx['lon_max']) / 2.0
    alt0 = bbx['alt_min']
    return lat0, lon0, alt0


def local_to_global(work_dir, xx, yy, zz):
    lat0, lon0, alt0 = _origin(work_dir)
    return enu_to_latlonalt(xx, yy, zz, lat0, lon0, alt0)


def global_to_local(work_dir, xx, yy, zz):
    lat0, lon0, alt0 = _origin(work_dir)
    return latlonalt_to_enu(xx, yy, zz, lat0, lon0, alt0)
