"""Drop-in for the reference's lib/latlonalt_enu_converter.py (:36-45): same names, arguments (degrees, metres;
scalars or arrays) and return order, computed by the exact float64 CUDA chain (pymap3d 1.7.15 formulas)."""
from ._geo_common import run3


def latlonalt_to_enu(lat, lon, alt, lat0, lon0, alt0):
    e, n, u = run3('vs_geodetic_to_enu', lat, lon, alt, float(lat0), float(lon0), float(alt0))
    return e, n, u


def enu_to_latlonalt(e, n, u, lat0, lon0, alt0):
    lat, lon, alt = run3('vs_enu_to_geodetic', e, n, u, float(lat0), float(lon0), float(alt0))
    return lat, lon, alt
