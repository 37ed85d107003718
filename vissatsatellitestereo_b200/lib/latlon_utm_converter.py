"""Drop-in for the reference's lib/latlon_utm_converter.py (:39-63), including the function name typo
`latlon_to_eastnorh`.  PROJ 6.2's +proj=utm (extended transverse Mercator, order 6) on the GPU in float64."""
import numpy as np

from ._geo_common import run2


def latlon_to_zone_number(latitude, longitude):
    """utm 0.4.2 zone rule (the reference takes only the zone number from utm.from_latlon, :48)."""
    if 56 <= latitude < 64 and 3 <= longitude < 12:
        return 32
    if 72 <= latitude <= 84 and longitude >= 0:
        if longitude <= 9:
            return 31
        elif longitude <= 21:
            return 33
        elif longitude <= 33:
            return 35
        elif longitude <= 42:
            return 37
    return int((longitude + 180) / 6) + 1


def latlon_to_eastnorh(lat, lon):
    # assume all the points are either on north or south hemisphere (:41)
    assert (np.all(lat >= 0) or np.all(lat < 0))
    south = not (lat[0, 0] >= 0)                                    # :43-46
    zone_number = latlon_to_zone_number(lat[0, 0], lon[0, 0])       # :48
    east, north = run2('vs_geodetic_to_utm', lat, lon, int(zone_number), 1 if south else 0)
    return east, north


def eastnorth_to_latlon(east, north, zone_number, hemisphere):
    south = hemisphere != 'N'                                       # :56-59
    lat, lon = run2('vs_utm_to_geodetic', east, north, int(zone_number), 1 if south else 0)
    return lat, lon
