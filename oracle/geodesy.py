"""Oracle geodesy: numpy float64 restatement of the third-party chain the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  No reference-run vectors exist for this file (pymap3d / pyproj / utm
are absent here); it is pinned instead to an independent 50-digit evaluation of the same maps
(tests/golden/make_geodesy_mp.py, tests/test_geodesy_pin.py: <= 5e-9 m over the benchmark AOIs) and anchored on
published known answers.

Reference call sites being restated:
  * lib/latlonalt_enu_converter.py:36-45  -> pymap3d 1.7.15 geodetic2enu / enu2geodetic
  * lib/latlon_utm_converter.py:39-63     -> utm 0.4.2 zone rule + pyproj 2.4.0 (PROJ 6.2)
                                            Proj(proj='utm', ellps='WGS84', zone, south)
                                            = Poder/Engsager extended transverse Mercator, order 6
  * coordinate_system.py:41-64            -> origin = bbox centre, alt_min
The operation ORDER of each upstream formula is kept (it determines the float64 rounding).
"""
import numpy as np

# ---------------------------------------------------------------------------------------------
# pymap3d 1.7.15 — Ellipsoid('wgs84')
# ---------------------------------------------------------------------------------------------
WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.2572235630
WGS84_B = WGS84_A * (1.0 - WGS84_F)


def get_radius_normal(lat_rad):
    """pymap3d.ecef.get_radius_normal: prime-vertical radius N(lat)."""
    a, b = WGS84_A, WGS84_B
    return a ** 2 / np.sqrt(a ** 2 * np.cos(lat_rad) ** 2 + b ** 2 * np.sin(lat_rad) ** 2)


def geodetic2ecef(lat, lon, alt):
    """pymap3d.ecef.geodetic2ecef (degrees in)."""
    lat = np.radians(lat)
    lon = np.radians(lon)
    N = get_radius_normal(lat)
    x = (N + alt) * np.cos(lat) * np.cos(lon)
    y = (N + alt) * np.cos(lat) * np.sin(lon)
    z = (N * (WGS84_B / WGS84_A) ** 2 + alt) * np.sin(lat)
    return x, y, z


def ecef2geodetic(x, y, z):
    """pymap3d.ecef.ecef2geodetic (You, 2000 closed form); degrees out."""
    a, b = WGS84_A, WGS84_B
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    E = np.sqrt(a ** 2 - b ** 2)
    # eqn. 4a
    u = np.sqrt(0.5 * (r ** 2 - E ** 2) + 0.5 * np.sqrt((r ** 2 - E ** 2) ** 2 + 4 * E ** 2 * z ** 2))
    Q = np.hypot(x, y)
    huE = np.hypot(u, E)
    # eqn. 4b
    with np.errstate(divide='ignore', invalid='ignore'):
        Beta = np.arctan(huE / u * z / np.hypot(x, y))
    # eqn. 13
    eps = ((b * u - a * huE + E ** 2) * np.sin(Beta)) / (a * huE * 1 / np.cos(Beta) - E ** 2 * np.cos(Beta))
    Beta = Beta + eps
    lat = np.arctan(a / b * np.tan(Beta))
    lon = np.arctan2(y, x)
    # eqn. 7
    alt = np.hypot(z - b * np.sin(Beta), Q - a * np.cos(Beta))
    with np.errstate(invalid='ignore'):
        inside = x ** 2 / a ** 2 + y ** 2 / a ** 2 + z ** 2 / b ** 2 < 1
    alt = np.where(inside, -alt, alt)
    return np.degrees(lat), np.degrees(lon), alt


def enu2uvw(east, north, up, lat0, lon0):
    """pymap3d.enu.enu2uvw (degrees in)."""
    lat0 = np.radians(lat0)
    lon0 = np.radians(lon0)
    t = np.cos(lat0) * up - np.sin(lat0) * north
    w = np.sin(lat0) * up + np.cos(lat0) * north
    u = np.cos(lon0) * t - np.sin(lon0) * east
    v = np.sin(lon0) * t + np.cos(lon0) * east
    return u, v, w


def uvw2enu(u, v, w, lat0, lon0):
    """pymap3d.enu.uvw2enu (degrees in)."""
    lat0 = np.radians(lat0)
    lon0 = np.radians(lon0)
    t = np.cos(lon0) * u + np.sin(lon0) * v
    East = -np.sin(lon0) * u + np.cos(lon0) * v
    Up = np.cos(lat0) * t + np.sin(lat0) * w
    North = -np.sin(lat0) * t + np.cos(lat0) * w
    return East, North, Up


def enu2geodetic(e, n, u, lat0, lon0, h0):
    """pymap3d.enu.enu2geodetic = enu2ecef then ecef2geodetic."""
    x0, y0, z0 = geodetic2ecef(lat0, lon0, h0)
    dx, dy, dz = enu2uvw(e, n, u, lat0, lon0)
    return ecef2geodetic(x0 + dx, y0 + dy, z0 + dz)


def geodetic2enu(lat, lon, h, lat0, lon0, h0):
    """pymap3d.enu.geodetic2enu = ECEF difference then uvw2enu."""
    x1, y1, z1 = geodetic2ecef(lat, lon, h)
    x2, y2, z2 = geodetic2ecef(lat0, lon0, h0)
    return uvw2enu(x1 - x2, y1 - y2, z1 - z2, lat0, lon0)


# reference wrappers, lib/latlonalt_enu_converter.py:36-45
def latlonalt_to_enu(lat, lon, alt, lat0, lon0, alt0):
    return geodetic2enu(lat, lon, alt, lat0, lon0, alt0)


def enu_to_latlonalt(e, n, u, lat0, lon0, alt0):
    return enu2geodetic(e, n, u, lat0, lon0, alt0)


# ---------------------------------------------------------------------------------------------
# utm 0.4.2 — zone rule (lib/latlon_utm_converter.py:48 uses only the zone number)
# ---------------------------------------------------------------------------------------------
def utm_zone_number(latitude, longitude):
    """utm.conversion.latlon_to_zone_number."""
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


# ---------------------------------------------------------------------------------------------
# PROJ 6.2 etmerc (what +proj=utm evaluates), order 6
# ---------------------------------------------------------------------------------------------
PROJ_A = 6378137.0
PROJ_RF = 298.257223563
PROJ_K0 = 0.9996
ETMERC_ORDER = 6


def _etmerc_setup():
    f0 = 1.0 / PROJ_RF
    es = 2 * f0 - f0 * f0
    # etmerc.cpp setup(): "f = P->es / (1 + sqrt(1 - P->es))"
    f = es / (1 + np.sqrt(1 - es))
    n = f / (2 - f)
    np_ = n
    cgb = np.zeros(6)
    cbg = np.zeros(6)
    utg = np.zeros(6)
    gtu = np.zeros(6)
    cgb[0] = n * (2 + n * (-2 / 3.0 + n * (-2 + n * (116 / 45.0 + n * (26 / 45.0 + n * (-2854 / 675.0))))))
    cbg[0] = n * (-2 + n * (2 / 3.0 + n * (4 / 3.0 + n * (-82 / 45.0 + n * (32 / 45.0 + n * (4642 / 4725.0))))))
    np_ *= n
    cgb[1] = np_ * (7 / 3.0 + n * (-8 / 5.0 + n * (-227 / 45.0 + n * (2704 / 315.0 + n * (2323 / 945.0)))))
    cbg[1] = np_ * (5 / 3.0 + n * (-16 / 15.0 + n * (-13 / 9.0 + n * (904 / 315.0 + n * (-1522 / 945.0)))))
    np_ *= n
    cgb[2] = np_ * (56 / 15.0 + n * (-136 / 35.0 + n * (-1262 / 105.0 + n * (73814 / 2835.0))))
    cbg[2] = np_ * (-26 / 15.0 + n * (34 / 21.0 + n * (8 / 5.0 + n * (-12686 / 2835.0))))
    np_ *= n
    cgb[3] = np_ * (4279 / 630.0 + n * (-332 / 35.0 + n * (-399572 / 14175.0)))
    cbg[3] = np_ * (1237 / 630.0 + n * (-12 / 5.0 + n * (-24832 / 14175.0)))
    np_ *= n
    cgb[4] = np_ * (4174 / 315.0 + n * (-144838 / 6237.0))
    cbg[4] = np_ * (-734 / 315.0 + n * (109598 / 31185.0))
    np_ *= n
    cgb[5] = np_ * (601676 / 22275.0)
    cbg[5] = np_ * (444337 / 155925.0)

    np_ = n * n
    Qn = PROJ_K0 / (1 + n) * (1 + np_ * (1 / 4.0 + np_ * (1 / 64.0 + np_ / 256.0)))
    utg[0] = n * (-0.5 + n * (2 / 3.0 + n * (-37 / 96.0 + n * (1 / 360.0 + n * (81 / 512.0 + n * (-96199 / 604800.0))))))
    gtu[0] = n * (0.5 + n * (-2 / 3.0 + n * (5 / 16.0 + n * (41 / 180.0 + n * (-127 / 288.0 + n * (7891 / 37800.0))))))
    utg[1] = np_ * (-1 / 48.0 + n * (-1 / 15.0 + n * (437 / 1440.0 + n * (-46 / 105.0 + n * (1118711 / 3870720.0)))))
    gtu[1] = np_ * (13 / 48.0 + n * (-3 / 5.0 + n * (557 / 1440.0 + n * (281 / 630.0 + n * (-1983433 / 1935360.0)))))
    np_ *= n
    utg[2] = np_ * (-17 / 480.0 + n * (37 / 840.0 + n * (209 / 4480.0 + n * (-5569 / 90720.0))))
    gtu[2] = np_ * (61 / 240.0 + n * (-103 / 140.0 + n * (15061 / 26880.0 + n * (167603 / 181440.0))))
    np_ *= n
    utg[3] = np_ * (-4397 / 161280.0 + n * (11 / 504.0 + n * (830251 / 7257600.0)))
    gtu[3] = np_ * (49561 / 161280.0 + n * (-179 / 168.0 + n * (6601661 / 7257600.0)))
    np_ *= n
    utg[4] = np_ * (-4583 / 161280.0 + n * (108847 / 3991680.0))
    gtu[4] = np_ * (34729 / 80640.0 + n * (-3418889 / 1995840.0))
    np_ *= n
    utg[5] = np_ * (-20648693 / 638668800.0)
    gtu[5] = np_ * (212378941 / 319334400.0)
    # Zb: origin northing for phi0 = 0 (UTM): gatg(cbg, 0) == 0 -> Zb = -Qn*(0 + clens(gtu, 0)) = -0.0
    Z = _gatg(cbg, np.float64(0.0))
    Zb = -Qn * (Z + _clens(gtu, 2 * Z))
    return dict(cgb=cgb, cbg=cbg, utg=utg, gtu=gtu, Qn=Qn, Zb=Zb, n=n)


def _gatg(p, B):
    cos_2B = 2 * np.cos(2 * B)
    h1 = p[5]
    h2 = 0.0
    h = 0.0
    for k in range(4, -1, -1):
        h = -h2 + cos_2B * h1 + p[k]
        h2 = h1
        h1 = h
    return B + h * np.sin(2 * B)


def _clens(a, arg_r):
    cos_arg_r = np.cos(arg_r)
    r = 2 * cos_arg_r
    hr1 = 0.0
    hr = a[5]
    for k in range(4, -1, -1):
        hr2 = hr1
        hr1 = hr
        hr = -hr2 + r * hr1 + a[k]
    return np.sin(arg_r) * hr


def _clenS(a, arg_r, arg_i):
    sin_arg_r = np.sin(arg_r)
    cos_arg_r = np.cos(arg_r)
    sinh_arg_i = np.sinh(arg_i)
    cosh_arg_i = np.cosh(arg_i)
    r = 2 * cos_arg_r * cosh_arg_i
    i = -2 * sin_arg_r * sinh_arg_i
    hi1 = 0.0
    hr1 = 0.0
    hi = 0.0
    hr = a[5]
    for k in range(4, -1, -1):
        hr2 = hr1
        hi2 = hi1
        hr1 = hr
        hi1 = hi
        hr = -hr2 + r * hr1 - i * hi1 + a[k]
        hi = -hi2 + i * hr1 + r * hi1
    r = sin_arg_r * cosh_arg_i
    i = cos_arg_r * sinh_arg_i
    R = r * hr - i * hi
    I = r * hi + i * hr
    return R, I


def _asinhy(x):
    y = np.abs(x)
    y = np.log1p(y * (1 + y / (np.hypot(1.0, y) + 1)))
    return np.where(x < 0, -y, y)


_ETMERC = None


def etmerc_consts():
    global _ETMERC
    if _ETMERC is None:
        _ETMERC = _etmerc_setup()
    return _ETMERC


def utm_lam0(zone_number):
    """PROJ utm setup: lam0 = (zone - 1 + .5) * pi / 30 - pi   (zone is 1-based in the string)."""
    zone = zone_number - 1
    return (zone + .5) * np.pi / 30. - np.pi


def utm_forward(lat_deg, lon_deg, zone_number, south):
    """pyproj.Proj(proj='utm', ellps='WGS84', zone, south)(lon, lat) -> (east, north)."""
    C = etmerc_consts()
    DG2RAD = np.radians(1.0)
    lat_deg = np.asarray(lat_deg, dtype=np.float64)
    lon_deg = np.asarray(lon_deg, dtype=np.float64)
    phi = DG2RAD * lat_deg
    lam = DG2RAD * lon_deg - utm_lam0(zone_number)
    Cn = _gatg(C['cbg'], phi)
    Ce = lam
    sin_Cn = np.sin(Cn)
    cos_Cn = np.cos(Cn)
    sin_Ce = np.sin(Ce)
    cos_Ce = np.cos(Ce)
    Cn = np.arctan2(sin_Cn, cos_Ce * cos_Cn)
    Ce = np.arctan2(sin_Ce * cos_Cn, np.hypot(sin_Cn, cos_Cn * cos_Ce))
    Ce = _asinhy(np.tan(Ce))
    dCn, dCe = _clenS(C['gtu'], 2 * Cn, 2 * Ce)
    Cn = Cn + dCn
    Ce = Ce + dCe
    y = C['Qn'] * Cn + C['Zb']
    x = C['Qn'] * Ce
    bad = np.abs(Ce) > 2.623395162778
    x0 = 500000.0
    y0 = 10000000.0 if south else 0.0
    east = PROJ_A * x + x0
    north = PROJ_A * y + y0
    east = np.where(bad, np.inf, east)
    north = np.where(bad, np.inf, north)
    return east, north


def utm_inverse(east, north, zone_number, south):
    """pyproj.Proj(...)(east, north, inverse=True) -> (lon, lat) ; returned here as (lat, lon)."""
    C = etmerc_consts()
    east = np.asarray(east, dtype=np.float64)
    north = np.asarray(north, dtype=np.float64)
    x0 = 500000.0
    y0 = 10000000.0 if south else 0.0
    ra = 1.0 / PROJ_A
    # inv_prepare: xy = (xy * to_meter - x0) * ra
    Ce = (east - x0) * ra
    Cn = (north - y0) * ra
    Cn = (Cn - C['Zb']) / C['Qn']
    Ce = Ce / C['Qn']
    dCn, dCe = _clenS(C['utg'], 2 * Cn, 2 * Ce)
    Cn = Cn + dCn
    Ce = Ce + dCe
    Ce = np.arctan(np.sinh(Ce))
    sin_Cn = np.sin(Cn)
    cos_Cn = np.cos(Cn)
    sin_Ce = np.sin(Ce)
    cos_Ce = np.cos(Ce)
    Ce = np.arctan2(sin_Ce, cos_Ce * cos_Cn)
    Cn = np.arctan2(sin_Cn * cos_Ce, np.hypot(sin_Ce, cos_Ce * cos_Cn))
    phi = _gatg(C['cgb'], Cn)
    lam = Ce + utm_lam0(zone_number)
    RAD2DG = np.degrees(1.0)
    return phi * RAD2DG, lam * RAD2DG


# reference wrappers, lib/latlon_utm_converter.py:39-63
def latlon_to_eastnorh(lat, lon):
    assert (np.all(lat >= 0) or np.all(lat < 0))
    south = not (lat[0, 0] >= 0)
    zone_number = utm_zone_number(lat[0, 0], lon[0, 0])
    return utm_forward(lat, lon, zone_number, south)


def eastnorth_to_latlon(east, north, zone_number, hemisphere):
    south = hemisphere != 'N'
    return utm_inverse(east, north, zone_number, south)


# utm 0.4.2 to_latlon is used by stereo_pipeline.py:202 (write_aoi) for the AOI corners only.  We use
# the PROJ inverse for the synthetic aoi.json (the two agree to ~1e-9 deg, irrelevant for the origin).
def enu_origin_from_aoi(aoi):
    """coordinate_system.py:45-47."""
    lat0 = (aoi['lat_min'] + aoi['lat_max']) / 2.0
    lon0 = (aoi['lon_min'] + aoi['lon_max']) / 2.0
    alt0 = aoi['alt_min']
    return lat0, lon0, alt0
