"""The slice of the reference's stereo_pipeline.py that surrounds the aggregation path:
`write_aoi` (:185-226) and the two step methods that call into the path (`run_aggregate_2p5d` :479-497,
`run_aggregate_3d` :461-477).  Every other step (image cropping, SfM, MVS: COLMAP and RPC code) is out of scope and
asking for it is an error.

`write_aoi` goes through `utm.to_latlon` (utm==0.4.2) in the reference; that package is not in /root/reference, so
its series is restated below from the published algorithm it implements (Snyder, USGS PP-1395, eqs. 8-17…8-25 and
3-24/7-19: footpoint latitude + truncated series in D).  No reference-run vector exists for it; the tests anchor it on the
package's own known-value vectors (metre-rounded) and on the exact inverse of the data path (agreement < 5e-8 deg).  It
only fixes the ENU origin stored in aoi.json.
"""
import json
import logging
import math
import os
import time

from . import aggregate_2p5d, aggregate_3d

_K0 = 0.9996
_E = 0.00669438
_E2 = _E * _E
_E3 = _E2 * _E
_E_P2 = _E / (1.0 - _E)
_SQRT_E = math.sqrt(1 - _E)
_e1 = (1 - _SQRT_E) / (1 + _SQRT_E)
_M1 = 1 - _E / 4 - 3 * _E2 / 64 - 5 * _E3 / 256
_P2 = 3. / 2 * _e1 - 27. / 32 * _e1 ** 3 + 269. / 512 * _e1 ** 5
_P3 = 21. / 16 * _e1 ** 2 - 55. / 32 * _e1 ** 4
_P4 = 151. / 96 * _e1 ** 3 - 417. / 128 * _e1 ** 5
_P5 = 1097. / 512 * _e1 ** 4
_R = 6378137


def utm_to_latlon(easting, northing, zone_number, northern):
    """utm.to_latlon(easting, northing, zone_number, northern=...) -> (lat, lon) in degrees (footpoint-latitude
    series; agrees with the PROJ inverse used on the data path to ~1e-8 deg)."""
    x = easting - 500000
    y = northing if northern else northing - 10000000
    mu = y / _K0 / (_R * _M1)
    p_rad = mu + _P2 * math.sin(2 * mu) + _P3 * math.sin(4 * mu) + _P4 * math.sin(6 * mu) + _P5 * math.sin(8 * mu)
    p_sin, p_cos = math.sin(p_rad), math.cos(p_rad)
    p_tan = p_sin / p_cos
    t2 = p_tan * p_tan
    t4 = t2 * t2
    ep_sin = 1 - _E * p_sin * p_sin
    n = _R / math.sqrt(ep_sin)
    r = (1 - _E) / ep_sin
    c = _E_P2 * p_cos ** 2
    c2 = c * c
    d = x / (n * _K0)
    d2 = d * d
    d3, d4 = d2 * d, d2 * d2
    d5, d6 = d4 * d, d4 * d2
    latitude = p_rad - (p_tan / r) * (d2 / 2 - d4 / 24 * (5 + 3 * t2 + 10 * c - 4 * c2 - 9 * _E_P2) +
                                      d6 / 720 * (61 + 90 * t2 + 298 * c + 45 * t4 - 252 * _E_P2 - 3 * c2))
    longitude = (d - d3 / 6 * (1 + 2 * t2 + c) +
                 d5 / 120 * (5 - 2 * c + 28 * t2 - 3 * c2 + 8 * _E_P2 + 24 * t4)) / p_cos
    return math.degrees(latitude), math.degrees(longitude) + (zone_number - 1) * 6 - 180 + 3


def write_aoi(config):
    """stereo_pipeline.py:185-226: config['bounding_box'] (UTM) + alt range -> <work_dir>/aoi.json; returns the dict."""
    bbx_utm = config['bounding_box']
    zone_number = bbx_utm['zone_number']
    hemisphere = bbx_utm['hemisphere']
    ul_easting = bbx_utm['ul_easting']
    ul_northing = bbx_utm['ul_northing']
    lr_easting = ul_easting + bbx_utm['width']
    lr_northing = ul_northing - bbx_utm['height']
    corners = [(ul_easting, ul_northing), (lr_easting, ul_northing), (lr_easting, lr_northing),
               (ul_easting, lr_northing)]
    latlon = [utm_to_latlon(e, n, zone_number, hemisphere == 'N') for (e, n) in corners]
    lats = [p[0] for p in latlon]
    lons = [p[1] for p in latlon]
    aoi_dict = {'zone_number': zone_number,
                'hemisphere': hemisphere,
                'ul_easting': ul_easting,
                'ul_northing': ul_northing,
                'lr_easting': lr_easting,
                'lr_northing': lr_northing,
                'width': bbx_utm['width'],
                'height': bbx_utm['height'],
                'lat_min': min(lats),
                'lat_max': max(lats),
                'lon_min': min(lons),
                'lon_max': max(lons),
                'alt_min': config['alt_min'],
                'alt_max': config['alt_max']}
    with open(os.path.join(config['work_dir'], 'aoi.json'), 'w') as fp:
        json.dump(aoi_dict, fp, indent=2)
    return aoi_dict


_IN_SCOPE = ('aggregate_2p5d', 'aggregate_3d')


class StereoPipeline(object):
    """Runs the aggregation steps of a reference config file; `steps_to_run` entries outside the path must be false."""

    def __init__(self, config_file):
        with open(config_file) as fp:
            self.config = json.load(fp)
        os.makedirs(os.path.join(self.config['work_dir'], 'logs'), exist_ok=True)

    def write_aoi(self):
        return write_aoi(self.config)

    def run(self):
        steps = self.config.get('steps_to_run', {})
        other = [k for k, v in steps.items() if v and k not in _IN_SCOPE]
        if other:
            raise NotImplementedError('steps outside the aggregation path are not part of this library: {}'
                                      .format(', '.join(sorted(other))))
        self.write_aoi()
        per_step_time = []
        for name, fn in (('aggregate_2p5d', self.run_aggregate_2p5d), ('aggregate_3d', self.run_aggregate_3d)):
            if steps.get(name, False):
                t0 = time.time()
                fn()
                per_step_time.append((True, name, (time.time() - t0) / 60.0))
            else:
                per_step_time.append((False, name, 0.0))
        with open(os.path.join(self.config['work_dir'], 'runtime.txt'), 'w') as fp:       # :170-183
            fp.write('step_name, status, duration (minutes)\n')
            total = 0.0
            for (has_run, step_name, duration) in per_step_time:
                if has_run:
                    fp.write('{}, success, {}\n'.format(step_name, duration))
                else:
                    fp.write('{}, skipped\n'.format(step_name))
                total += duration
            fp.write('\ntotal: {} minutes\n'.format(total))

    def run_aggregate_3d(self):
        aggregate_3d.run_fuse(self.config['work_dir'])
        logging.info('3D aggregation done')

    def run_aggregate_2p5d(self):
        max_processes = self.config.get('aggregate_max_processes', -1)
        aggregate_2p5d.run_fuse(self.config['work_dir'], max_processes=max_processes)
        logging.info('2.5D aggregation done')
