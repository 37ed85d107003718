"""Drop-in for the reference's aggregate_2p5d_util.py (:45-146): depth maps -> per-view DSM GeoTIFFs.

Same entry points and on-disk outputs (<out_dir>/dsm_tif/<stem>.tif, dsm_jpg/<stem>.jpg, dsm_img_grid/<stem>.jpg).
What changes is where the work happens: the reference forks one process per view and does the arithmetic in
numpy/pymap3d/pyproj; here one process owns the GPU, `max_processes` only bounds the host I/O threads (CUDA
contexts do not survive fork), and every view goes through libvissat_b200 (stage A + B).  Under torchrun the
sorted view list is split into contiguous blocks, one per rank.
"""
import collections
import logging
import os
import shutil
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import engine as _engine
from . import produce_dsm as _produce_dsm
from .colmap.read_dense import read_array
from .lib.dsm_util import write_dsm_tif
from .visualization.plot_height_map import plot_height_map

# per-out_dir cache of what convert_depth_maps left on the device, picked up by aggregate_2p5d.run_fuse
_RESULTS = {}


def load_inv_proj_mats(mvs_dir):
    """aggregate_2p5d_util.py:54-61: one 'name m00 ... m33' line per image."""
    inv_proj_mats = {}
    with open(os.path.join(mvs_dir, 'inv_proj_mats.txt')) as fp:
        for line in fp.readlines():
            tmp = line.split(' ')
            if len(tmp) < 17:
                continue
            inv_proj_mats[tmp[0]] = np.array([float(tmp[i]) for i in range(1, 17)]).reshape((4, 4))
    return inv_proj_mats


def _make_engine(work_dir):
    import json
    with open(os.path.join(work_dir, 'aoi.json')) as fp:
        aoi_dict = json.load(fp)
    return _engine.DsmEngine(aoi_dict, _produce_dsm.e_resolution, _produce_dsm.n_resolution), aoi_dict


def _ensure_dirs(out_dir):
    for subdir in [out_dir, os.path.join(out_dir, 'dsm_tif'), os.path.join(out_dir, 'dsm_jpg'),
                   os.path.join(out_dir, 'dsm_img_grid')]:
        os.makedirs(subdir, exist_ok=True)


def _write_view_outputs(out_dir, stem, dsm, height_map, eng, aoi_dict):
    tif_to_write = os.path.join(out_dir, 'dsm_tif', stem + '.tif')
    write_dsm_tif(dsm, tif_to_write, (aoi_dict['ul_easting'], aoi_dict['ul_northing'], eng.e_resolution, eng.n_resolution),
                  (aoi_dict['zone_number'], aoi_dict['hemisphere']), nodata_val=-10000)
    if _produce_dsm.write_previews:
        plot_height_map(np.clip(dsm, aoi_dict['alt_min'], aoi_dict['alt_max']),
                        os.path.join(out_dir, 'dsm_jpg', stem + '.jpg'), save_cbar=True)
        if height_map is not None and not np.isnan(height_map).all():
            min_val, max_val = np.nanpercentile(height_map, [1, 99])                      # :104-106
            plot_height_map(np.clip(height_map, min_val, max_val), os.path.join(out_dir, 'dsm_img_grid', stem + '.jpg'))


def convert_depth_map_worker(work_dir, out_dir, item, depth_type, _state=None):
    """One view (:45-106).  Returns the per-view DSM as a device tensor, or None if the view was skipped."""
    _ensure_dirs(out_dir)
    mvs_dir = os.path.join(work_dir, 'colmap/mvs')
    if _state is None:
        eng, aoi_dict = _make_engine(work_dir)
        _state = {'eng': eng, 'aoi': aoi_dict, 'mats': load_inv_proj_mats(mvs_dir)}
    eng, aoi_dict = _state['eng'], _state['aoi']
    idx = item.rfind('.{}.bin'.format(depth_type))
    if idx == -1:
        logging.info('something funny is happening: {}'.format(item))      # :66-69
        return None
    img_name = item[:idx]
    logging.info('converting depth map to dsm: {}'.format(img_name))
    if img_name not in _state['mats']:
        raise KeyError('no inv_proj_mats.txt row for {}'.format(img_name))
    depth_map = _state.get('depth_host')
    if depth_map is None:
        depth_map = read_array(os.path.join(mvs_dir, 'stereo/depth_maps', item))
    depth = torch.from_numpy(np.ascontiguousarray(depth_map, dtype=np.float32)).to(eng.device, non_blocking=True)
    want_hm = _produce_dsm.write_previews
    height_map = torch.empty_like(depth) if want_hm else None
    stats_on = eng.collect_stats
    eng.collect_stats = True
    eng.rasterize(depth, _state['mats'][img_name], height_map=height_map)
    dsm = eng.finalize(count_nan=True)
    st = eng.stats()
    eng.collect_stats = stats_on
    if st['valid'] == 0:
        # the reference dies here (lat[0, 0] on an empty array, lib/latlon_utm_converter.py:43) inside a pool
        # worker whose exception is never fetched: the view silently produces no tif.  Same outcome, but logged.
        logging.warning('no valid depth pixel in {}: view skipped'.format(item))
        return None
    stem = img_name[:-4]
    _write_view_outputs(out_dir, stem, dsm.cpu().numpy(), height_map.cpu().numpy() if want_hm else None, eng, aoi_dict)
    return dsm, eng.last_nan_count(), stem


def split_big_list(big_list, num_small_lists):
    cnt = len(big_list)
    indices = np.array_split(np.arange(cnt, dtype=np.int32), num_small_lists)
    small_lists = []
    for sub in indices:
        if sub.size > 0:
            small_lists.append(big_list[sub[0]:sub[-1] + 1])
    return small_lists


def convert_depth_maps(work_dir, out_dir, depth_type, max_processes=-1):
    mvs_dir = os.path.join(work_dir, 'colmap/mvs')
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    if rank == 0 and os.path.exists(out_dir):
        shutil.rmtree(out_dir)                                              # :127-128
    if distributed:
        dist.barrier()
    _ensure_dirs(out_dir)

    # all items to be converted (:131-135)
    depth_dir = os.path.join(mvs_dir, 'stereo/depth_maps')
    all_items = [item for item in sorted(os.listdir(depth_dir)) if depth_type in item]
    logging.info('{} to be processed...'.format(len(all_items)))
    if max_processes <= 0:
        max_processes = os.cpu_count() or 1

    from .distributed import split_views
    a, b = split_views(len(all_items), world)[rank]
    my_items = all_items[a:b]
    eng, aoi_dict = _make_engine(work_dir)
    state = {'eng': eng, 'aoi': aoi_dict, 'mats': load_inv_proj_mats(mvs_dir)}
    results = []
    n_io = max(1, min(max_processes, len(my_items), 8))
    with ThreadPoolExecutor(n_io) as pool:
        # host threads read ahead (at most 2 * n_io depth maps wait in host memory); the GPU work itself is serialised
        # on this process's stream
        def load(item):
            if item.rfind('.{}.bin'.format(depth_type)) == -1:
                return None
            return read_array(os.path.join(depth_dir, item))
        todo = iter(my_items)
        pending = collections.deque()

        def refill():
            while len(pending) < 2 * n_io:
                nxt = next(todo, None)
                if nxt is None:
                    return
                pending.append((nxt, pool.submit(load, nxt)))
        refill()
        while pending:
            item, fut = pending.popleft()
            state['depth_host'] = fut.result()
            refill()
            r = convert_depth_map_worker(work_dir, out_dir, item, depth_type, _state=state)
            if r is not None:
                results.append(r)
    state.pop('depth_host', None)
    _RESULTS[os.path.abspath(out_dir)] = {'engine': eng, 'aoi': aoi_dict, 'views': results, 'n_items': len(all_items),
                                          'rank': rank, 'world': world}


if __name__ == '__main__':
    pass
