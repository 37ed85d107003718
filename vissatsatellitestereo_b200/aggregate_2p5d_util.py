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
from ._native import lib, check
from .engine import _ptr
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


def check_zone_and_hemisphere(depth_map, inv_proj_mat, aoi_dict, item=''):
    """lib/latlon_utm_converter.py:39-48: the reference takes the UTM zone and the hemisphere of a whole view from its
    FIRST valid point.  The fused kernel takes them from aoi.json (include/vissat_b200.h, vs_aoi); the two agree for an AOI
    inside one zone.  This makes the other case loud instead of silently different: the first valid pixel (raster order,
    aggregate_2p5d_util.py:92-93) is unprojected on the host, converted with the exact chain, and its zone / hemisphere
    must be aoi.json's.  Returns False if the view has no valid pixel."""
    from .lib.latlonalt_enu_converter import enu_to_latlonalt
    from .lib.latlon_utm_converter import latlon_to_zone_number
    flat = np.asarray(depth_map).reshape(-1)
    pos = np.flatnonzero(flat > 0)
    if pos.size == 0:
        return False
    width = depth_map.shape[1]
    M = np.asarray(inv_proj_mat, dtype=np.float64).reshape(4, 4)
    for idx in pos[:64]:          # the first valid pixel whose unprojection is finite (:92)
        row, col = divmod(int(idx), width)
        X = M.dot(np.array([col, row, 1.0, float(flat[idx])]))
        if X[3] != 0 and np.all(np.isfinite(X[:3] / X[3])):
            x, y, z = X[:3] / X[3]
            break
    else:
        return True
    lat0 = (aoi_dict['lat_min'] + aoi_dict['lat_max']) / 2.0
    lon0 = (aoi_dict['lon_min'] + aoi_dict['lon_max']) / 2.0
    lat, lon, _ = enu_to_latlonalt(float(x), float(y), float(z), lat0, lon0, aoi_dict['alt_min'])
    zone = latlon_to_zone_number(lat, lon)
    hemi = 'N' if lat >= 0 else 'S'
    if zone != int(aoi_dict['zone_number']) or hemi != aoi_dict['hemisphere']:
        raise _engine._native.VisSatError(
            'view {}: its first valid point lies in UTM zone {}{} (lat {:.6f}, lon {:.6f}) but aoi.json says {}{}; the '
            'reference would project this view in zone {}{} (lib/latlon_utm_converter.py:43-48), the fused kernel uses '
            'aoi.json -- an AOI that straddles a zone boundary or the equator is not supported'.format(
                item, zone, hemi, lat, lon, aoi_dict['zone_number'], aoi_dict['hemisphere'], zone, hemi))
    return True


def convert_depth_map_worker(work_dir, out_dir, item, depth_type, _state=None):
    """One view (:45-106).  Returns (per-view DSM device tensor, number of empty cells, stem), or None if the view was
    skipped.  Stand-alone form of what convert_depth_maps pipelines."""
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
    depth_map = np.ascontiguousarray(depth_map, dtype=np.float32)
    if not check_zone_and_hemisphere(depth_map, _state['mats'][img_name], aoi_dict, item):
        # the reference dies here (lat[0, 0] on an empty array, lib/latlon_utm_converter.py:43) inside a pool
        # worker whose exception is never fetched: the view silently produces no tif.  Same outcome, but logged.
        logging.warning('no valid depth pixel in {}: view skipped'.format(item))
        return None
    depth = torch.from_numpy(depth_map).to(eng.device, non_blocking=True)
    want_hm = _produce_dsm.write_previews
    height_map = torch.empty_like(depth) if want_hm else None
    eng.rasterize(depth, _state['mats'][img_name], height_map=height_map)
    dsm = eng.finalize(count_nan=True)
    stem = img_name[:-4]
    _write_view_outputs(out_dir, stem, dsm.cpu().numpy(), height_map.cpu().numpy() if want_hm else None, eng, aoi_dict)
    return dsm, eng.last_nan_count(), stem


class _ViewPipeline:
    """Views of one rank through the GPU with everything overlapped: while view v is in stages A+B, view v+1 is copied
    host -> device and the DSM (and height map) of view v-1 is copied back and written to disk by the I/O pool.
    The per-view DSMs are written straight into ONE preallocated (n_views, n_size, e_size) device stack, which is what
    aggregate_2p5d.run_fuse fuses -- no per-view tensors, no torch.stack copy (the reference re-reads the tifs instead,
    aggregate_2p5d.py:57-66)."""

    N_SLOTS = 3

    def __init__(self, eng, aoi_dict, mats, out_dir, depth_type, n_views, io_pool):
        self.eng, self.aoi, self.mats, self.out_dir, self.depth_type = eng, aoi_dict, mats, out_dir, depth_type
        self.pool = io_pool
        dev = eng.device
        self.stack = torch.empty((max(n_views, 1), eng.n_size, eng.e_size), dtype=torch.float32, device=dev)
        self.nan_counts = torch.zeros(max(n_views, 1), dtype=torch.int64, device=dev)
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(device=dev) for _ in range(3))
        self.slots = [None] * self.N_SLOTS
        self.n = 0                  # planes of the stack in use
        self.views = []             # (plane index, stem)
        self.jobs = []
        self.want_hm = _produce_dsm.write_previews

    def _slot(self, i, shape):
        s = self.slots[i]
        if s is None or s['shape'] != shape:
            dev = self.eng.device
            s = {'shape': shape,
                 'h_depth': torch.empty(shape, dtype=torch.float32).pin_memory(),
                 'd_depth': torch.empty(shape, dtype=torch.float32, device=dev),
                 'd_hm': torch.empty(shape, dtype=torch.float32, device=dev) if self.want_hm else None,
                 'h_hm': torch.empty(shape, dtype=torch.float32).pin_memory() if self.want_hm else None,
                 'h_dsm': torch.empty((self.eng.n_size, self.eng.e_size), dtype=torch.float32).pin_memory(),
                 'busy': None}
            self.slots[i] = s
        return s

    def submit(self, item, depth_host):
        """Returns (plane index, stem) or None if the item is skipped."""
        idx = item.rfind('.{}.bin'.format(self.depth_type))
        if idx == -1:
            logging.info('something funny is happening: {}'.format(item))      # :66-69
            return None
        img_name = item[:idx]
        logging.info('converting depth map to dsm: {}'.format(img_name))
        if img_name not in self.mats:
            raise KeyError('no inv_proj_mats.txt row for {}'.format(img_name))
        depth_host = np.ascontiguousarray(depth_host, dtype=np.float32)
        if not check_zone_and_hemisphere(depth_host, self.mats[img_name], self.aoi, item):
            logging.warning('no valid depth pixel in {}: view skipped'.format(item))
            return None
        eng, i = self.eng, self.n
        slot = self._slot(i % self.N_SLOTS, tuple(depth_host.shape))
        if slot['busy'] is not None:
            slot['busy'].result()                 # the writer of the view that used this slot before is done with it
        slot['h_depth'].numpy()[...] = depth_host
        with torch.cuda.stream(self.s_in):
            slot['d_depth'].copy_(slot['h_depth'], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.s_in)
        self.s_comp.wait_event(ev_in)
        with torch.cuda.stream(self.s_comp):
            eng.rasterize(slot['d_depth'], self.mats[img_name], height_map=slot['d_hm'])
            check(lib.vs_grid_finalize(eng.ctx.handle, _ptr(eng.keygrid), eng.e_size, eng.n_size, _ptr(self.stack[i]),
                                       eng.simd_lanes, _ptr(self.nan_counts[i:i + 1]), _stream_of(self.s_comp)),
                  'vs_grid_finalize')
            ev_comp = torch.cuda.Event()
            ev_comp.record(self.s_comp)
        self.s_out.wait_event(ev_comp)
        self.s_in.wait_event(ev_comp)              # the next upload into this slot's device buffer comes after its use
        with torch.cuda.stream(self.s_out):
            slot['h_dsm'].copy_(self.stack[i], non_blocking=True)
            if self.want_hm:
                slot['h_hm'].copy_(slot['d_hm'], non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(self.s_out)
        stem = img_name[:-4]

        def write(slot=slot, ev_out=ev_out, stem=stem):
            ev_out.synchronize()
            _write_view_outputs(self.out_dir, stem, slot['h_dsm'].numpy(), slot['h_hm'].numpy() if self.want_hm else None,
                                self.eng, self.aoi)
        slot['busy'] = self.pool.submit(write)
        self.jobs.append(slot['busy'])
        self.views.append((i, stem))
        self.n += 1
        return i, stem

    def finish(self):
        for j in self.jobs:
            j.result()                            # surfaces I/O errors
        torch.cuda.current_stream(self.eng.device).wait_stream(self.s_comp)
        torch.cuda.synchronize(self.eng.device)
        counts = self.nan_counts[:self.n].cpu().tolist()
        return self.stack[:self.n], [(i, counts[i], stem) for i, stem in self.views]


def _stream_of(s):
    import ctypes as C
    return C.c_void_p(s.cuda_stream)


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
    mats = load_inv_proj_mats(mvs_dir)
    n_io = max(1, min(max_processes, len(my_items), 8))
    with ThreadPoolExecutor(n_io) as pool, ThreadPoolExecutor(max(1, min(n_io, 4))) as writers:
        pipe = _ViewPipeline(eng, aoi_dict, mats, out_dir, depth_type, len(my_items), writers)
        # host threads read ahead (at most 2 * n_io depth maps wait in host memory); the GPU work is enqueued by this
        # thread, in sorted order, on the pipeline's streams
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
            depth_host = fut.result()
            refill()
            pipe.submit(item, depth_host)
        stack, views = pipe.finish()
    _RESULTS[os.path.abspath(out_dir)] = {'engine': eng, 'aoi': aoi_dict, 'stack': stack, 'views': views,
                                          'n_items': len(all_items), 'rank': rank, 'world': world}


if __name__ == '__main__':
    pass
