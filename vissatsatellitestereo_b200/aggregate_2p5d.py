"""Drop-in for the reference's aggregate_2p5d.py (:45-113): `run_fuse(work_dir, max_processes=-1)`.

Outputs, as in the reference: colmap/mvs/dsm/{dsm_tif,dsm_jpg,dsm_img_grid}/<stem>.* per view and
mvs_results/aggregate_2p5d/{aggregate_2p5d_dsm.tif, aggregate_2p5d_dsm.jpg, aggregate_2p5d.ply}.
The robust fusion (:65-78) and the final 3x3 median (:81) run on the GPU; under torchrun the per-view DSM row
bands are exchanged with one NCCL all-to-all and each rank fuses its own rows (rank 0 writes the result).
"""
import json
import logging
import os

import numpy as np
import torch

from . import aggregate_2p5d_util as _util
from .aggregate_2p5d_util import convert_depth_maps
from .lib.dsm_util import read_dsm_tif
from .lib.ply_np_converter import np2ply
from .produce_dsm import produce_dsm_from_height


def _all_ranks_ok(err, device, what):
    """Under torchrun a rank that raised alone would leave the others inside the next collective: exchange a flag first
    and fail everywhere."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        if err is not None:
            raise err
        return
    flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if err is not None:
        raise err
    if int(flag.item()) != 0:
        raise RuntimeError('{} failed on another rank'.format(what))


def run_fuse(work_dir, max_processes=-1):
    # first convert depth maps
    dsm_dir = os.path.join(work_dir, 'colmap/mvs/dsm')
    err = None
    try:
        convert_depth_maps(work_dir, dsm_dir, depth_type='geometric', max_processes=max_processes)
    except Exception as e:          # noqa: BLE001 -- re-raised on every rank below
        err = e
    dev = torch.device('cuda', torch.cuda.current_device())
    _all_ranks_ok(err, dev, 'convert_depth_maps')
    res = _util._RESULTS.pop(os.path.abspath(dsm_dir))
    eng, rank, world = res['engine'], res['rank'], res['world']

    out_dir = os.path.join(work_dir, 'mvs_results/aggregate_2p5d')
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)

    # the per-view DSMs are still on the device, in sorted file order (:59), as planes of ONE stack written by stage B;
    # the reference re-reads the tifs here
    local = res['stack']
    for plane, n_nan, stem in res['views']:
        logging.info('dsm {} empty ratio: {} '.format(stem + '.tif', n_nan / (eng.n_size * eng.e_size)))
    if world > 1:
        import torch.distributed as dist
        from . import distributed as D
        counts = [torch.zeros(1, dtype=torch.int64, device=eng.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64, device=eng.device))
        band, _ = D.fuse_distributed(eng, local, [int(c.item()) for c in counts])
        fused = D.gather_bands(band, eng.n_size, eng.e_size)
        if rank != 0:
            return
    else:
        if local.shape[0] == 0:
            raise RuntimeError('no per-view DSM was produced')
        fused = eng.fuse_and_blur(local)
    all_dsm_mean_no_outliers = fused.cpu().numpy()

    # write tif
    tif_to_write = os.path.join(out_dir, 'aggregate_2p5d_dsm.tif')
    jpg_to_write = os.path.join(out_dir, 'aggregate_2p5d_dsm.jpg')
    ul_e, ul_n, e_size, n_size, e_resolution, n_resolution = produce_dsm_from_height(
        work_dir, all_dsm_mean_no_outliers, tif_to_write, jpg_to_write)

    void_ratio = np.sum(np.isnan(all_dsm_mean_no_outliers)) / all_dsm_mean_no_outliers.size
    logging.info('\n After aggregation, empty ratio: {} '.format(void_ratio))

    # create a colored point cloud (:91-113): one vertex per non-empty cell at its upper-left corner
    xx = ul_n - np.arange(n_size) * n_resolution
    yy = ul_e + np.arange(e_size) * e_resolution
    xx, yy = np.meshgrid(xx, yy, indexing='ij')
    zz = all_dsm_mean_no_outliers.reshape(-1)
    valid_mask = np.logical_not(np.isnan(zz))
    # vertex colours: the reference reads its preview jpg back (:100); the same height -> colour mapping (same table, same
    # alt_min / alt_max clip as produce_dsm_from_height's preview, same [1, 99] percentile range) applied directly, so
    # the colours neither depend on the previews being written nor carry JPEG noise
    from .visualization.plot_height_map import height_to_rgb
    with open(os.path.join(work_dir, 'aoi.json')) as fp:
        aoi_for_color = json.load(fp)
    rgb, _, _ = height_to_rgb(np.clip(all_dsm_mean_no_outliers, aoi_for_color['alt_min'], aoi_for_color['alt_max']))
    color = rgb.reshape((-1, 3))
    utm_points = np.stack((yy.reshape(-1)[valid_mask], xx.reshape(-1)[valid_mask], zz[valid_mask]), axis=1)
    with open(os.path.join(work_dir, 'aoi.json')) as fp:
        aoi_dict = json.load(fp)
    comments = ['projection: UTM {}{}'.format(aoi_dict['zone_number'], aoi_dict['hemisphere'])]
    np2ply(utm_points, os.path.join(out_dir, 'aggregate_2p5d.ply'), color=color[valid_mask], comments=comments,
           use_double=True)


def fuse_dsm_tifs(tif_files, engine=None):
    """Stage C on existing per-view GeoTIFFs (the reference's :57-81 in isolation): read, NaN-ify nodata, fuse."""
    from . import engine as _engine
    imgs = [read_dsm_tif(f)[0] for f in tif_files]
    if engine is None:
        ctx_engine = None
        stack = torch.from_numpy(np.stack(imgs)).cuda()
        # an AOI-less context is enough for stages C + blur
        ctx, dev = _engine.default_context()
        import ctypes as C
        from ._native import lib, check
        V, H, W = stack.shape
        mean = torch.empty((H, W), dtype=torch.float32, device=dev)
        out = torch.empty_like(mean)
        st = _engine._stream(dev)
        check(lib.vs_fuse_views(ctx.handle, _engine._ptr(stack), stack.stride(0), V, H, W, _engine._ptr(mean), st))
        check(lib.vs_median3x3(ctx.handle, _engine._ptr(mean), 0, H, W, 0, H, _engine._ptr(out),
                               _engine.DEFAULT_SIMD_LANES, C.c_void_p(0), st))
        del ctx_engine
        return out.cpu().numpy()
    return engine.fuse_and_blur(torch.from_numpy(np.stack(imgs)).to(engine.device)).cpu().numpy()


if __name__ == '__main__':
    pass
