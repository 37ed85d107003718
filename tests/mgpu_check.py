"""torchrun helper: N-rank fused DSM == single-rank fused DSM, bit for bit (run by test_gpu_run_fuse.py and by
hand: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_check.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    from vissatsatellitestereo_b200 import distributed as D, engine as E, synthetic as S
    from vissatsatellitestereo_b200.lib import latlon_utm_converter as geo
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    cfg = S.scaled(S.CONFIGS['C3'], views=13, depth=192, grid=301, name='mgpu')
    cfg.n_size = 257
    aoi = S.make_aoi(cfg, geo)
    terrain = S.Terrain(cfg, device=dev)
    eng = E.DsmEngine(aoi, cfg.res, cfg.res, device=dev)

    def view(v):
        M, _ = S.make_camera(cfg, v, aoi['alt_min'])
        return eng.view_dsm(S.make_depth_map(cfg, v, M, terrain, device=dev), M).clone()

    a, b = D.split_views(cfg.n_views, world)[rank]
    local = torch.stack([view(v) for v in range(a, b)]) if b > a else \
        torch.empty((0, eng.n_size, eng.e_size), dtype=torch.float32, device=dev)
    counts = [hi - lo for lo, hi in D.split_views(cfg.n_views, world)]
    band, (r0, r1) = D.fuse_distributed(eng, local, counts)
    full = D.gather_bands(band, eng.n_size, eng.e_size)
    ok = True
    if rank == 0:
        single = eng.fuse_and_blur(torch.stack([view(v) for v in range(cfg.n_views)]))
        ok = torch.equal(torch.nan_to_num(full, nan=-1e9), torch.nan_to_num(single, nan=-1e9))
        print('MGPU_OK' if ok else 'MGPU_MISMATCH', 'world', world, 'nan frac', float(torch.isnan(single).float().mean()))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
