"""torchrun helper: N-rank fused DSM == single-rank fused DSM, bit for bit, through both exchange transports (NCCL
row-band all-to-all; stage-B peer stores).  Run by test_gpu_run_fuse.py and by hand:
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_check.py"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def same(a, b):
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=-1e9), torch.nan_to_num(b, nan=-1e9))


def check_config(name, grid_w, grid_h, n_views, dev, rank, world):
    from vissatsatellitestereo_b200 import distributed as D, engine as E, synthetic as S
    from vissatsatellitestereo_b200.lib import latlon_utm_converter as geo
    cfg = S.scaled(S.CONFIGS['C3'], views=n_views, depth=192, grid=grid_w, name='mgpu_' + name)
    cfg.n_size = grid_h
    aoi = S.make_aoi(cfg, geo)
    terrain = S.Terrain(cfg, device=dev)
    eng = E.DsmEngine(aoi, cfg.res, cfg.res, device=dev)
    assert (eng.e_size, eng.n_size) == (grid_w, grid_h)
    mats = [S.make_camera(cfg, v, aoi['alt_min'])[0] for v in range(cfg.n_views)]
    depths = {}

    def depth(v):
        if v not in depths:
            depths[v] = S.make_depth_map(cfg, v, mats[v], terrain, device=dev)
        return depths[v]

    def view(v):
        return eng.view_dsm(depth(v), mats[v]).clone()

    a, b = D.split_views(cfg.n_views, world)[rank]
    local = torch.stack([view(v) for v in range(a, b)]) if b > a else \
        torch.empty((0, eng.n_size, eng.e_size), dtype=torch.float32, device=dev)
    counts = [hi - lo for lo, hi in D.split_views(cfg.n_views, world)]

    # ---- transport 1: NCCL all-to-all of row bands
    band, (r0, r1) = D.fuse_distributed(eng, local, counts)
    full = D.gather_bands(band, eng.n_size, eng.e_size)
    want_stack, _, _ = D.exchange_rowbands(local, counts, eng.n_size)

    # ---- transport 2: stage B stores the row bands into the peers' stacks (three steps: both buffers get reused)
    ok_peer, peer_state = True, 'OK'
    try:
        px = D.PeerExchange(eng, torch.empty_like(local), counts)      # fails on every rank or on none
    except Exception as e:
        px, peer_state = None, 'UNAVAILABLE ({})'.format(e)
    for step in range(3 if px is not None else 0):
        px.local.fill_(7.0)
        px.begin_step()
        eng.views_to_dsm([depth(v) for v in range(a, b)], mats[a:b], px.local)
        band_stack, (q0, q1), (h0, h1) = px.finish()
        ok_peer &= (q0, q1) == (r0, r1) and same(px.local, local) and same(band_stack, want_stack)
        if q1 > q0:
            mean = eng.fuse(band_stack)
            band2 = eng.median3x3(mean, row_begin=q0, row_end=q1, in_row0=h0, h_total=eng.n_size)
            ok_peer &= same(band2, band)
        torch.cuda.synchronize()
    if px is not None:
        px.close()
    t = torch.tensor([1 if ok_peer else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_peer = int(t.item()) == 1

    ok = True
    if rank == 0:
        single = eng.fuse_and_blur(torch.stack([view(v) for v in range(cfg.n_views)]))
        ok = same(full, single)
        print('MGPU_OK' if ok else 'MGPU_MISMATCH', name, 'world', world, 'grid', (grid_h, grid_w), 'nan frac',
              float(torch.isnan(single).float().mean()))
        print('MGPU_PEER_' + peer_state if ok_peer else 'MGPU_PEER_MISMATCH', name, 'world', world, flush=True)
    eng.close()
    return ok and ok_peer


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    ok = True
    # odd row pitch (scalar stores), bands cutting through 32-row tiles; then 16-byte-aligned rows; then fewer rows
    # than ranks x 2 (1-row bands: a row is band, upper halo and lower halo at once)
    for name, w, h, nv in (('odd', 301, 257, 13), ('vec', 320, 130, 9), ('thin', 64, max(world + 1, 3), 5)):
        ok &= check_config(name, w, h, nv, dev, rank, world)
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == '__main__':
    main()
