"""N-rank result == single-rank result, bit for bit, on small grids through both exchange transports (NCCL row-band
all-to-all; stage-B peer stores with the sparse occupancy exchange).  SURVEY.md §8(e): the fused DSM must not depend
on the number of ranks.  Used by `bench.py --gpus N` (N > 1; the outcome goes into its JSON line as
`mgpu_bit_identical`) and by tests/mgpu_check.py under torchrun.  GPU-vs-GPU comparison: no oracle involved.

Requires an initialised NCCL process group and the current CUDA device set to this rank's GPU.
"""
import torch
import torch.distributed as dist


def _same(a, b):
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=-1e9), torch.nan_to_num(b, nan=-1e9))


# (name, grid width, grid height (None: world + 1 rows -> 1-row bands), views, base config):
#   odd row pitch (scalar stores), bands cutting through 32-row tiles; 16-byte-aligned rows; fewer rows than
#   2 x ranks; a sparse-coverage case (views at random offsets over a grid of several tile rows)
CASES = (('odd', 301, 257, 13, 'C3'), ('vec', 320, 130, 9, 'C3'), ('thin', 64, None, 5, 'C3'),
         ('sparse', 448, 416, 24, 'C3'))


def check_case(name, grid_w, grid_h, n_views, base, dev, rank, world, log=None):
    from . import distributed as D, engine as E, synthetic as S
    from .lib import latlon_utm_converter as geo
    if grid_h is None:
        grid_h = max(world + 1, 3)
    cfg = S.scaled(S.CONFIGS[base], views=n_views, depth=192 if name != 'sparse' else 160, grid=grid_w, name='mgpu_' + name)
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

    # ---- transport 1: NCCL all-to-all of row bands (dense)
    band, (r0, r1) = D.fuse_distributed(eng, local, counts)
    full = D.gather_bands(band, eng.n_size, eng.e_size)
    want_stack, _, _ = D.exchange_rowbands(local, counts, eng.n_size)

    # ---- transport 2: stage B stores the row bands into the peers' stacks (three steps: both buffers get reused)
    ok_peer, peer_state = True, 'OK'
    try:
        px = D.PeerExchange(eng, torch.empty_like(local), counts)      # fails on every rank or on none
        px.keep_last_occ = True
    except Exception as e:
        px, peer_state = None, 'UNAVAILABLE ({})'.format(e)
    for step in range(3 if px is not None else 0):
        px.local.fill_(7.0)
        px.begin_step()
        eng.views_to_dsm([depth(v) for v in range(a, b)], mats[a:b], px.local)
        band2, (q0, q1) = px.fuse_band()
        checks = {'band_rows': (q0, q1) == (r0, r1), 'local_planes': _same(px.local, local),
                  'fused_band': _same(band2, band), 'band_stack': px.check_band_stack(want_stack)}
        if not all(checks.values()) and log:
            log('MGPU_PEER_DETAIL', name, 'rank', rank, 'step', step, {k: bool(v) for k, v in checks.items()})
        ok_peer &= all(checks.values())
        torch.cuda.synchronize()
    if px is not None:
        px.close()
    t = torch.tensor([1 if ok_peer else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_peer = int(t.item()) == 1

    ok = True
    if rank == 0:
        single = eng.fuse_and_blur(torch.stack([view(v) for v in range(cfg.n_views)]))
        ok = _same(full, single)
        if log:
            log('MGPU_OK' if ok else 'MGPU_MISMATCH', name, 'world', world, 'grid', (grid_h, grid_w), 'nan frac',
                float(torch.isnan(single).float().mean()))
            log('MGPU_PEER_' + peer_state if ok_peer else 'MGPU_PEER_MISMATCH', name, 'world', world)
    eng.close()
    return ok, ok_peer, peer_state


def run(dev, rank, world, log=None):
    """All cases.  Returns {'bit_identical': bool, 'nccl': bool, 'peer_store': bool | None (unavailable), 'cases': [...]}
    identical on every rank."""
    ok_n = ok_p = True
    peer_avail = True
    for name, w, h, nv, base in CASES:
        a, b, state = check_case(name, w, h, nv, base, dev, rank, world, log)
        ok_n &= a
        ok_p &= b
        peer_avail &= state == 'OK'
    t = torch.tensor([1 if ok_n else 0, 1 if ok_p else 0, 1 if peer_avail else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_n, ok_p, peer_avail = (int(x) == 1 for x in t.tolist())
    return {'bit_identical': ok_n and ok_p, 'nccl': ok_n, 'peer_store': ok_p if peer_avail else None,
            'cases': [c[0] for c in CASES], 'world': world}
