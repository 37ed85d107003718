"""Sparse coverage (large AOIs, BASELINE.json configs[2]): occupancy bitmap written by stage B, fusion that reads only
the marked (tile, view) pairs.  The sparse result must be bit-identical to the dense fusion of the same stack with the
unmarked pairs set to NaN -- and the memory of unmarked pairs must never be read (it is poisoned here)."""
import numpy as np
import pytest
import torch

from oracle import geodesy, pipeline as op

pytestmark = pytest.mark.gpu

TW, TH = 64, 32


@pytest.fixture(scope='module')
def eng():
    from vissatsatellitestereo_b200 import engine, synthetic as S
    engine.require_cuda()
    cfg = S.scaled(S.CONFIGS['C1'], views=1, depth=64, grid=32)
    return engine.DsmEngine(S.make_aoi(cfg, geodesy), cfg.res, cfg.res)


def _same(a, b):
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=-1e9), torch.nan_to_num(b, nan=-1e9))


def _random_sparse_stack(V, H, W, seed, p_profile):
    """Dense cube (V, H, W) whose (tile, view) occupancy follows p_profile(ty, tx) + the matching bitmap."""
    rng = np.random.default_rng(seed)
    Ty, Tx, words = -(-H // TH), -(-W // TW), -(-V // 32)
    cube = (30 + 5 * rng.normal(size=(V, H, W))).astype(np.float32)
    cube[rng.random(cube.shape) < 0.25] = np.nan
    cube[:, 1::7, :] = np.round(cube[:, 1::7, :])                 # ties
    occ = np.zeros((Ty, Tx, words), dtype=np.uint32)
    for ty in range(Ty):
        for tx in range(Tx):
            present = rng.random(V) < p_profile(ty, tx)
            sl = (slice(None), slice(ty * TH, (ty + 1) * TH), slice(tx * TW, (tx + 1) * TW))
            blk = cube[sl]
            blk[~present] = np.nan
            cube[sl] = blk
            for g in np.nonzero(present)[0]:
                occ[ty, tx, g >> 5] |= np.uint32(1) << np.uint32(g & 31)
    return cube, occ.view(np.int32)


@pytest.mark.parametrize('V,H,W', [(5, 70, 130), (50, 97, 200), (130, 64, 128), (200, 96, 192), (300, 40, 100), (600, 33, 70)])
def test_sparse_fusion_bit_identical_to_dense(eng, V, H, W):
    # occupancy from 0 to 100 % across the tiles: every bin up to V is hit, including empty tiles and full ones
    Ty, Tx = -(-H // TH), -(-W // TW)
    prof = lambda ty, tx: ((ty * Tx + tx) / max(Ty * Tx - 1, 1)) ** 0.7                          # noqa: E731
    cube, occ = _random_sparse_stack(V, H, W, V, prof)
    dense = torch.from_numpy(cube).cuda()
    want = eng.fuse(dense)                                                   # dense kernels (bit-exact vs numpy elsewhere)
    occ_t = torch.from_numpy(occ).cuda()
    # poison everything the bitmap does not mark: the sparse path must not read it
    poisoned = torch.where(torch.isnan(eng.densify(torch.zeros_like(dense), occ_t)), torch.full_like(dense, 1e30), dense)
    got = eng.fuse(poisoned, occ=occ_t)
    assert _same(got, want)
    # numpy on a few rows (the sparse path keeps numpy's summation order over the ORIGINAL view axis)
    rows = slice(0, min(H, 20))
    assert np.array_equal(got[rows].cpu().numpy(), op.fuse_dsms([cube[v, rows] for v in range(V)], blur=False), equal_nan=True)
    # row-band form: planes hold grid rows [row0, row0 + rows) with row0 inside a tile
    for row0, rows_n in ((5, H - 9), (TH, H - TH), (H - 3, 3)):
        if rows_n <= 0:
            continue
        band = poisoned[:, row0:row0 + rows_n].contiguous()
        got_b = eng.fuse(band, occ=occ_t, row0=row0)
        assert _same(got_b, want[row0:row0 + rows_n]), (row0, rows_n)


def test_stage_b_occupancy_and_sparse_end_to_end(eng):
    """A C3-like scene (views at random offsets over a larger grid): the bitmap stage B writes marks every tile whose
    plane is not all-NaN (plus, conservatively, tiles with data within the key box a CTA loads: 2 rows / 4 columns
    around the tile), and the sparse fusion of the stack equals the dense one."""
    from vissatsatellitestereo_b200 import engine as E, synthetic as S
    cfg = S.scaled(S.CONFIGS['C3'], views=40, depth=160, grid=448, name='sparse_e2e')
    cfg.n_size = 416
    scene = S.make_scene(cfg, geodesy, device='cuda')
    e = E.DsmEngine(scene.aoi, cfg.res, cfg.res)
    V = cfg.n_views
    stack = torch.full((V, e.n_size, e.e_size), 123.0, dtype=torch.float32, device='cuda')
    occ = e.alloc_occupancy(V)
    assert tuple(occ.shape) == (13, 7, 2)
    e.set_occupancy(occ, stack, 0)
    e.views_to_dsm(scene.depths, scene.mats, stack)
    e.set_occupancy(None)
    dense_ref = torch.empty_like(stack)
    e.views_to_dsm(scene.depths, scene.mats, dense_ref)                      # same kernels without the marking
    assert _same(stack, dense_ref)
    assert _same(e.densify(stack, occ), stack), 'a tile with data is not marked'
    # marked tiles: those with data, or with data within a few cells of them (the per-view DSM has data wherever the
    # key grid has, plus the 1-cell hole fill; the key box of a tile reaches 2 rows / 4 columns beyond it)
    has = (~torch.isnan(stack)).float()
    near = torch.nn.functional.max_pool2d(has[None], (7, 11), 1, (3, 5))[0]
    Ty, Tx = occ.shape[:2]
    pad = torch.zeros((V, Ty * TH, Tx * TW), device='cuda')
    pad[:, :e.n_size, :e.e_size] = near
    tile_has = pad.view(V, Ty, TH, Tx, TW).amax(dim=(2, 4)) > 0                 # (V, Ty, Tx)
    g = torch.arange(V, device='cuda')
    bits = ((occ[:, :, g // 32] >> (g % 32).to(torch.int32)) & 1).permute(2, 0, 1).bool()
    assert not (bits & ~tile_has).any(), 'a tile without data (incl. its 2-cell halo) is marked'
    frac = bits.float().mean().item()
    assert 0.05 < frac < 0.9, frac
    want = e.fuse_and_blur(stack)
    got = e.median3x3(e.fuse(stack, occ=occ), count_nan=True)
    assert _same(got, want)
    # the captured step with occupancy replays identically
    gr = e.capture_step(scene.depths, scene.mats, stack, fuse=True, occ=occ)
    for _ in range(2):
        stack.fill_(5.0)
        assert _same(gr.replay(), want)
    e.close()
