"""The row-band form of the oracle (used for parity at BASELINE.json's full sizes) equals the full-grid oracle bit
for bit, including bands at the grid's top and bottom edges, 1-row bands, and views that miss a band."""
import numpy as np
import pytest

from oracle import geodesy, pipeline as op
from vissatsatellitestereo_b200 import synthetic as S


@pytest.mark.parametrize('name,depth,grid', [('C2', 384, 384), ('C3', 256, 640), ('C5', 320, 320)])
def test_banded_oracle_equals_full_oracle(name, depth, grid):
    cfg = S.scaled(S.CONFIGS[name], views=4, depth=depth, grid=grid, name='bands')
    scene = S.make_scene(cfg, geodesy, device='cpu')
    n = cfg.n_size
    bands = [(0, 9), (n // 3, n // 3 + 1), (n // 2 - 7, n // 2 + 12), (n - 5, n)]
    full_views = []
    for v in range(cfg.n_views):
        d, M = scene.depths[v].numpy(), scene.mats[v]
        full, _ = op.convert_depth_map(d, M, scene.aoi, cfg.res, cfg.res, fast=(v != 0))    # view 0: the literal loop
        full_views.append(full)
        outs, info = op.convert_depth_map_rows(d, M, scene.aoi, cfg.res, cfg.res, bands)
        assert info['n_valid'] == int(np.sum(d > 0))
        for (rb, re), o in zip(bands, outs):
            assert o.dtype == np.float32 and o.shape == (re - rb, cfg.e_size)
            assert np.array_equal(o, full[rb:re], equal_nan=True), (v, rb, re)
        lo, hi = op.view_row_range(d, M, scene.aoi, cfg.res)
        raw = op._scatter_nanmax(op.enu_points_to_utm(op.unproject_depth(d, M)[1][:, :3], scene.aoi),
                                 scene.aoi['ul_easting'], scene.aoi['ul_northing'], cfg.res, cfg.res, cfg.e_size, cfg.n_size)
        rows_hit = np.nonzero((~np.isnan(raw)).any(axis=1))[0]
        assert lo <= rows_hit.min() and rows_hit.max() <= hi
    fused = op.fuse_dsms(full_views)
    for rb, re in bands:
        h0, h1 = max(rb - 1, 0), min(re + 1, n)
        got = op.fuse_rows([f[h0:h1] for f in full_views], rb, re, n)
        assert np.array_equal(got, fused[rb:re], equal_nan=True)
