#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, device-resident inputs): python tools/microbench.py [what ...]
what: fuse k1 k2 k4 (default: all)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from vissatsatellitestereo_b200 import engine as E, synthetic as S  # noqa: E402
from vissatsatellitestereo_b200.lib import latlon_utm_converter as geo  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    what = sys.argv[1:] or ['fuse', 'k1', 'k2', 'k4']
    cfg = S.SynthConfig(**S.CONFIGS['C2'].__dict__)
    NB = int(os.environ.get('VISSAT_MB_BASE', '4'))      # distinct synthetic views behind the fusion stacks
    cfg.n_views = NB
    aoi = S.make_aoi(cfg, geo)
    eng = E.DsmEngine(aoi, cfg.res, cfg.res)
    eng.collect_stats = False
    dev = eng.device
    terrain = S.Terrain(cfg, device=dev)
    mats = [S.make_camera(cfg, v, aoi['alt_min'])[0] for v in range(cfg.n_views)]
    depths = [S.make_depth_map(cfg, v, mats[v], terrain, device=dev) for v in range(cfg.n_views)]
    G = eng.n_size * eng.e_size
    P = cfg.height * cfg.width
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if 'k1' in what:
        i = [0]

        def k1():
            eng.clear_keygrid()
            eng.rasterize(depths[i[0] % NB], mats[i[0] % NB], clear=False)
            i[0] += 1
        t = timeit(k1, 20)
        print('k1 (clear + unproject_scatter) {:.1f} us/view  {:.1f} Gpix/s  {:.0f} GB/s'.format(t * 1e3, P / t / 1e6, 4 * P / t / 1e6))
        for cname in ('C4', 'C5'):
            c2 = S.SynthConfig(**S.CONFIGS[cname].__dict__)
            a2 = S.make_aoi(c2, geo)
            e2 = E.DsmEngine(a2, c2.res, c2.res)
            e2.collect_stats = False
            t2 = S.Terrain(c2, device=dev)
            M2 = S.make_camera(c2, 0, a2['alt_min'])[0]
            d2 = S.make_depth_map(c2, 0, M2, t2, device=dev)
            del t2
            t = timeit(lambda: e2.rasterize(d2, M2), 10)
            print('k1 {} ({}x{} -> {}^2 @ {}) {:.1f} us/view  {:.1f} Gpix/s'.format(cname, c2.height, c2.width, c2.e_size, c2.res, t * 1e3, c2.height * c2.width / t / 1e6))
            out = torch.empty((e2.n_size, e2.e_size), dtype=torch.float32, device=dev)
            t = timeit(lambda: e2.finalize(out=out), 10)
            print('k2 {} {:.1f} us/view {:.0f} GB/s'.format(cname, t * 1e3, 8 * e2.n_size * e2.e_size / t / 1e6))
            del e2, d2, out
    if 'k2' in what:
        eng.rasterize(depths[0], mats[0])
        out = torch.empty((eng.n_size, eng.e_size), dtype=torch.float32, device=dev)
        t = timeit(lambda: eng.finalize(out=out), 20)
        print('k2 grid_finalize {:.1f} us/view  {:.0f} GB/s (8 B/cell)'.format(t * 1e3, 8 * G / t / 1e6))
    if 'k4' in what:
        img = eng.view_dsm(depths[0], mats[0])
        out = torch.empty_like(img)
        t = timeit(lambda: eng.median3x3(img, out=out), 20)
        print('k4 median3x3 {:.1f} us  {:.0f} GB/s (8 B/cell)'.format(t * 1e3, 8 * G / t / 1e6))
    if 'fuse' in what:
        base = torch.stack([eng.view_dsm(depths[v], mats[v]).clone() for v in range(NB)])
        for V, rows in ((8, 2048), (16, 2048), (24, 2048), (32, 2048), (40, 2048), (50, 2048), (64, 2048), (100, 1024),
                        (200, 512), (400, 256)):
            # VISSAT_MB_SHUFFLE=1: base view drawn at random per view (lanes of the multi-lane kernels see alike runs);
            # default: view v uses base view v % 4 (lane = view mod 4 / 8: every lane sees ONE base view)
            idx = torch.randint(0, NB, (V,), device=dev) if os.environ.get('VISSAT_MB_SHUFFLE') == '1' else torch.arange(V, device=dev) % NB
            stack = (base[idx, :rows] + torch.randn((V, 1, 1), device=dev) * 0.5).contiguous()
            stack[torch.rand(stack.shape, device=dev) < 0.1] = float('nan')
            out = torch.empty((rows, eng.e_size), dtype=torch.float32, device=dev)

            def f():
                flush.zero_() if False else None
                eng.fuse(stack, out=out)
            t = timeit(f, 5, 2)
            cells = rows * eng.e_size
            print('fuse V={:4d} rows={:5d}: {:8.1f} us  {:7.1f} Mcell/s  {:6.0f} GB/s read'.format(
                V, rows, t * 1e3, cells / t / 1e3, 4 * V * cells / t / 1e6))
            del stack, out


if __name__ == '__main__':
    main()
