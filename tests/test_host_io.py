"""Host-side pieces of the drop-in that need no GPU: GeoTIFF and PLY I/O, view/row-band partitioning, and the
world_size-2 row-band exchange on the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch


def test_geotiff_round_trip(tmp_path):
    from vissatsatellitestereo_b200.lib import dsm_util
    rng = np.random.default_rng(0)
    for shape in [(1, 1), (37, 53), (600, 700)]:
        img = rng.normal(30, 10, size=shape).astype(np.float32)
        img[rng.random(shape) < 0.2] = np.nan
        f = str(tmp_path / 'a_{}_{}.tif'.format(*shape))
        dsm_util.write_dsm_tif(img, f, (354052.3651180889, 6182702.10540914, 0.3, 0.3), (21, 'S'), nodata_val=-10000)
        got, meta = dsm_util.read_dsm_tif(f)
        assert got.dtype == np.float32 and np.array_equal(got, img, equal_nan=True)
        assert meta['geo'] == (354052.3651180889, 0.3, 0.0, 6182702.10540914, 0.0, -0.3)       # lib/dsm_util.py:150
        assert meta['zone_number'] == 21 and meta['hemisphere'] == 'S' and meta['nodata'] == -10000.0
        assert meta['img_width'] == shape[1] and meta['img_height'] == shape[0]
        assert meta['meta'] == {'AREA_OR_POINT': 'Area'}                                      # :157
        assert meta['lr_easting'] == meta['ul_easting'] + (shape[1] - 1) * 0.3
        assert dsm_util.parse_proj_str(meta['proj']) == (21, 'S')
        if not np.isnan(img).all():
            assert meta['alt_min'] == float(np.nanmin(img))
    # values within np.isclose of the nodata value come back as NaN, like the reference (:69-72)
    img = np.array([[1.0, -10000.05, -9999.0]], dtype=np.float32)
    f = str(tmp_path / 'b.tif')
    dsm_util.write_dsm_tif(img, f, (0.0, 0.0, 0.5, 0.5), (32, 'N'), nodata_val=-10000)
    got, meta = dsm_util.read_dsm_tif(f)
    assert np.isnan(got[0, 1]) and got[0, 2] == -9999.0 and meta['hemisphere'] == 'N' and meta['zone_number'] == 32
    # libtiff (through OpenCV) reads the raster too
    import cv2
    assert np.array_equal(cv2.imread(f, cv2.IMREAD_UNCHANGED), img)
    assert dsm_util.get_driver('x.tif') is not None and dsm_util.get_driver('x.xyz') is None


def test_ply_round_trip(tmp_path):
    from vissatsatellitestereo_b200.lib.ply_np_converter import np2ply, ply2np
    rng = np.random.default_rng(1)
    v = rng.normal(size=(100, 3)) * 1e5
    c = (rng.random((100, 3)) * 255).astype(np.uint8)
    f = str(tmp_path / 'p.ply')
    np2ply(v, f, color=c, comments=['projection: UTM 21S'], use_double=True)     # aggregate_2p5d.py:110-113
    d, col, com = ply2np(f)
    assert np.array_equal(d, v) and np.array_equal(col, c) and com == ['projection: UTM 21S']
    np2ply(v, f, use_double=False)
    d, col, com = ply2np(f)
    assert np.array_equal(d, v.astype(np.float32)) and col is None and com is None
    np2ply(v[:5], f, color=c[:5], text=True)
    d, col, _ = ply2np(f)
    assert np.allclose(d, v[:5], atol=1e-4) and np.array_equal(col, c[:5])


def test_partitions():
    from vissatsatellitestereo_b200 import distributed as D
    from vissatsatellitestereo_b200.aggregate_2p5d_util import split_big_list
    assert D.split_views(50, 8) == [(0, 7), (7, 14), (14, 20), (20, 26), (26, 32), (32, 38), (38, 44), (44, 50)]
    assert D.split_views(3, 4) == [(0, 1), (1, 2), (2, 3), (0, 0)]
    bands = D.row_bands(2048, 8)
    assert bands[0] == (0, 256) and bands[-1] == (1792, 2048)
    assert D.band_with_halo((0, 256), 2048) == (0, 257) and D.band_with_halo((256, 512), 2048) == (255, 513)
    assert D.band_with_halo((1792, 2048), 2048) == (1791, 2048)
    # aggregate_2p5d_util.py:109-122
    assert split_big_list(list(range(10)), 3) == [[0, 1, 2, 3], [4, 5, 6], [7, 8, 9]]
    assert split_big_list([1, 2], 4) == [[1], [2]]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _exchange_worker(rank, world, port, n_rows, W, counts, out):
    import torch.distributed as dist
    from vissatsatellitestereo_b200 import distributed as D
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{}'.format(port), rank=rank, world_size=world)
    try:
        first = sum(counts[:rank])
        # plane v of the global stack holds v*1e6 + row*1e3 + col
        rows = torch.arange(n_rows, dtype=torch.float32)[:, None] * 1e3 + torch.arange(W, dtype=torch.float32)[None, :]
        local = torch.stack([rows + (first + i) * 1e6 for i in range(counts[rank])]) if counts[rank] else \
            torch.empty((0, n_rows, W))
        band_stack, (r0, r1), (h0, h1) = D.exchange_rowbands(local, counts, n_rows)
        want = torch.stack([rows[h0:h1] + v * 1e6 for v in range(sum(counts))])
        ok = torch.equal(band_stack, want)
        # the wave form (what the benchmark overlaps with compute) must deliver the same stack
        xch = D.WaveExchanger(local, counts)
        vmax = max(counts)
        for a in range(0, vmax, 2):
            xch.send_wave(a, min(a + 2, vmax))
        bs2, band2, halo2 = xch.finish()
        ok = ok and torch.equal(bs2, want) and band2 == (r0, r1) and halo2 == (h0, h1)
        # a fake "fusion" (mean over views of the band) and the gather of the bands on rank 0
        band = band_stack.mean(dim=0)[r0 - h0: r1 - h0].contiguous()
        full = D.gather_bands(band, n_rows, W)
        if rank == 0:
            ok = ok and torch.allclose(full, rows + (sum(counts) - 1) / 2 * 1e6)
        out[rank] = bool(ok) and (r0, r1) == D.row_bands(n_rows, world)[rank]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_rows,counts', [(37, [3, 2]), (8, [1, 4]), (5, [2, 0])])
def test_rowband_exchange_world2_gloo(n_rows, counts):
    """the all-to-all that turns view-sharded per-view DSMs into row-band-sharded stacks."""
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, n_rows, 11, counts, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


@pytest.mark.parametrize('world,n_rows,counts', [(3, 10, [3, 2, 2]), (4, 3, [1, 2, 0, 1])])
def test_rowband_exchange_more_ranks_gloo(world, n_rows, counts):
    """Uneven view blocks, row counts that do not divide by the world size, middle ranks with two halo neighbours, and
    fewer rows than ranks (an empty band)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, n_rows, 7, counts, out), nprocs=world, join=True)
    assert dict(out) == {r: True for r in range(world)}


def test_bind_to_gpu_numa_node_with_fake_sysfs(tmp_path, monkeypatch):
    """hostbind: affinity follows <sysfs>/bus/pci/devices/<bus>/numa_node -> node<N>/cpulist; unknown -> unchanged."""
    import os
    from vissatsatellitestereo_b200 import hostbind as HB
    assert HB._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    if not hasattr(os, 'sched_setaffinity'):
        return
    allowed = sorted(os.sched_getaffinity(0))
    monkeypatch.setattr(HB, 'gpu_pci_bus_id', lambda i: '0000:1b:00.0')
    sysfs = tmp_path / 'sys'
    (sysfs / 'bus/pci/devices/0000:1b:00.0').mkdir(parents=True)
    (sysfs / 'devices/system/node/node1').mkdir(parents=True)
    assert HB.bind_to_gpu_numa_node(0, str(sysfs)) == -1                 # no numa_node file
    (sysfs / 'bus/pci/devices/0000:1b:00.0/numa_node').write_text('-1\n')
    assert HB.bind_to_gpu_numa_node(0, str(sysfs)) == -1                 # kernel does not know
    (sysfs / 'bus/pci/devices/0000:1b:00.0/numa_node').write_text('1\n')
    (sysfs / 'devices/system/node/node1/cpulist').write_text('{}\n'.format(allowed[0]))
    try:
        assert HB.gpu_numa_node(0, str(sysfs)) == 1
        assert HB.bind_to_gpu_numa_node(0, str(sysfs)) == 1
        assert os.sched_getaffinity(0) == {allowed[0]}
        (sysfs / 'devices/system/node/node1/cpulist').write_text('100000\n')   # none of our CPUs: leave as is
        assert HB.bind_to_gpu_numa_node(0, str(sysfs)) == -1
        assert os.sched_getaffinity(0) == {allowed[0]}
    finally:
        os.sched_setaffinity(0, set(allowed))


def test_convert_depth_maps_read_ahead_is_bounded_and_ordered(tmp_path, monkeypatch):
    """Host logic of aggregate_2p5d_util.convert_depth_maps (:131-146) with the GPU parts stubbed: every matching item
    is processed once, in sorted order; other files are passed through to the worker (which logs and skips them); at
    most 2 * n_io depth maps are loaded ahead of the consumer; stale output is wiped (:127-128)."""
    import threading
    from vissatsatellitestereo_b200 import aggregate_2p5d_util as U
    work = tmp_path / 'work'
    depth_dir = work / 'colmap/mvs/stereo/depth_maps'
    depth_dir.mkdir(parents=True)
    names = ['{:04d}.png.geometric.bin'.format(i) for i in range(23)] + ['0003.png.photometric.bin', 'notes.geometric.txt']
    for n in names:
        (depth_dir / n).write_bytes(b'x')
    out_dir = work / 'colmap/mvs/dsm'
    (out_dir / 'dsm_tif').mkdir(parents=True)
    (out_dir / 'dsm_tif/stale.tif').write_bytes(b'old')
    lock = threading.Lock()
    live = {'loaded': 0, 'max': 0}
    seen = []

    def fake_read(path):
        with lock:
            live['loaded'] += 1
            live['max'] = max(live['max'], live['loaded'])
        return os.path.basename(path)

    class FakePipeline:
        """stands in for the GPU side (_ViewPipeline): records the order, releases the prefetched array"""

        def __init__(self, eng, aoi, mats, out, depth_type, n_views, io_pool):
            assert n_views == 24 and depth_type == 'geometric'   # 23 depth maps + notes.geometric.txt
            self.views = []

        def submit(self, item, depth_host):
            if depth_host is not None:
                assert depth_host == item                 # the prefetched array belongs to this item
                with lock:
                    live['loaded'] -= 1
            seen.append(item)
            if item.endswith('.bin'):
                self.views.append((len(self.views), 0, item))
                return len(self.views) - 1, item
            return None

        def finish(self):
            return 'stack', self.views

    monkeypatch.setattr(U, 'read_array', fake_read)
    monkeypatch.setattr(U, '_ViewPipeline', FakePipeline)
    monkeypatch.setattr(U, '_make_engine', lambda wd: ('engine', {'aoi': 1}))
    monkeypatch.setattr(U, 'load_inv_proj_mats', lambda d: {})
    U.convert_depth_maps(str(work), str(out_dir), 'geometric', max_processes=3)
    want = sorted(n for n in names if 'geometric' in n)
    assert seen == want
    assert live['max'] <= 2 * 3 + 1 and live['loaded'] == 0
    assert not (out_dir / 'dsm_tif/stale.tif').exists()
    res = U._RESULTS.pop(os.path.abspath(str(out_dir)))
    assert [v[2] for v in res['views']] == [n for n in want if n.endswith('.bin')] and res['world'] == 1
    assert res['stack'] == 'stack'


def test_preview_colour_table_and_psnr(tmp_path):
    """N3: previews use the reference's colour table with matplotlib's ListedColormap rule (plot_height_map.py:39-57,
    save_image_only.py:62-101).  The written JPEG is compared (PSNR) with an independent rendering of that rule."""
    import cv2
    from vissatsatellitestereo_b200.visualization import plot_height_map as PH
    from vissatsatellitestereo_b200.visualization._colormap_height import COLORMAP_HEIGHT as LUT
    assert LUT.shape == (197, 3) and LUT.dtype == np.uint8
    assert tuple(LUT[0]) == (10, 100, 68) and tuple(LUT[-1]) == (255, 255, 255)        # colormap_height.txt first / last row
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:240, 0:320]
    h = (20 + 15 * np.sin(xx / 40.0) * np.cos(yy / 55.0) + rng.normal(0, 0.2, xx.shape)).astype(np.float32)
    h[50:80, 100:160] = np.nan
    out = str(tmp_path / 'p.jpg')
    PH.plot_height_map(h, out, save_cbar=True)
    # independent statement of the mapping: percentile range, pins, x = (h - lo) / (hi - lo), index int(x * N) capped
    lo, hi = np.nanpercentile(h.astype(np.float64), [1, 99])
    c = np.clip(h.astype(np.float64), lo, hi)
    c[0, 0], c[0, 1] = lo, hi
    nan = np.isnan(c)
    x = (np.where(nan, lo, c) - lo) / (hi - lo)
    want = np.zeros(h.shape + (3,), dtype=np.uint8)
    for i in range(h.shape[0]):
        for j in range(h.shape[1]):
            want[i, j] = (0, 0, 0) if nan[i, j] else LUT[min(int(x[i, j] * 197), 196)]
    rgb, mask, rng_used = PH.height_to_rgb(h)
    assert np.array_equal(rgb, want) and np.array_equal(mask, nan) and rng_used == (lo, hi)
    assert tuple(rgb[0, 0]) == tuple(LUT[0]) and tuple(rgb[0, 1]) == tuple(LUT[196])
    got = cv2.imread(out)[:, :, ::-1].astype(np.float64)
    mse = np.mean((got - want.astype(np.float64)) ** 2)
    psnr = 10 * np.log10(255.0 ** 2 / mse)
    assert psnr > 30.0, psnr
    m = cv2.imread(str(tmp_path / 'p.mask.jpg'), cv2.IMREAD_GRAYSCALE)
    assert (m[60, 120] < 30) and (m[10, 10] > 225)
    assert os.path.exists(str(tmp_path / 'p.cbar.jpg'))
    # force_range and maskout
    rgb2, mask2, r2 = PH.height_to_rgb(h, force_range=(0.0, 40.0), maskout=(xx < 5))
    assert r2 == (0.0, 40.0) and mask2[:, :5].all() and (rgb2[:, :5] == 0).all()


def _write_tiled_tif(path, a, tile, bo='<', deflate=False):
    """Minimal tiled float32 TIFF (classic, one IFD) for the reader test."""
    import struct
    import zlib
    h, w = a.shape
    tl, tw = tile
    tiles = []
    for r0 in range(0, h, tl):
        for c0 in range(0, w, tw):
            t = np.zeros((tl, tw), dtype=np.float32)
            blk = a[r0:r0 + tl, c0:c0 + tw]
            t[:blk.shape[0], :blk.shape[1]] = blk
            raw = t.astype(bo + 'f4').tobytes()
            tiles.append(zlib.compress(raw) if deflate else raw)
    n = len(tiles)
    entries = [(256, 4, [w]), (257, 4, [h]), (258, 3, [32]), (259, 3, [8 if deflate else 1]), (262, 3, [1]), (277, 3, [1]),
               (322, 4, [tw]), (323, 4, [tl]), (339, 3, [3])]
    header = 8
    ifd_size = 2 + 12 * (len(entries) + 2) + 4
    arrays_off = header + ifd_size
    data_off = arrays_off + 8 * n
    offs, pos = [], data_off
    for t in tiles:
        offs.append(pos)
        pos += len(t)
    entries += [(324, 4, offs), (325, 4, [len(t) for t in tiles])]
    entries.sort()
    with open(path, 'wb') as fp:
        fp.write((b'II' if bo == '<' else b'MM') + struct.pack(bo + 'HI', 42, header))
        fp.write(struct.pack(bo + 'H', len(entries)))
        arr_pos = arrays_off
        blobs = b''
        for tag, typ, vals in entries:
            fmt = {3: 'H', 4: 'I'}[typ]
            payload = struct.pack(bo + str(len(vals)) + fmt, *vals)
            if len(payload) <= 4:
                fp.write(struct.pack(bo + 'HHI', tag, typ, len(vals)) + payload.ljust(4, b'\0'))
            else:
                fp.write(struct.pack(bo + 'HHII', tag, typ, len(vals), arr_pos))
                blobs += payload
                arr_pos += len(payload)
        fp.write(struct.pack(bo + 'I', 0))
        assert fp.tell() == arrays_off and len(blobs) == 8 * n
        fp.write(blobs)
        for t in tiles:
            fp.write(t)


def test_tif_reader_takes_compressed_predicted_and_tiled_files(tmp_path):
    """lib/dsm_util.read_dsm_tif on files other writers produce: Pillow/libtiff strips with LZW / Deflate / PackBits and
    predictors 1-3 (an independent encoder), and hand-written tiled files of both byte orders."""
    from vissatsatellitestereo_b200.lib.dsm_util import read_dsm_tif
    Image = pytest.importorskip('PIL.Image')
    rng = np.random.default_rng(7)
    a = (rng.normal(size=(70, 53)) * 100).astype(np.float32)
    a[3:9, 4:20] = 12.5                                   # runs, so that PackBits/LZW have something to find
    im = Image.fromarray(a, mode='F')
    for comp in (None, 'tiff_lzw', 'tiff_adobe_deflate', 'packbits'):
        for pred in (None, 2, 3):
            path = str(tmp_path / 'p_{}_{}.tif'.format(comp, pred))
            kw = {}
            if comp:
                kw['compression'] = comp
            if pred:
                kw['tiffinfo'] = {317: pred}
            im.save(path, **kw)
            got, meta = read_dsm_tif(path)
            assert np.array_equal(got, a), (comp, pred)
            assert (meta['img_height'], meta['img_width']) == a.shape and meta['zone_number'] is None
    for bo in ('<', '>'):
        for deflate in (False, True):
            path = str(tmp_path / 'tiled_{}_{}.tif'.format('le' if bo == '<' else 'be', int(deflate)))
            _write_tiled_tif(path, a, (32, 16), bo=bo, deflate=deflate)
            if not (bo == '>' and deflate):      # (Pillow mis-swaps big-endian data that went through its libtiff decoder)
                assert np.array_equal(np.array(Image.open(path)), a)      # the test writer itself is sane
            got, _ = read_dsm_tif(path)
            assert np.array_equal(got, a), (bo, deflate)


def test_ply_reader_takes_colmap_style_and_meshes(tmp_path):
    """ply2np on what the pipeline feeds it: COLMAP's fused.ply (x y z nx ny nz red green blue, binary little endian)
    and, for robustness, big-endian data, an ASCII file and a mesh whose face element follows the vertices."""
    from vissatsatellitestereo_b200.lib.ply_np_converter import ply2np
    rng = np.random.default_rng(3)
    n = 17
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, size=(n, 3)).astype(np.uint8)
    for bo, fmt in (('<', 'binary_little_endian'), ('>', 'binary_big_endian')):
        dt = np.dtype([(k, bo + 'f4') for k in ('x', 'y', 'z', 'nx', 'ny', 'nz')] + [(k, 'u1') for k in ('red', 'green', 'blue')])
        rec = np.empty(n, dtype=dt)
        for i, k in enumerate(('x', 'y', 'z')):
            rec[k] = xyz[:, i]
        for i, k in enumerate(('nx', 'ny', 'nz')):
            rec[k] = nrm[:, i]
        for i, k in enumerate(('red', 'green', 'blue')):
            rec[k] = rgb[:, i]
        header = ['ply', 'format {} 1.0'.format(fmt), 'comment made by a test', 'element vertex {}'.format(n)] + \
                 ['property float {}'.format(k) for k in ('x', 'y', 'z', 'nx', 'ny', 'nz')] + \
                 ['property uchar {}'.format(k) for k in ('red', 'green', 'blue')] + \
                 ['element face 2', 'property list uchar int vertex_indices', 'end_header']
        path = str(tmp_path / (fmt + '.ply'))
        with open(path, 'wb') as fp:
            fp.write(('\n'.join(header) + '\n').encode('ascii'))
            fp.write(rec.tobytes())
            fp.write(np.array([3], np.uint8).tobytes() + np.array([0, 1, 2], bo + 'i4').tobytes())
            fp.write(np.array([3], np.uint8).tobytes() + np.array([2, 1, 3], bo + 'i4').tobytes())
        data, color, comments = ply2np(path)
        assert np.array_equal(data, xyz) and np.array_equal(color, rgb) and comments == ['made by a test']
    path = str(tmp_path / 'ascii.ply')
    with open(path, 'w') as fp:
        fp.write('ply\nformat ascii 1.0\nelement vertex 2\nproperty double x\nproperty double y\nproperty double z\n'
                 'property uint8 red\nproperty uint8 green\nproperty uint8 blue\nelement face 1\n'
                 'property list uchar int vertex_indices\nend_header\n1.5 2.5 -3 1 2 3\n4 5 6.25 255 0 7\n3 0 1 1\n')
    data, color, comments = ply2np(path)
    assert np.array_equal(data, [[1.5, 2.5, -3], [4, 5, 6.25]]) and np.array_equal(color, [[1, 2, 3], [255, 0, 7]])
    assert comments is None
    with open(str(tmp_path / 'bad.ply'), 'w') as fp:
        fp.write('plx\n')
    with pytest.raises(ValueError):
        ply2np(str(tmp_path / 'bad.ply'))


def test_native_band_split_equals_numpy_array_split():
    """The row bands the stage-B peer stores use (exchange.cu: vs_band_rows, a host function of the shipped library)
    must be the ones distributed.row_bands (numpy.array_split) hands to the fusion, for every grid height and rank count."""
    import ctypes as C
    from vissatsatellitestereo_b200 import _native, distributed as D
    try:
        fn = getattr(_native.lib, '_Z12vs_band_rowsiiPi')
    except AttributeError:
        pytest.skip('internal symbol not exported by this build')
    fn.restype, fn.argtypes = None, [C.c_int, C.c_int, C.POINTER(C.c_int)]
    for n in range(1, _native.VS_MAX_RANKS + 1):
        for H in list(range(1, 70)) + [257, 2048, 8191]:
            row0 = (C.c_int * (n + 1))()
            fn(H, n, row0)
            assert row0[0] == 0 and row0[n] == H
            for j, (a, b) in enumerate(D.row_bands(H, n)):
                assert b - a == row0[j + 1] - row0[j]
                if b > a:
                    assert (a, b) == (row0[j], row0[j + 1])
