"""Host-side driver of the CUDA path for ONE GPU: depth maps -> per-view DSMs -> fused DSM.

torch is used for device buffers and streams only; every computation is a call into libvissat_b200.so.
The multi-GPU driver (one process per GPU) is in ``distributed.py``.
"""
import ctypes as C

import numpy as np
import torch

from . import _native
from ._native import lib, check

DEFAULT_SIMD_LANES = 16     # cv2's float32 vector width on AVX-512 hosts; only matters for grids narrower than lanes+2


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda():
    if not torch.cuda.is_available() or _native.device_count() == 0:
        raise _native.VisSatError('no CUDA device: vissatsatellitestereo_b200 has no CPU fallback')


def grid_shape(aoi_dict, e_resolution, n_resolution):
    """produce_dsm.py:54-55."""
    e_size = int(aoi_dict['width'] / e_resolution) + 1
    n_size = int(aoi_dict['height'] / n_resolution) + 1
    return e_size, n_size


def aoi_struct(aoi_dict, e_resolution, n_resolution, e_size=None, n_size=None, alt_margin=(50.0, 100.0)):
    """vs_aoi from aoi.json, following coordinate_system.py:45-47, produce_dsm.py:51-56 and the
    (swapped-name) resolution use of lib/proj_to_grid.py:42-43."""
    if e_size is None or n_size is None:
        e_size, n_size = grid_shape(aoi_dict, e_resolution, n_resolution)
    a = _native.vs_aoi()
    a.lat0 = (aoi_dict['lat_min'] + aoi_dict['lat_max']) / 2.0
    a.lon0 = (aoi_dict['lon_min'] + aoi_dict['lon_max']) / 2.0
    a.alt0 = aoi_dict['alt_min']
    a.zone = int(aoi_dict['zone_number'])
    a.south = 0 if aoi_dict['hemisphere'] == 'N' else 1
    a.ul_e = aoi_dict['ul_easting']
    a.ul_n = aoi_dict['ul_northing']
    a.row_res = e_resolution      # proj_to_grid(points, ul_e, ul_n, e_resolution, n_resolution, ...): rows / xresolution
    a.col_res = n_resolution      # cols / yresolution
    a.xsize = e_size
    a.ysize = n_size
    a.alt_lo = aoi_dict['alt_min'] - alt_margin[0]
    a.alt_hi = aoi_dict['alt_max'] + alt_margin[1]
    return a


class StepGraph:
    """A captured step of DsmEngine.capture_step."""

    def __init__(self, graph, fused, launches):
        self.graph, self.fused, self.launches_per_replay = graph, fused, int(launches)

    def replay(self):
        self.graph.replay()
        return self.fused


class DsmEngine:
    """One AOI on one GPU.  Buffers are allocated once and reused across views."""

    def __init__(self, aoi_dict, e_resolution=0.5, n_resolution=0.5, device=None, max_degree=5,
                 simd_lanes=DEFAULT_SIMD_LANES, ambiguity_eps=1e-7):
        require_cuda()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        self.aoi_dict = dict(aoi_dict)
        self.e_resolution, self.n_resolution = float(e_resolution), float(n_resolution)
        self.e_size, self.n_size = grid_shape(aoi_dict, e_resolution, n_resolution)
        self.simd_lanes = int(simd_lanes)
        self.collect_stats = True      # per-view counters (valid / in-grid / ambiguous / exact); one 32-byte memset
        self.ctx = _native.Context(self.device.index)
        self._aoi = aoi_struct(aoi_dict, e_resolution, n_resolution, self.e_size, self.n_size)
        self.fit = self.ctx.set_aoi(self._aoi, max_degree)
        self.ctx.set_ambiguity_eps(ambiguity_eps)
        with torch.cuda.device(self.device):
            self.keygrid = torch.empty((self.n_size, self.e_size), dtype=torch.int32, device=self.device)
            self._stats = torch.zeros(_native.VS_NUM_STATS, dtype=torch.int64, device=self.device)
            self._nan_count = torch.zeros(1, dtype=torch.int64, device=self.device)

    # ---- stage A + B -------------------------------------------------------------------------------------
    def rasterize(self, depth, inv_proj_mat, height_map=None, clear=True):
        """aggregate_2p5d_util.py:75-98 + lib/proj_to_grid.py:42-61 -> self.keygrid (device)."""
        assert depth.is_cuda and depth.dtype == torch.float32 and depth.dim() == 2 and depth.is_contiguous()
        M = np.ascontiguousarray(np.asarray(inv_proj_mat, dtype=np.float64).reshape(16))
        H, W = depth.shape
        check(lib.vs_unproject_rasterize(self.ctx.handle, _ptr(depth), H, W, M.ctypes.data_as(C.POINTER(C.c_double)),
                                         _ptr(self.keygrid), 1 if clear else 0, _ptr(height_map),
                                         _ptr(self._stats) if self.collect_stats else C.c_void_p(0),
                                         _stream(self.device)), 'vs_unproject_rasterize')

    def clear_keygrid(self):
        check(lib.vs_keygrid_clear(self.ctx.handle, _ptr(self.keygrid), self.keygrid.numel(), 4,
                                   _stream(self.device)), 'vs_keygrid_clear')

    def finalize(self, out=None, count_nan=False):
        """lib/proj_to_grid.py:62-79 + produce_dsm.py:58 -> float32 (n_size, e_size) per-view DSM (device)."""
        if out is None:
            out = torch.empty((self.n_size, self.e_size), dtype=torch.float32, device=self.device)
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and \
            tuple(out.shape) == (self.n_size, self.e_size)
        check(lib.vs_grid_finalize(self.ctx.handle, _ptr(self.keygrid), self.e_size, self.n_size, _ptr(out),
                                   self.simd_lanes, _ptr(self._nan_count) if count_nan else C.c_void_p(0),
                                   _stream(self.device)), 'vs_grid_finalize')
        return out

    def view_dsm(self, depth, inv_proj_mat, out=None, height_map=None):
        """One view of convert_depth_map_worker (aggregate_2p5d_util.py:75-102) without file I/O."""
        self.rasterize(depth, inv_proj_mat, height_map=height_map)
        return self.finalize(out=out)

    def views_to_dsm(self, depths, mats, stack, first=0, count_nan=None, stats=None):
        """Stages A + B for a batch of views in ONE library call (vs_views_to_dsm): view i of `depths` -> stack[first + i].
        depths: list of (H, W) float32 device tensors; mats: list of 4x4.  count_nan: optional int64 (n,) device tensor
        (empty cells per view); stats: optional int64 (n, VS_NUM_STATS) device tensor (K1 counters per view)."""
        n = len(depths)
        if n == 0:
            return
        ptrs = (C.c_void_p * n)(*[d.data_ptr() for d in depths])
        Hs = (C.c_int32 * n)(*[d.shape[0] for d in depths])
        Ws = (C.c_int32 * n)(*[d.shape[1] for d in depths])
        M = np.ascontiguousarray(np.stack([np.asarray(m, dtype=np.float64).reshape(16) for m in mats]))
        out = stack[first:first + n]
        assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape[1:]) == (self.n_size, self.e_size)
        check(lib.vs_views_to_dsm(self.ctx.handle, n, ptrs, Hs, Ws, M.ctypes.data_as(C.POINTER(C.c_double)),
                                  _ptr(self.keygrid), _ptr(out), stack.stride(0), self.simd_lanes,
                                  _ptr(count_nan), _ptr(stats), _stream(self.device)), 'vs_views_to_dsm')

    def capture_step(self, depths, mats, stack, fuse=True, occ=None):
        """Stages A-C for a fixed set of device buffers as ONE CUDA graph (102 kernel launches + 50 memsets for 50 views): replaying
        it costs one launch call on the host and removes the per-launch gaps on the device.  The step is run eagerly
        once first (every lazy allocation inside the library happens there), then captured; the internal streams of
        vs_views_to_dsm fork from and join into the capturing stream, so the overlap of stage A and stage B is part of
        the graph.  Returns a StepGraph; `replay()` returns the fused DSM tensor (same storage every time)."""
        self.set_timing(False)

        def body():
            if occ is not None:
                occ.zero_()
                self.set_occupancy(occ, stack, 0)
            self.views_to_dsm(depths, mats, stack)
            if not fuse:
                return None
            return self.median3x3(self.fuse(stack, occ=occ), count_nan=True)

        body()
        torch.cuda.synchronize(self.device)
        n0 = self.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fused = body()
        return StepGraph(graph, fused, self.launch_count() - n0)

    # ---- occupancy bitmap (sparse coverage: large AOIs) -----------------------------------------------------------
    def occupancy_shape(self, n_views):
        """(tile rows, tile columns, words) of the bitmap for n_views views (vs_set_occupancy layout)."""
        return (-(-self.n_size // _native.VS_TILE_H), -(-self.e_size // _native.VS_TILE_W), -(-int(n_views) // 32))

    def alloc_occupancy(self, n_views):
        return torch.zeros(self.occupancy_shape(n_views), dtype=torch.int32, device=self.device)

    def set_occupancy(self, occ, stack=None, view0=0):
        """Make stage B (views_to_dsm) record in `occ` which (tile, view) pairs of `stack` hold data; plane i of `stack`
        is global view view0 + i.  occ=None switches it off.  The caller zeroes `occ` before each pass over the views."""
        if occ is None:
            check(lib.vs_set_occupancy(self.ctx.handle, C.c_void_p(0), 0, C.c_void_p(0), 0), 'vs_set_occupancy')
            return
        assert occ.is_cuda and occ.dtype == torch.int32 and occ.is_contiguous() and occ.dim() == 3
        assert tuple(occ.shape[:2]) == self.occupancy_shape(1)[:2]
        check(lib.vs_set_occupancy(self.ctx.handle, _ptr(occ), occ.shape[2], _ptr(stack), int(view0)), 'vs_set_occupancy')

    def set_exchange(self, ex):
        """Enable (a _native.vs_exchange) or disable (None) the peer stores of stage B (distributed.PeerExchange)."""
        check(lib.vs_set_exchange(self.ctx.handle, C.byref(ex) if ex is not None else None), 'vs_set_exchange')

    def set_streams(self, n):
        check(lib.vs_set_streams(self.ctx.handle, int(n)), 'vs_set_streams')

    def set_coschedule(self, enable):
        """Co-scheduled stage A+B kernel of views_to_dsm on (default) / off (the separate stage-A and stage-B kernels)."""
        check(lib.vs_set_coschedule(self.ctx.handle, 1 if enable else 0), 'vs_set_coschedule')

    def set_timing(self, enable):
        check(lib.vs_set_timing(self.ctx.handle, 1 if enable else 0), 'vs_set_timing')

    def get_timing(self, max_views=65536):
        a = (C.c_float * max_views)()
        b = (C.c_float * max_views)()
        n = C.c_int32(0)
        check(lib.vs_get_timing(self.ctx.handle, max_views, a, b, C.byref(n)), 'vs_get_timing')
        k = min(n.value, max_views)
        return np.array(a[:k], dtype=np.float64), np.array(b[:k], dtype=np.float64)

    def stats(self):
        """Counters of the last rasterize call (synchronises)."""
        s = self._stats.cpu().numpy()
        return {'valid': int(s[0]), 'in_grid': int(s[1]), 'ambiguous': int(s[2]), 'exact': int(s[3])}

    def last_nan_count(self):
        return int(self._nan_count.cpu().item())

    # ---- stage C ------------------------------------------------------------------------------------------
    def fuse(self, views, out=None, occ=None, row0=0):
        """aggregate_2p5d.py:65-78 on a (V, rows, W) float32 stack of per-view DSMs (device).
        occ: optional occupancy bitmap filled by stage B (set_occupancy / the sparse peer exchange): only the marked
        (tile, view) pairs are read, the rest count as NaN; `views` then holds grid rows [row0, row0 + rows)."""
        assert views.is_cuda and views.dtype == torch.float32 and views.dim() == 3
        V, rows, W = views.shape
        assert views.stride(2) == 1 and views.stride(1) == W, 'planes must be row-major contiguous'
        if out is None:
            out = torch.empty((rows, W), dtype=torch.float32, device=self.device)
        if occ is not None:
            assert occ.is_cuda and occ.dtype == torch.int32 and occ.is_contiguous() and occ.dim() == 3
            assert occ.shape[1] == -(-W // _native.VS_TILE_W) and occ.shape[0] * _native.VS_TILE_H >= row0 + rows
            check(lib.vs_fuse_views_sparse(self.ctx.handle, _ptr(views), views.stride(0), V, rows, W, int(row0), _ptr(occ),
                                           occ.shape[2], _ptr(out), _stream(self.device)), 'vs_fuse_views_sparse')
            return out
        check(lib.vs_fuse_views(self.ctx.handle, _ptr(views), views.stride(0), V, rows, W, _ptr(out),
                                _stream(self.device)), 'vs_fuse_views')
        return out

    def densify(self, views, occ, row0=0):
        """Copy of `views` with every (tile, view) pair the bitmap does not mark set to NaN -- what the stack means under
        the sparse convention (diagnostics / tests; plain torch indexing, not a hot path)."""
        V, rows, W = views.shape
        g = torch.arange(V, device=views.device)
        bits = (occ[:, :, (g // 32)] >> (g % 32).to(torch.int32)) & 1                       # (Ty, Tx, V)
        m = bits.permute(2, 0, 1).bool()
        m = m.repeat_interleave(_native.VS_TILE_H, dim=1)[:, row0:row0 + rows]
        m = m.repeat_interleave(_native.VS_TILE_W, dim=2)[:, :, :W]
        return torch.where(m, views, torch.full_like(views, float('nan')))

    def median3x3(self, img, out=None, row_begin=0, row_end=None, in_row0=0, h_total=None, count_nan=False):
        """aggregate_2p5d.py:81: cv2.medianBlur(float32, 3) of rows [row_begin, row_end) of an h_total-row image
        whose rows in_row0.. are in `img`."""
        assert img.is_cuda and img.dtype == torch.float32 and img.dim() == 2 and img.is_contiguous()
        W = img.shape[1]
        if h_total is None:
            h_total = in_row0 + img.shape[0]
        if row_end is None:
            row_end = h_total
        if out is None:
            out = torch.empty((row_end - row_begin, W), dtype=torch.float32, device=self.device)
        check(lib.vs_median3x3(self.ctx.handle, _ptr(img), in_row0, h_total, W, row_begin, row_end, _ptr(out),
                               self.simd_lanes, _ptr(self._nan_count) if count_nan else C.c_void_p(0),
                               _stream(self.device)), 'vs_median3x3')
        return out

    def fuse_and_blur(self, views):
        """aggregate_2p5d.py:65-81 -> float32 (n_size, e_size) fused DSM (device)."""
        mean = self.fuse(views)
        return self.median3x3(mean, count_nan=True)

    # ---- host-buffer entry point (what aggregate_2p5d.run_fuse uses) ---------------------------------------
    def process_host(self, depths_host, mats, views_out_host=None, fused_out_host=None, stack=None, fuse=True):
        """Depth maps in (pinned) host memory -> per-view DSMs and fused DSM back in host memory.

        Three streams: H2D of view v+1, kernels of view v and D2H of DSM v-1 overlap; two device depth slots.
        depths_host: list of float32 (H,W) CPU tensors (pinned for async copies); mats: list of 4x4 arrays.
        views_out_host: optional (V, n_size, e_size) pinned CPU tensor; fused_out_host: optional (n_size, e_size).
        Returns (stack (device), fused (device or None)).  Synchronises before returning."""
        V = len(depths_host)
        dev = self.device
        if stack is None:
            stack = torch.empty((V, self.n_size, self.e_size), dtype=torch.float32, device=dev)
        if not hasattr(self, '_streams'):
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        s_in, s_comp, s_out = self._streams
        cur = torch.cuda.current_stream(dev)
        for s in self._streams:
            s.wait_stream(cur)
        shapes = {tuple(d.shape) for d in depths_host}
        slots = {}
        for shp in shapes:
            key = ('slot', shp)
            if key not in self.__dict__.setdefault('_slots', {}):
                self._slots[key] = [torch.empty(shp, dtype=torch.float32, device=dev) for _ in range(2)]
            slots[shp] = self._slots[key]
        comp_done = [None] * V
        for v in range(V):
            shp = tuple(depths_host[v].shape)
            slot = slots[shp][v % 2]
            if v >= 2:
                s_in.wait_event(comp_done[v - 2])
            with torch.cuda.stream(s_in):
                slot.copy_(depths_host[v], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            s_comp.wait_event(ev_in)
            with torch.cuda.stream(s_comp):
                self.view_dsm(slot, mats[v], out=stack[v])
                comp_done[v] = torch.cuda.Event()
                comp_done[v].record(s_comp)
            if views_out_host is not None:
                s_out.wait_event(comp_done[v])
                with torch.cuda.stream(s_out):
                    views_out_host[v].copy_(stack[v], non_blocking=True)
        fused = None
        if fuse:
            with torch.cuda.stream(s_comp):
                fused = self.fuse_and_blur(stack)
                ev = torch.cuda.Event()
                ev.record(s_comp)
            if fused_out_host is not None:
                s_out.wait_event(ev)
                with torch.cuda.stream(s_out):
                    fused_out_host.copy_(fused, non_blocking=True)
        for s in self._streams:
            cur.wait_stream(s)
        cur.synchronize()
        return stack, fused

    def eval_poly(self, enu):
        """Diagnostics (vs_fit_eval): the validated ENU -> (fractional col, fractional row, altitude) polynomial of this
        AOI on (n, 3) host points, in K1's arithmetic."""
        pts = np.ascontiguousarray(np.asarray(enu, dtype=np.float64).reshape(-1, 3))
        n = pts.shape[0]
        out = [np.empty(n, dtype=np.float64) for _ in range(3)]
        dp = C.POINTER(C.c_double)
        check(lib.vs_fit_eval(self.ctx.handle, pts.ctypes.data_as(dp), n, *[o.ctypes.data_as(dp) for o in out]),
              'vs_fit_eval')
        return tuple(out)

    def launch_count(self):
        return self.ctx.launch_count()

    def close(self):
        self.ctx.close()


# ---- AOI-independent entry points (public proj_to_grid / converters) -----------------------------------------
_ctx_cache = {}


def default_context(device=None):
    require_cuda()
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _ctx_cache:
        _ctx_cache[idx] = _native.Context(idx)
    return _ctx_cache[idx], torch.device('cuda', idx)


def proj_to_grid_device(points, xoff, yoff, xresolution, yresolution, xsize, ysize, blur=False,
                        simd_lanes=DEFAULT_SIMD_LANES, device=None):
    """lib/proj_to_grid.py:41-81 on the GPU with exact float64 semantics.
    points: (N,3) float64 (host numpy or device tensor).  Returns device float64 (ysize, xsize)
    [, device float32 blurred grid if blur]."""
    ctx, dev = default_context(device)
    if not torch.is_tensor(points):
        points = torch.from_numpy(np.ascontiguousarray(np.asarray(points, dtype=np.float64)))
    pts = points.to(device=dev, dtype=torch.float64).contiguous()
    assert pts.dim() == 2 and pts.shape[1] >= 3
    if pts.shape[1] != 3:
        pts = pts[:, :3].contiguous()
    n = pts.shape[0]
    keys = torch.empty((ysize, xsize), dtype=torch.int64, device=dev)
    filled = torch.empty((ysize, xsize), dtype=torch.float64, device=dev)
    blurred = torch.empty((ysize, xsize), dtype=torch.float32, device=dev) if blur else None
    st = _stream(dev)
    check(lib.vs_points_rasterize(ctx.handle, _ptr(pts), n, float(xoff), float(yoff), float(xresolution),
                                  float(yresolution), int(xsize), int(ysize), _ptr(keys), 1, C.c_void_p(0), st),
          'vs_points_rasterize')
    check(lib.vs_grid_finalize64(ctx.handle, _ptr(keys), int(xsize), int(ysize), _ptr(filled), _ptr(blurred),
                                 int(simd_lanes), st), 'vs_grid_finalize64')
    return (filled, blurred) if blur else filled
