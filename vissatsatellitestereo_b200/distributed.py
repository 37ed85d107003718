"""Multi-GPU driver: one process per GPU (torchrun), NCCL over NVLink for the one exchange the path has.

SURVEY.md §8(e): stages A+B are independent per view -> the sorted view list is split into contiguous blocks,
one per rank (so the fused sum order stays the reference's sorted order, aggregate_2p5d.py:59).  Stage C is
independent per cell -> each rank owns a contiguous band of grid rows (+1 halo row each side for the final
3x3 blur).  The transpose between the two layouts is an all-to-all of per-view DSM row bands; it replaces the
file-system hand-off of aggregate_2p5d.py:57-66.  The result is independent of the number of ranks.

The host logic here runs on any backend (tests use gloo on CPU tensors); the compute calls need CUDA.

Two transports for that transpose:
  * `PeerExchange` (default on NVLink boxes): stage B's kernel stores every output row straight into the band stack of
    the rank that fuses it, through CUDA-IPC peer-mapped memory -- compute and exchange are one kernel, there is no
    pack/copy pass and no collective on the data path (only a 1-element all-reduce as the stream-ordered barrier);
  * `WaveExchanger`: grouped NCCL send/recv in waves, overlapped with stages A/B of later views (also what the gloo CPU
    tests exercise).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def split_views(n_views, world):
    """Contiguous blocks like np.array_split (aggregate_2p5d_util.py:109-122 uses the same split)."""
    idx = np.array_split(np.arange(n_views), world)
    return [(int(a[0]), int(a[-1]) + 1) if a.size else (0, 0) for a in idx]


def row_bands(n_rows, world):
    idx = np.array_split(np.arange(n_rows), world)
    return [(int(a[0]), int(a[-1]) + 1) if a.size else (n_rows, n_rows) for a in idx]


def band_with_halo(band, n_rows, halo=1):
    r0, r1 = band
    if r0 == r1:
        return (r0, r1)
    return (max(r0 - halo, 0), min(r1 + halo, n_rows))


def pack_rowbands(local_stack, bands_h):
    """(V_local, n_rows, W) -> flat send buffer laid out [dest rank][view][band rows + halo][W]."""
    parts = [local_stack[:, a:b, :].reshape(-1) for (a, b) in bands_h]
    return torch.cat(parts) if len(parts) > 1 else parts[0].contiguous()


def exchange_rowbands(local_stack, view_counts, n_rows, group=None, halo=1):
    """All-to-all of per-view DSM row bands.

    local_stack : (V_local, n_rows, W) float32, this rank's per-view DSMs (its block of the sorted view list)
    view_counts : list of V_local for every rank
    Returns (band_stack (V_total, rows_with_halo, W), (r0, r1) own band, (h0, h1) rows present in band_stack).
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    W = local_stack.shape[2]
    bands = row_bands(n_rows, world)
    bands_h = [band_with_halo(b, n_rows, halo) for b in bands]
    h0, h1 = bands_h[rank]
    my_rows = h1 - h0
    send = pack_rowbands(local_stack, bands_h)
    in_splits = [local_stack.shape[0] * (b - a) * W for (a, b) in bands_h]
    out_splits = [vc * my_rows * W for vc in view_counts]
    recv = torch.empty(sum(out_splits), dtype=local_stack.dtype, device=local_stack.device)
    if world == 1:
        recv.copy_(send)
    else:
        dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
    v_total = sum(view_counts)
    return recv.view(v_total, my_rows, W), bands[rank], (h0, h1)


class WaveExchanger:
    """The same all-to-all, issued in waves on a side stream so that it overlaps stages A/B of later views.

    Rank r calls `send_wave(a, b)` once views [a, b) of its local stack are final on the current stream; the wave is
    packed and exchanged on `self.stream`.  `finish()` makes the current stream wait for every wave and returns the
    (V_total, rows_with_halo, W) stack of this rank's row band, in global (sorted) view order.
    """

    def __init__(self, local_stack, view_counts, group=None, halo=1):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local = local_stack
        self.view_counts = list(view_counts)
        self.n_rows, self.W = local_stack.shape[1], local_stack.shape[2]
        self.bands = row_bands(self.n_rows, self.world)
        self.bands_h = [band_with_halo(b, self.n_rows, halo) for b in self.bands]
        self.h0, self.h1 = self.bands_h[self.rank]
        self.first = [sum(self.view_counts[:i]) for i in range(self.world)]
        v_total = sum(self.view_counts)
        self.band_stack = torch.empty((v_total, self.h1 - self.h0, self.W), dtype=local_stack.dtype,
                                      device=local_stack.device)
        self.stream = torch.cuda.Stream(device=local_stack.device) if local_stack.is_cuda else None
        # persistent pack buffers, one per destination rank: [view][band rows + halo][W].  Allocated once: the object
        # is reused for every pass over the data (a fresh allocation per pass on a side stream would fall through the
        # caching allocator to cudaMalloc and synchronise the device).
        self.send_bufs = [torch.empty((local_stack.shape[0], h1 - h0, self.W), dtype=local_stack.dtype,
                                      device=local_stack.device) for (h0, h1) in self.bands_h]
        self._reqs = []

    def send_wave(self, a, b):
        """Exchange local views [a, b).  Every rank must call this with the same sequence of (a, b) (view counts are
        equal per rank in the benchmark; ranks with fewer views pass empty ranges clipped to their count)."""
        ranges = [(min(a, c), min(b, c)) for c in self.view_counts]      # views [a, b) of every source rank
        la, lb = ranges[self.rank]
        cur = torch.cuda.current_stream(self.local.device) if self.stream is not None else None
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _NullCtx()
        if self.stream is not None:
            self.stream.wait_stream(cur)
        with ctx:
            send = []
            for buf, (h0, h1) in zip(self.send_bufs, self.bands_h):
                buf[la:lb].copy_(self.local[la:lb, h0:h1, :])          # pack (strided -> contiguous)
                send.append(buf[la:lb])
            recv = [self.band_stack[self.first[i] + ra: self.first[i] + rb] for i, (ra, rb) in enumerate(ranges)]
            # own band: local copy (all_to_all on NCCL rewrites it with the same data); peers: grouped send/recv
            if recv[self.rank].numel():
                recv[self.rank].copy_(send[self.rank])
            ops = []
            if self.world > 1 and self.local.is_cuda:
                # NCCL: the list form of all_to_all is one grouped ncclSend/ncclRecv batch on the main communicator
                dist.all_to_all(recv, send, group=self.group)
            elif self.world > 1:
                # gloo (CPU tests) has no list all_to_all: batched isend/irecv
                for peer in range(self.world):
                    if peer == self.rank:
                        continue
                    if send[peer].numel():
                        ops.append(dist.P2POp(dist.isend, send[peer], peer, group=self.group))
                    if recv[peer].numel():
                        ops.append(dist.P2POp(dist.irecv, recv[peer], peer, group=self.group))
            if ops:
                self._reqs.extend(dist.batch_isend_irecv(ops))

    def finish(self):
        ctx = torch.cuda.stream(self.stream) if self.stream is not None else _NullCtx()
        with ctx:
            for q in self._reqs:
                q.wait()             # on NCCL: stream-level wait on the side stream; on gloo: host wait
        if self.stream is not None:
            torch.cuda.current_stream(self.local.device).wait_stream(self.stream)
        self._reqs = []
        return self.band_stack, self.bands[self.rank], (self.h0, self.h1)


class _DeviceArray:
    """Device memory owned by libvissat_b200 (vs_peer_alloc / vs_peer_open) exposed through
    __cuda_array_interface__ so that torch can wrap it without a copy."""

    def __init__(self, ptr, shape, typestr='<f4'):
        self.__cuda_array_interface__ = {'shape': tuple(int(x) for x in shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}


class PeerExchange:
    """Row-band exchange written by stage B itself (vs_set_exchange, include/vissat_b200.h).

    Every rank owns `buffers` band stacks (V_total, rows + halo, W) allocated by the library and mapped into every
    other rank's address space (CUDA IPC; the handles travel through all_gather_object).  `begin_step()` points the
    engine at the next stack of every rank; the following `engine.views_to_dsm(..., local_stack, ...)` calls then
    write both the local per-view DSMs and the remote row bands.  `finish()` is the stream-ordered barrier (1-element
    all-reduce: every rank's stage-B kernels have completed, hence their peer stores have landed) and returns this
    rank's stack.  Two stacks alternate so that step s+1 may write while a slower rank still fuses step s.
    """

    def __init__(self, engine, local_stack, view_counts, group=None, halo=1, buffers=2, sparse=True):
        """sparse=True (default): stage B also maintains the owners' occupancy bitmaps and does not send all-empty
        tiles; the owner fuses with vs_fuse_views_sparse ("skip all-NaN row-bands via an occupancy
        flag").  sparse=False: every tile is stored, the band stacks are dense."""
        from . import _native
        self._native = _native
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > _native.VS_MAX_RANKS:
            raise ValueError('PeerExchange supports up to {} ranks'.format(_native.VS_MAX_RANKS))
        self.engine = engine
        self.local = local_stack
        self.view_counts = list(view_counts)
        self.halo = int(halo)
        self.n_rows, self.W = local_stack.shape[1], local_stack.shape[2]
        assert (self.n_rows, self.W) == (engine.n_size, engine.e_size)
        self.bands = row_bands(self.n_rows, self.world)
        self.bands_h = [band_with_halo(b, self.n_rows, halo) for b in self.bands]
        self.h0, self.h1 = self.bands_h[self.rank]
        self.v_total = sum(self.view_counts)
        self.view0 = sum(self.view_counts[:self.rank])
        self.n_buffers = int(buffers)
        lib, check = _native.lib, _native.check
        ctx = engine.ctx.handle
        my_rows = self.h1 - self.h0
        self.sparse = bool(sparse)
        self.occ_shape = engine.occupancy_shape(self.v_total)                 # full-grid tile indexing on every rank
        # the bitmap sits behind the stack in the same allocation, at an offset that is THE SAME ON EVERY RANK (peers
        # compute each other's bitmap address from it): the largest band stack of any rank, rounded up
        stack_bytes = max(max(self.v_total * (h1 - h0) * self.W * 4 for (h0, h1) in self.bands_h), 4)
        self._occ_offset = (stack_bytes + 255) // 256 * 256
        occ_bytes = int(np.prod(self.occ_shape)) * 4
        nbytes = self._occ_offset + occ_bytes
        # Set-up failures must be seen by every rank (a rank that raised alone would leave the others in a collective):
        # each phase ends with an exchange of success flags.
        self._own, self._opened, self._ptrs, handles, err = [], [], [], [], None
        try:
            for _ in range(self.n_buffers):
                ptr = C.c_void_p()
                hbuf = (C.c_uint8 * _native.VS_IPC_HANDLE_BYTES)()
                check(lib.vs_peer_alloc(ctx, nbytes, C.byref(ptr), hbuf), 'vs_peer_alloc')
                self._own.append(ptr.value)
                handles.append(bytes(hbuf))
        except Exception as e:
            err, handles = e, None
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        if any(g is None for g in gathered):
            self._release()
            raise _native.VisSatError('PeerExchange: allocation failed on rank(s) {}: {}'.format(
                [r for r, g in enumerate(gathered) if g is None], err))
        try:
            for b in range(self.n_buffers):   # [buffer][rank] -> device pointer valid in this process
                row = []
                for r in range(self.world):
                    if r == self.rank:
                        row.append(self._own[b])
                        continue
                    ptr = C.c_void_p()
                    hbuf = (C.c_uint8 * _native.VS_IPC_HANDLE_BYTES).from_buffer_copy(gathered[r][b])
                    check(lib.vs_peer_open(ctx, hbuf, C.byref(ptr)), 'vs_peer_open')
                    self._opened.append(ptr.value)
                    row.append(ptr.value)
                self._ptrs.append(row)
        except Exception as e:
            err = e
        # also the barrier "every rank has mapped every stack before anyone writes"
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=local_stack.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) != 1:
            self._release()
            raise _native.VisSatError('PeerExchange: mapping peer memory failed on some rank: {}'.format(err))
        self.band_stacks = [torch.as_tensor(_DeviceArray(p, (self.v_total, my_rows, self.W)), device=local_stack.device)
                            for p in self._own]
        self.occs = [torch.as_tensor(_DeviceArray(p + self._occ_offset, self.occ_shape, typestr='<i4'),
                                     device=local_stack.device) for p in self._own]
        for o in self.occs:
            o.zero_()
        torch.cuda.synchronize(local_stack.device)
        dist.barrier(group=group)              # every bitmap is zero before any peer sets a bit
        self._flag = torch.zeros(1, dtype=torch.int32, device=local_stack.device)
        self._cur = -1
        self.keep_last_occ = False     # diagnostics (check_band_stack)

    def _release(self):
        lib, ctx = self._native.lib, self.engine.ctx.handle
        for p in self._opened:
            lib.vs_peer_close(ctx, C.c_void_p(p))
        self._opened = []
        for p in self._own:
            lib.vs_peer_free(ctx, C.c_void_p(p))
        self._own = []

    def begin_step(self):
        self._cur = (self._cur + 1) % self.n_buffers
        self._barrier_done = False
        if self.local.shape[0] == 0:           # a rank without views launches nothing; it still takes part in finish()
            return
        ex = self._native.vs_exchange()
        ex.n_ranks, ex.rank, ex.halo = self.world, self.rank, self.halo
        ex.view0, ex.n_views_total = self.view0, self.v_total
        ex.local_stack = self.local.data_ptr()
        ex.occ_words = self.occ_shape[2] if self.sparse else 0
        for r in range(self.world):
            ex.band_stack[r] = self._ptrs[self._cur][r]
            ex.occ[r] = self._ptrs[self._cur][r] + self._occ_offset if self.sparse else None
        self.engine.set_exchange(ex)

    def finish_barrier_only(self):
        """The stream-ordered barrier of the step ("every rank's stage-B kernels, hence their peer stores, are complete")."""
        if not getattr(self, '_barrier_done', False):
            self.engine.set_exchange(None)
            dist.all_reduce(self._flag, group=self.group)
            self._barrier_done = True

    def finish(self):
        """Barrier on the current stream, then this rank's (V_total, rows + halo, W) stack of the step."""
        self.finish_barrier_only()
        return self.band_stacks[self._cur], self.bands[self.rank], (self.h0, self.h1)

    def fuse_band(self, count_nan=True, out=None, after_barrier=False):
        """finish() + stage C on this rank's rows: (fused band (rows, W), (r0, r1)).  In sparse mode only the marked
        (tile, view) pairs of the stack are read, and the bitmap is cleared for the step after next."""
        eng = self.engine
        band_stack, (r0, r1), (h0, h1) = self.finish()
        if r1 == r0:
            return torch.empty((0, self.W), dtype=torch.float32, device=self.local.device), (r0, r1)
        if not hasattr(self, '_mean_buf'):
            self._mean_buf = torch.empty((h1 - h0, self.W), dtype=torch.float32, device=self.local.device)
        occ = self.occs[self._cur] if self.sparse else None
        mean = eng.fuse(band_stack, out=self._mean_buf, occ=occ, row0=h0)
        if occ is not None:
            if self.keep_last_occ:
                self._last_occ = occ.clone()
            occ.zero_()      # before this rank reaches the next barrier, i.e. before any peer writes this buffer again
        out = eng.median3x3(mean, out=out, row_begin=r0, row_end=r1, in_row0=h0, h_total=self.n_rows, count_nan=count_nan)
        return out, (r0, r1)

    def check_band_stack(self, want_dense):
        """Diagnostics: the current stack, read under the sparse convention, equals the dense stack `want_dense`.
        Call between finish() and the next begin_step(); in sparse mode BEFORE fuse_band() clears the bitmap -- so
        fuse_band() keeps a copy of the last bitmap for this check."""
        stack = self.band_stacks[self._cur]
        if self.sparse:
            occ = getattr(self, '_last_occ', None)
            if occ is None:
                return False
            stack = self.engine.densify(stack, occ, row0=self.h0)
        return stack.shape == want_dense.shape and torch.equal(torch.nan_to_num(stack, nan=-1e9),
                                                               torch.nan_to_num(want_dense, nan=-1e9))

    def close(self):
        lib, ctx = self._native.lib, self.engine.ctx.handle
        torch.cuda.synchronize(self.local.device)
        self.engine.set_exchange(None)
        dist.barrier(group=self.group)         # nobody is still writing into a stack that is about to be unmapped
        self.band_stacks = []
        self.occs = []
        self._last_occ = None
        for p in self._opened:
            lib.vs_peer_close(ctx, C.c_void_p(p))
        self._opened = []
        dist.barrier(group=self.group)         # every mapping is gone before the owner frees the memory
        for p in self._own:
            lib.vs_peer_free(ctx, C.c_void_p(p))
        self._own = []


def make_exchange(engine, local_stack, view_counts, group=None, prefer='peer', sparse=True):
    """PeerExchange when every rank can set it up (its constructor fails on all ranks or on none), else WaveExchanger.
    Returns (exchanger, kind)."""
    if prefer == 'peer' and local_stack.is_cuda:
        try:
            return PeerExchange(engine, local_stack, view_counts, group=group, sparse=sparse), 'peer-store'
        except Exception as e:      # e.g. no peer access between the devices, IPC not permitted in this container
            if dist.get_rank(group) == 0:
                print('PeerExchange unavailable ({}); using NCCL waves'.format(e), flush=True)
    return WaveExchanger(local_stack, view_counts, group=group), 'nccl-waves'


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def fuse_distributed(engine, local_stack, view_counts, group=None):
    """Stage C across ranks: exchange, fuse own row band, blur it.  Returns (band float32 (rows, W), (r0, r1))."""
    n_rows = local_stack.shape[1]
    band_stack, (r0, r1), (h0, h1) = exchange_rowbands(local_stack, view_counts, n_rows, group=group)
    if r1 == r0:
        return torch.empty((0, local_stack.shape[2]), dtype=torch.float32, device=local_stack.device), (r0, r1)
    mean = engine.fuse(band_stack)
    out = engine.median3x3(mean, row_begin=r0, row_end=r1, in_row0=h0, h_total=n_rows, count_nan=True)
    return out, (r0, r1)


def gather_bands(band, n_rows, W, group=None, dst=0):
    """Collect the fused row bands on rank `dst` -> (n_rows, W) there, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bands = row_bands(n_rows, world)
    if world == 1:
        return band
    if rank == dst:
        outs = [torch.empty((b - a, W), dtype=band.dtype, device=band.device) for (a, b) in bands]
        outs[dst] = band
        reqs = [dist.irecv(outs[r], src=r, group=group) for r in range(world) if r != dst and outs[r].numel()]
        for q in reqs:
            q.wait()
        return torch.cat(outs, dim=0)
    if band.numel():
        dist.send(band.contiguous(), dst=dst, group=group)
    return None
