#!/usr/bin/env python
"""bench.py -- depth Mpix/s -> fused DSM (BASELINE.json metric) for the B200 path and the CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One "step" = one pass of the whole hot path over one batch of synthetic views: every view through stage A
(unproject + geodesy + scatter-max) and stage B (hole fill + 3x3 median) into a per-view DSM, then stage C
(robust cross-view fusion + 3x3 median) into the fused DSM.  At N = 1 the workload is BASELINE.json configs[1]
(C2: 50 views x 2048^2 depth, 2048^2 grid @ 0.3 m).  At N > 1 every rank processes its own C2-sized block of views
(weak scaling: 50*N views in total), the per-view DSM row bands are exchanged with one NCCL all-to-all and every
rank fuses its own band of grid rows.

`value`  : Mpix/s with the depth maps already resident in HBM (device -> device).
`e2e`    : same metric through the host-buffer entry point (pinned host depth maps -> H2D -> kernels -> D2H of
           every per-view DSM and of the fused DSM).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = 'depth_mpix_per_s_to_fused_dsm'
UNIT = 'Mpix/s'


def load_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fp:
            return float(json.load(fp)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every ~5 ms during the timed region (the region is a few
    tens of ms, too short for `nvidia-smi -lms`); same fields as the B200_PROFILING.md clocks line."""

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.thread = None
        self.stop_flag = False
        self.smax = None

    def _run(self):
        import pynvml as nv
        R = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown,
             'hw_thermal_slowdown': nv.nvmlClocksEventReasonHwThermalSlowdown,
             'sw_thermal_slowdown': nv.nvmlClocksEventReasonSwThermalSlowdown,
             'sw_power_cap': nv.nvmlClocksEventReasonSwPowerCap}
        h = self.handle
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((sm, pw, [k for k, bit in R.items() if mask & bit]))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(',')[self.index])
                except Exception:
                    pass
            self.handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.smax = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': self.smax, 'reasons': [], 'samples': 0}
        if self.thread is None:
            return out
        self.stop_flag = True
        self.thread.join(timeout=2)
        if self.samples:
            sm = np.array([s[0] for s in self.samples], dtype=float)
            reasons = sorted({r for s in self.samples for r in s[2]})
            out = {'sm_mhz': float(np.median(sm)), 'sm_min_mhz': float(sm.min()), 'sm_max_mhz': float(self.smax),
                   'reasons': reasons, 'samples': len(self.samples),
                   'power_w_max': float(max(s[1] for s in self.samples))}
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_view_worker(args):
    from oracle import pipeline as op
    depth, M, aoi, res = args
    dsm, _ = op.convert_depth_map(depth, M, aoi, res, res, fast=False)     # faithful: Python hole-fill loop
    return dsm


def cpu_sample(cfg, n_views, band_rows, cores, geo=None, shrink=1):
    """Time the oracle (CPU restatement of the reference) on a bounded sample of the workload and extrapolate
    linearly to the whole config: n_views views through the per-view stage with a
    multiprocessing.Pool(cores) exactly like aggregate_2p5d_util.py:138-146, then the single-process fusion
    (aggregate_2p5d.py:65-81) on band_rows grid rows of a V-view cube."""
    import multiprocessing as mp
    from oracle import geodesy, pipeline as op
    from vissatsatellitestereo_b200 import synthetic as S
    full = cfg
    if shrink > 1:
        # same recipe on a sub-AOI: depth maps and grid both 1/shrink in each dimension (same GSD, same density);
        # every stage is linear in pixels / cells, so the throughput is extrapolated by area
        cfg = S.scaled(cfg, depth=cfg.height // shrink, grid=cfg.e_size // shrink, name=cfg.name)
    scene = S.make_scene(cfg, geodesy, device='cpu', views=range(n_views))
    jobs = [(d.numpy(), M, scene.aoi, cfg.res) for d, M in zip(scene.depths, scene.mats)]
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(min(cores, n_views)) as pool:
        dsms = pool.map(_cpu_view_worker, jobs, chunksize=1)
    t_views = time.perf_counter() - t0
    t_fuse = 0.0
    if cfg.fuse:
        band_rows = min(band_rows, dsms[0].shape[0])
        cube = [dsms[v % n_views][:band_rows].copy() for v in range(cfg.n_views)]
        t0 = time.perf_counter()
        op.fuse_dsms(cube)
        t_fuse = time.perf_counter() - t0
    n_rows = dsms[0].shape[0]
    area = float(shrink * shrink)
    total = (t_views * (cfg.n_views / n_views) + t_fuse * (n_rows / max(band_rows, 1))) * area
    mpix = full.n_views * full.height * full.width / 1e6
    return {'value': mpix / total, 't_views_s': t_views, 't_fuse_s': t_fuse, 'extrapolated_total_s': total,
            'sample': '{} of {} views{} through the per-view stage (Pool({})), fusion on {} of {} grid rows x {} views; '
                      'extrapolated linearly'.format(
                          n_views, cfg.n_views,
                          '' if shrink == 1 else ' on a 1/{0} x 1/{0} sub-AOI ({1}x{2} depth -> {3}x{4} grid)'.format(
                              shrink, cfg.height, cfg.width, cfg.n_size, cfg.e_size),
                          min(cores, n_views), band_rows, n_rows, cfg.n_views)}


def cpu_full_pass(cfg, cores):
    """The reference's CPU path on the WHOLE config, measured (nothing extrapolated): every view through the per-view
    stage with multiprocessing.Pool(cores) exactly like aggregate_2p5d_util.py:138-146 (faithful Python hole-fill
    loop), then the single-process fusion of aggregate_2p5d.py:65-81 over the whole grid."""
    import multiprocessing as mp
    from oracle import geodesy, pipeline as op
    from vissatsatellitestereo_b200 import synthetic as S
    scene = S.make_scene(cfg, geodesy, device='cpu')
    jobs = [(d.numpy(), M, scene.aoi, cfg.res) for d, M in zip(scene.depths, scene.mats)]
    n_proc = min(cores, cfg.n_views)
    t0 = time.perf_counter()
    with mp.get_context('fork').Pool(n_proc) as pool:
        dsms = pool.map(_cpu_view_worker, jobs, chunksize=1)
    t_views = time.perf_counter() - t0
    t_fuse = 0.0
    if cfg.fuse:
        t0 = time.perf_counter()
        op.fuse_dsms(dsms)
        t_fuse = time.perf_counter() - t0
    total = t_views + t_fuse
    mpix = cfg.n_views * cfg.height * cfg.width / 1e6
    return {'value': mpix / total, 't_views_s': t_views, 't_fuse_s': t_fuse, 'total_s': total, 'cores': n_proc,
            'extrapolated': False,
            'sample': 'the whole config, measured once: all {} views through the per-view stage (Pool({})), fusion of '
                      'all {} rows x {} views in one process; nothing extrapolated'.format(
                          cfg.n_views, n_proc, cfg.n_size, cfg.n_views)}


def cpu_pass_estimate_s(cfg, cores):
    """Rough cost model (build-container measurements: 35 s per 2048^2 view with the Python hole-fill loop, 2.6 s per
    256 rows x 2048 cols x 50 views of fusion) used only to decide whether the whole config fits a few minutes."""
    per_view = 35.0 * (cfg.height * cfg.width) / (2048.0 * 2048.0)
    rounds = -(-cfg.n_views // max(1, min(cores, cfg.n_views)))
    fuse = 2.6 * (cfg.n_size * cfg.e_size * cfg.n_views) / (256.0 * 2048.0 * 50.0) if cfg.fuse else 0.0
    return per_view * rounds + fuse


def run_reference_arm(args, cfg):
    """--impl reference: the CPU path of the reference (oracle port) on this box's host cores, all of them
    (Pool(os.cpu_count()) as aggregate_2p5d_util.py:138-141).  ONE pass over the workload is timed regardless of
    --steps/--warmup (a pass takes about a minute and a half for C2; repeating it 25 times would not fit a few
    minutes) and the line says so.  The whole config is run and measured when it fits (C1, C2 at N = 1); otherwise
    BASELINE.md section 3 is followed literally (>= 2 x cpu_count views, >= 1/8 of the fusion rows, linear
    extrapolation, labelled)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from vissatsatellitestereo_b200 import synthetic as S
    cores = os.cpu_count() or 1
    job = S.SynthConfig(**cfg.__dict__)
    job.n_views = cfg.n_views * max(1, args.gpus)          # the b200 arm's workload at this N (weak scaling)
    budget_s = float(os.environ.get('VISSAT_CPU_ARM_BUDGET_S', '330'))
    est = cpu_pass_estimate_s(job, cores)
    if est <= budget_s:
        r = cpu_full_pass(job, cores)
    else:
        n_views = min(job.n_views, 2 * cores)
        r = cpu_sample(job, n_views, max(1, job.n_size // 8), cores)
        r['extrapolated'] = True
        r['cores'] = min(cores, n_views)
    value = r['value']
    mpix = job.n_views * job.height * job.width / 1e6
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * mpix / value,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(cfg, max(1, args.gpus)),
            'cpu_passes_timed': 1, 'measured_not_extrapolated': not r['extrapolated'],
            'estimated_pass_s': est, 't_views_s': r['t_views_s'], 't_fuse_s': r['t_fuse_s'],
            'note': 'one CPU pass over the workload is timed whatever --steps/--warmup say (a pass takes minutes); '
                    'ms_per_step is that pass',
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': r['cores'], 'kind': 'port',
                             'sample': r['sample'], 'host_cpu_count': cores},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def workload_config(cfg, n_gpus):
    return {'workload': '{}: {} views x {}x{} float32 depth per GPU -> {}x{} grid @ {} m ({})'.format(
        cfg.name, cfg.n_views, cfg.height, cfg.width, cfg.n_size, cfg.e_size, cfg.res,
        'BASELINE.json configs[1]' if cfg.name == 'C2' else 'BASELINE.json config'),
        'views_per_gpu': cfg.n_views, 'views_total': cfg.n_views * n_gpus,
        'depth_hw': [cfg.height, cfg.width], 'grid_hw': [cfg.n_size, cfg.e_size], 'resolution_m': cfg.res,
        'fuse': cfg.fuse, 'parallelism': 'views sharded over {} GPU(s); fusion by grid-row band'.format(n_gpus),
        'cache': 'inputs: {:.0f} MB of depth per GPU per step (> 126 MB L2 for C2..C5); no explicit flush'.format(
            cfg.n_views * cfg.height * cfg.width * 4 / 1e6)}


# ------------------------------------------------------------------------------------------------ GPU arm
def measure(cfg, total_views, block, world, rank, dev, steps, warmup, sparse, want_e2e, want_graph, label):
    """Time `steps` passes of the hot path over views [block[0], block[1]) of a `total_views`-view job on this rank
    (every rank calls this with its own block; fusion is sharded by grid-row band when world > 1).
    Returns a dict of measurements (identical on every rank where it matters: times are the max over ranks)."""
    import torch
    import torch.distributed as dist
    from vissatsatellitestereo_b200 import engine as E, synthetic as S, distributed as D
    from vissatsatellitestereo_b200.lib import latlon_utm_converter as geo

    job = S.SynthConfig(**cfg.__dict__)
    job.n_views = total_views
    aoi = S.make_aoi(job, geo)
    terrain = S.Terrain(job, device=dev)
    mats, depths = [], []
    for v in range(block[0], block[1]):
        M, _ = S.make_camera(job, v, aoi['alt_min'])
        mats.append(M)
        depths.append(S.make_depth_map(job, v, M, terrain, device=dev))
    del terrain
    torch.cuda.synchronize()
    V = block[1] - block[0]
    view_counts = [hi - lo for lo, hi in D.split_views(total_views, world)] if world > 1 else [V]
    assert view_counts[rank if world > 1 else 0] == V

    eng = E.DsmEngine(aoi, cfg.res, cfg.res, device=dev)
    eng.collect_stats = False
    stack = torch.empty((V, eng.n_size, eng.e_size), dtype=torch.float32, device=dev)
    P = cfg.height * cfg.width
    G = eng.n_size * eng.e_size
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    stage_events, exch_events, ab_events = [], [], []

    xch, xch_kind = (None, 'none')
    if world > 1 and cfg.fuse:
        xch, xch_kind = D.make_exchange(eng, stack, view_counts, prefer=os.environ.get('VISSAT_EXCHANGE', 'peer'), sparse=sparse)
    peer = xch_kind == 'peer-store'
    n_waves = 5 if (world > 1 and not peer) else 1
    wave_edges = [round(i * V / n_waves) for i in range(n_waves + 1)]
    occ = eng.alloc_occupancy(V) if (sparse and world == 1 and cfg.fuse) else None
    # outputs are allocated once: an allocation inside the step would put host time between the recorded events
    rows_mine = eng.n_size if world == 1 else (lambda b: b[1] - b[0])(D.row_bands(eng.n_size, world)[rank])
    mean_buf = torch.empty((eng.n_size, eng.e_size), dtype=torch.float32, device=dev) if world == 1 else None
    fused_buf = torch.empty((rows_mine, eng.e_size), dtype=torch.float32, device=dev)

    def step(record):
        if record:
            ab0, ab1 = ev(), ev()
            ab0.record()
        if peer:
            xch.begin_step()
            eng.views_to_dsm(depths, mats, stack)                        # stage B also writes the peers' row bands
        else:
            if occ is not None:
                occ.zero_()
                eng.set_occupancy(occ, stack, 0)
            for w in range(n_waves):
                a, b = wave_edges[w], wave_edges[w + 1]
                eng.views_to_dsm(depths[a:b], mats[a:b], stack, first=a)     # one library call per wave (stages A + B)
                if xch is not None:
                    xch.send_wave(a, b)                                      # overlaps stages A/B of the next wave
        if record:
            ab1.record()
            ab_events.append((ab0, ab1))
        fe, fused = None, None
        if cfg.fuse:
            if record:
                f0, f1, f2 = ev(), ev(), ev()
                f0.record()
            if world == 1:
                eng.fuse(stack, out=mean_buf, occ=occ)
                if record:
                    f1.record()
                fused = eng.median3x3(mean_buf, out=fused_buf, count_nan=True)
            elif peer:
                if record:
                    x0 = ev()
                    xch.finish_barrier_only()
                    x0.record()
                    exch_events.append((f0, x0))
                fused, _ = xch.fuse_band(out=fused_buf, after_barrier=record)
                if record:
                    f1.record()
            else:
                if record:
                    x0 = ev()
                band_stack, (r0, r1), (h0, h1) = xch.finish()       # waits only for what is still in flight
                if record:
                    x0.record()
                    exch_events.append((f0, x0))
                mean = eng.fuse(band_stack)
                if record:
                    f1.record()
                fused = eng.median3x3(mean, out=fused_buf, row_begin=r0, row_end=r1, in_row0=h0, h_total=eng.n_size,
                                      count_nan=True)
            if record:
                f2.record()
                fe = (f0, f1, f2)
        if record:
            stage_events.append(fe)
        return fused

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The whole step (stages A-C) is captured once as a CUDA graph and the timed region replays it (VISSAT_GRAPH=0:
    # eager launches).  The work is identical; the per-stage breakdown then comes from an eager instrumented pass
    # right after the timed region.  One GPU: one step per graph.  Several GPUs (peer-store exchange): TWO steps per
    # graph, because the exchange alternates between two band stacks per rank; the stream-ordered barrier between
    # stage B and the fusion (a 1-element NCCL all-reduce) is captured with the rest.
    graph, graph_note, steps_per_replay, graph_parity = None, 'eager launches', 1, 0
    if want_graph and os.environ.get('VISSAT_GRAPH', '1') != '0':
        if world == 1:
            try:
                graph = eng.capture_step(depths, mats, stack, fuse=cfg.fuse, occ=occ)
                graph_note = 'CUDA graph replay ({} kernel launches per step captured once)'.format(graph.launches_per_replay)
            except Exception as e:
                graph, graph_note = None, 'eager launches (graph capture failed: {})'.format(str(e)[:120])
                torch.cuda.synchronize()
        elif peer and cfg.fuse:
            ok = 1
            try:
                for _ in range(2):
                    step(False)                      # lazy allocations (library scratch, NCCL communicator) happen here
                barrier()
                n0 = eng.launch_count()
                graph_parity = xch._cur % 2          # the captured pair starts with the OTHER band stack
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    step(False)
                    fused_cap = step(False)
                graph = E.StepGraph(g, fused_cap, eng.launch_count() - n0)
                steps_per_replay = 2
            except Exception as e:
                ok, graph = 0, None
                graph_note = 'eager launches (graph capture failed: {})'.format(str(e)[:120])
                torch.cuda.synchronize()
            t = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)              # all ranks replay, or none does
            if int(t.item()) != 1:
                graph, steps_per_replay = None, 1
                if ok:
                    graph_note = 'eager launches (graph capture failed on another rank)'
            else:
                graph_note = 'CUDA graph replay, 2 steps per graph ({} kernel launches + 2 NCCL barriers captured once)'.format(
                    graph.launches_per_replay)

    def run_steps(k, record):
        """k steps: graph replays (steps_per_replay each) + eager steps for the remainder"""
        out = None
        done = 0
        if graph is not None and steps_per_replay == 2 and k > 0 and xch._cur % 2 != graph_parity:
            # an odd number of eager steps since the capture: the graph's first step would reuse the band stack the last
            # eager step wrote (its owner may still be fusing it).  One eager step restores the alternation.
            out = step(False)
            done = 1
        if graph is not None:
            while done + steps_per_replay <= k:
                out = graph.replay()
                done += steps_per_replay
        while done < k:
            out = step(record and graph is None)
            done += 1
        return out

    run_steps(warmup, False)
    barrier()
    launches0 = eng.launch_count()
    if graph is None:
        eng.set_timing(True)       # CUDA events around stage A / stage B of every view, recorded inside the library
    t0, t1 = ev(), ev()
    barrier()
    t0.record()
    fused = run_steps(steps, True)
    t1.record()
    barrier()
    if graph is not None:
        n_rep = steps // steps_per_replay
        launches = graph.launches_per_replay * n_rep + (eng.launch_count() - launches0)
    else:
        launches = eng.launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / steps
    mpix_step = total_views * P / 1e6
    value = mpix_step / (ms_step * 1e-3)

    # per-stage device times: inside the timed region (eager), or from an eager instrumented pass of the same step
    if graph is not None:
        fused_graph = fused.clone() if fused is not None else None
        eng.set_timing(True)
        for _ in range(max(1, min(steps, 5))):
            fused = step(True)
        torch.cuda.synchronize()
        if fused_graph is not None:
            assert torch.equal(torch.nan_to_num(fused_graph, nan=-1e9), torch.nan_to_num(fused, nan=-1e9)), \
                'graph replay and eager step disagree'
    k1, k2 = eng.get_timing()
    eng.set_timing(False)
    # Kernel-quality numbers (roofline) come from an instrumented single-stream pass of the same step: in the timed
    # region stage A and stage B kernels of different views overlap on 4 internal streams, which stretches every
    # per-kernel event duration.
    eng.set_streams(1)
    eng.set_timing(True)
    for _ in range(max(1, min(steps, 3))):
        step(False)
    torch.cuda.synchronize()
    k1_iso, k2_iso = eng.get_timing()
    eng.set_timing(False)
    eng.set_streams(4)
    fuse_ms = np.array([fe[0].elapsed_time(fe[1]) for fe in stage_events if fe]) if cfg.fuse else np.array([0.0])
    blur_ms = np.array([fe[1].elapsed_time(fe[2]) for fe in stage_events if fe]) if cfg.fuse else np.array([0.0])
    exch_ms = np.array([a.elapsed_time(b) for a, b in exch_events]) if exch_events else np.array([0.0])
    ab_ms = np.array([a.elapsed_time(b) for a, b in ab_events])          # stages A+B of all V views, per step
    inst_total = float(ab_ms.sum() + fuse_ms.sum() + blur_ms.sum())
    occupancy_frac = None
    nvlink = None
    if occ is not None:
        occupancy_frac = float(_bitmap_fraction(occ, V))
    if peer and xch.sparse:
        xch.keep_last_occ = True            # one more step, keeping a copy of the bitmap the fusion consumed
        step(False)
        torch.cuda.synchronize()
        xch.keep_last_occ = False
        last = getattr(xch, '_last_occ', None)
        if last is not None:
            recv = _received_bytes(xch, last, rank, world, view_counts)
            t = torch.tensor([recv], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dense = total_views * 4.0 * G * (world - 1) / world ** 2
            t_ab = float(ab_ms.mean()) * 1e-3
            nvlink = {'bytes_per_gpu_per_dir_sparse_max': float(t.item()), 'bytes_per_gpu_per_dir_dense': dense,
                      'sparse_over_dense': float(t.item()) / dense if dense else None,
                      'gbs_during_stages_ab': float(t.item()) / t_ab / 1e9 if t_ab > 0 else None,
                      'frac_of_770_gbs': float(t.item()) / t_ab / 1e9 / 770.0 if t_ab > 0 else None,
                      'note': 'the exchange is written by the stage-B kernels (peer stores over NVLink), so it is spread over '
                              'stages A+B; bytes are counted from the owners\' occupancy bitmaps (marked tiles of remote views)'}
    stages = {'k1_unproject_scatter_ms_per_view': float(k1.mean()), 'k2_grid_finalize_ms_per_view': float(k2.mean()),
              'k1_isolated_ms_per_view': float(k1_iso.mean()), 'k2_isolated_ms_per_view': float(k2_iso.mean()),
              'k3_fuse_ms_per_step' if world == 1 else 'exchange_plus_fuse_ms_per_step': float(fuse_ms.mean()),
              'exchange_ms_per_step': float(exch_ms.mean()),
              'stages_ab_ms_per_step': float(ab_ms.mean()), 'stages_ab_effective_ms_per_view': float(ab_ms.mean() / max(V, 1)),
              'note': 'stage A of view v+1 overlaps stage B of view v on internal streams, so the per-kernel '
                      'durations k1/k2 are measured under concurrency and add up to more than stages_ab',
              'k4_median3x3_ms_per_step': float(blur_ms.mean()),
              'launch': graph_note, 'sparse': bool(sparse), 'tile_occupancy': occupancy_frac,
              'share_of_step': {'k1': float(k1.sum() / inst_total), 'k2': float(k2.sum() / inst_total),
                                'fuse': float(fuse_ms.sum() / inst_total), 'blur': float(blur_ms.sum() / inst_total)}}

    res = {'label': label, 'value': value, 'ms_per_step': ms_step, 'launches': int(launches), 'stages': stages,
           'V': V, 'P': P, 'G': G, 'k1_iso': float(k1_iso.mean()), 'k2_iso': float(k2_iso.mean()),
           'k1': float(k1.mean()), 'k2': float(k2.mean()), 'fuse_ms': float(fuse_ms.mean()), 'blur_ms': float(blur_ms.mean()),
           'ab_ms': float(ab_ms.mean()), 'graph_note': graph_note, 'xch_kind': xch_kind, 'fit': eng.fit, 'nvlink': nvlink,
           'mpix_step': mpix_step, 'rows_mine': rows_mine, 'e2e': None}

    # ---- e2e: pinned host depth maps -> H2D -> kernels -> D2H of per-view DSMs + fused DSM, every step
    if want_e2e:
        host_depths = [d.cpu().pin_memory() for d in depths]
        host_views = torch.empty((V, eng.n_size, eng.e_size), dtype=torch.float32).pin_memory()
        host_fused = torch.empty((rows_mine, eng.e_size), dtype=torch.float32).pin_memory()

        def e2e_step():
            if world == 1:
                eng.process_host(host_depths, mats, host_views, host_fused, stack=stack, fuse=cfg.fuse)
                return
            # every rank: its views host -> device -> per-view DSMs -> host; then exchange, fuse own band, band -> host
            eng.process_host(host_depths, mats, host_views, None, stack=stack, fuse=False)
            if cfg.fuse:
                band, _ = D.fuse_distributed(eng, stack, view_counts)
                host_fused.copy_(band, non_blocking=True)
                torch.cuda.synchronize()

        eng.set_occupancy(None)
        for _ in range(max(1, min(warmup, 2))):
            e2e_step()
        n_e2e = max(1, min(steps, 5))
        barrier()
        w0 = time.perf_counter()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        e1.record()
        barrier()
        wall = (time.perf_counter() - w0) / n_e2e
        dev_ms = e0.elapsed_time(e1) / n_e2e
        t_e2e = max(wall, dev_ms * 1e-3)
        if world > 1:
            t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        res['e2e'] = {'value': mpix_step / t_e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(V * P * 4) * world,
                      'd2h_bytes_per_step': int(V * G * 4) * world + (G * 4 if cfg.fuse else 0), 'ms_per_step': 1e3 * t_e2e,
                      'steps': n_e2e, 'api': 'DsmEngine.process_host (pinned host buffers in/out, 3 streams)' +
                      ('' if world == 1 else ' per rank + NCCL row-band exchange + per-rank fused band to host')}
        if cfg.fuse and world == 1 and fused is not None:
            assert np.array_equal(host_fused.numpy(), fused.cpu().numpy(), equal_nan=True), 'e2e result differs'
        # secondary: the same call when only the fused DSM is wanted on the host (no per-view DSMs back): the device->host
        # direction is what the hosts of the GPU boxes are slowest at (tools/pcie_probe.py, profiles/r2b_pcie_probe_n8.txt)
        if cfg.fuse and world == 1:
            def e2e_fused_only():
                eng.process_host(host_depths, mats, None, host_fused, stack=stack, fuse=True)
            e2e_fused_only()
            w0 = time.perf_counter()
            for _ in range(n_e2e):
                e2e_fused_only()
            t_fo = (time.perf_counter() - w0) / n_e2e
            res['e2e']['fused_only'] = {'value': mpix_step / t_fo, 'unit': UNIT, 'ms_per_step': 1e3 * t_fo,
                                        'h2d_bytes_per_step': int(V * P * 4), 'd2h_bytes_per_step': int(G * 4),
                                        'note': 'same API without views_out_host: per-view DSMs stay on the device'}
    if peer:
        xch.close()
    eng.close()
    del stack, depths
    torch.cuda.empty_cache()
    return res


def _bitmap_fraction(occ, V):
    import torch
    g = torch.arange(V, device=occ.device)
    bits = (occ[:, :, g // 32] >> (g % 32).to(torch.int32)) & 1
    return bits.float().mean().item()


def _received_bytes(xch, occ, rank, world, view_counts):
    """Bytes this rank's band stack received from OTHER ranks in one step, from its occupancy bitmap: for every marked
    (tile, view) with a remote view, the rows of the tile inside this rank's band (+ halo) x 64 columns x 4 bytes."""
    import torch
    from vissatsatellitestereo_b200 import _native
    Ty, Tx, _ = occ.shape
    V = sum(view_counts)
    first = sum(view_counts[:rank])
    g = torch.arange(V, device=occ.device)
    remote = (g < first) | (g >= first + view_counts[rank])
    bits = ((occ[:, :, g // 32] >> (g % 32).to(torch.int32)) & 1).bool() & remote[None, None, :]
    ty = torch.arange(Ty, device=occ.device)
    rows = (torch.clamp((ty + 1) * _native.VS_TILE_H, max=xch.h1) - torch.clamp(ty * _native.VS_TILE_H, min=xch.h0)).clamp(min=0)
    tx = torch.arange(Tx, device=occ.device)
    cols = (torch.clamp((tx + 1) * _native.VS_TILE_W, max=xch.W) - tx * _native.VS_TILE_W).clamp(min=0)
    cells = rows[:, None] * cols[None, :]
    return float((bits.sum(dim=2) * cells).sum().item()) * 4.0


def run_b200_arm(args, cfg):
    import torch
    import torch.distributed as dist
    from vissatsatellitestereo_b200 import engine as E, synthetic as S, distributed as D

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    E.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # run on the CPUs of this GPU's NUMA node: the pinned staging buffers of the e2e leg are then allocated there
    # (several ranks stream 1.7 GB per step over PCIe at once).  Undone before the CPU baseline is timed.
    from vissatsatellitestereo_b200 import hostbind
    affinity0 = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
    numa_node = hostbind.bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus, 'launch with torchrun --nproc-per-node {} (WORLD_SIZE={})'.format(args.gpus, world)

    # ---- headline: weak scaling, rank r holds views [r*V, (r+1)*V) of a (V * world)-view job over the same grid
    V = cfg.n_views
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    head = measure(cfg, V * world, (rank * V, (rank + 1) * V), world, rank, dev, args.steps, args.warmup,
                   sparse=cfg.random_offsets, want_e2e=not args.no_e2e, want_graph=True, label=cfg.name)
    clocks = sampler.stop() if rank == 0 else None

    # ---- second record: BASELINE.json configs[2] (C3, the config north_star names for the all-to-all), STRONG scaling:
    # its 200 views are split over the ranks, fusion by row band with the sparse exchange.  Few steps (a step is ~0.1 s).
    c3 = None
    if not args.no_c3 and cfg.name == 'C2':
        c3cfg = S.SynthConfig(**S.CONFIGS['C3'].__dict__)
        blk = D.split_views(c3cfg.n_views, world)[rank]
        r3 = measure(c3cfg, c3cfg.n_views, blk, world, rank, dev, steps=3, warmup=1, sparse=True, want_e2e=False,
                     want_graph=True, label='C3')
        peak3, _ = load_peaks()
        b_alg3 = c3cfg.n_views * 4.0 * r3['P'] + 2 * c3cfg.n_views * 4.0 * r3['G'] + 4.0 * r3['G']
        c3 = {'workload': 'C3 (BASELINE.json configs[2]): 200 views x 4096x4096 depth -> 8192x8192 grid @ 0.3 m, the 200 views '
                          'split over {} GPU(s), fusion by grid-row band'.format(world),
              'scaling': 'strong', 'n_gpus': world, 'value': r3['value'], 'unit': UNIT, 'ms_per_step': r3['ms_per_step'],
              'steps': 3, 'warmup': 1, 'algorithmic_bytes_per_step': b_alg3,
              'achieved_gbs': b_alg3 / (r3['ms_per_step'] * 1e-3) / 1e9,
              'frac_of_hbm_peak': b_alg3 / (r3['ms_per_step'] * 1e-3) / 1e9 / (peak3 * world),
              'stages': r3['stages'], 'exchange': r3['xch_kind'], 'nvlink': r3['nvlink'], 'fit': r3['fit'],
              'gpu_launches': r3['launches']}

    # ---- N > 1: the N-rank result must equal the single-rank result bit for bit (both transports, small grids)
    mgpu = None
    if world > 1:
        from vissatsatellitestereo_b200 import mgpu_selfcheck
        mgpu = mgpu_selfcheck.run(dev, rank, world)

    if affinity0 is not None:
        os.sched_setaffinity(0, affinity0)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if mgpu is not None and not mgpu['bit_identical']:
            sys.exit(3)
        return

    peak, peak_src = load_peaks()
    P, G = head['P'], head['G']
    # dominant kernel and its roofline (algorithmic bytes per launch / measured launch duration)
    per_step = {'k1': head['k1_iso'] * V, 'k2': head['k2_iso'] * V, 'fuse': head['fuse_ms'], 'blur': head['blur_ms']}
    alg = {'k1': (4.0 * P, head['k1_iso'], 'k_unproject_scatter: 4 B/pixel depth read'),
           'k2': (8.0 * G, head['k2_iso'], 'k_grid_finalize: 4 B/cell key read + 4 B/cell DSM write'),
           'fuse': (4.0 * G * V + 4.0 * G / world, head['fuse_ms'], 'k_fuse: 4 B/cell/view read + 4 B/cell write'),
           'blur': (8.0 * G / world, head['blur_ms'], 'k_median3x3: 4 B read + 4 B write per cell')}
    top = max(per_step, key=per_step.get)
    if world > 1 and top in ('fuse', 'blur'):
        top = 'k1' if per_step['k1'] >= per_step['k2'] else 'k2'      # exchange time is not a kernel roofline
    bytes_launch, ms_launch, what = alg[top]
    achieved = bytes_launch / (ms_launch * 1e-3) / 1e9
    # stages A+B of one view with SURVEY 8(d) bytes: depth read + per-view DSM write (key-grid traffic is overhead)
    pair_bytes = 4.0 * P + 4.0 * G
    pair_gbs = pair_bytes / (head['ab_ms'] / V * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': what, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'stages_ab_pair': {'algorithmic_bytes_per_view': pair_bytes, 'effective_ms_per_view': head['ab_ms'] / V,
                                   'achieved': pair_gbs, 'frac': pair_gbs / peak,
                                   'what': 'K1 + K2 of one view with SURVEY 8(d) bytes (4 B/pixel depth read + 4 B/cell per-view '
                                           'DSM write; the key grid between them is overhead, not credit) over the effective '
                                           'per-view time of the overlapped pipeline'},
                'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': bytes_launch, 'avg_launch_ms': float(ms_launch),
                'avg_launch_ms_in_timed_region': float({'k1': head['k1'], 'k2': head['k2'], 'fuse': head['fuse_ms'],
                                                        'blur': head['blur_ms']}[top]),
                'timing': 'CUDA events recorded inside the library around every launch; avg_launch_ms is from an '
                          'instrumented single-stream pass right after the timed region (in the timed region stage A '
                          'and stage B kernels of different views run concurrently on 4 streams, which stretches '
                          'per-kernel durations: avg_launch_ms_in_timed_region)',
                'all_kernels_gbs': {k: float(alg[k][0] / (alg[k][1] * 1e-3) / 1e9) for k in alg if alg[k][1] > 0}}
    traffic_file = os.path.join(REPO, 'profiles', 'traffic.json')
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fp:
                tj = json.load(fp)
            roofline['traffic'] = tj.get(top)
            roofline['traffic_source'] = tj.get('_source', 'profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum '
                                                'per launch from a committed `ncu --set full` capture (not measured in this run)')
        except Exception:
            pass
    # whole-pipeline algorithmic bytes (SURVEY.md 8d B_alg) for context
    Vt = V * world
    b_alg = Vt * 4.0 * P + Vt * 4.0 * G + (Vt * 4.0 * G + 4.0 * G if cfg.fuse else 0.0)
    pipeline_gbs = b_alg / (head['ms_per_step'] * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        # bounded sample: one Pool round of views (one per core) and 1/8 of the fusion rows, extrapolated linearly
        r = cpu_sample(cfg, max(1, min(cores, cfg.n_views)), max(1, cfg.n_size // 8), cores)
        cpu = {'value': r['value'], 'unit': UNIT, 'cores': min(cores, cfg.n_views), 'kind': 'port',
               'sample': r['sample'], 'host_cpu_count': cores}

    e2e = head['e2e']
    if e2e is not None:
        e2e['host_numa_node'] = numa_node
    line = {'metric': METRIC, 'value': head['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(cfg, world),
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(head['launches']), 'roofline': roofline,
            'cpu_baseline': cpu, 'stages': head['stages'],
            'pipeline': {'algorithmic_bytes_per_step': b_alg, 'achieved_gbs': pipeline_gbs,
                         'frac_of_hbm_peak': pipeline_gbs / (peak * world)},
            'fit': head['fit'], 'c3': c3}
    line['config']['launch'] = head['graph_note']
    if world > 1:
        line['config']['exchange'] = {'peer-store': 'stage-B kernel stores row bands into the peers\' stacks '
                                                    '(CUDA IPC peer memory over NVLink) + 1-element all-reduce barrier',
                                      'nccl-waves': 'NCCL grouped send/recv in 5 waves overlapped with stages A/B',
                                      'none': 'none (no fusion)'}[head['xch_kind']]
        line['mgpu_bit_identical'] = bool(mgpu['bit_identical'])
        line['mgpu_check'] = mgpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
        if not mgpu['bit_identical']:
            sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C2')
    ap.add_argument('--views', type=int, default=None, help='override views per GPU (debugging)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer end-to-end leg (kernel experiments)')
    ap.add_argument('--no-c3', action='store_true', help='skip the C3 (strong-scaling, sparse exchange) sub-record')
    args = ap.parse_args()
    from vissatsatellitestereo_b200 import synthetic as S
    cfg = S.SynthConfig(**S.CONFIGS[args.config].__dict__)
    if args.views:
        cfg.n_views = args.views
    if args.impl == 'reference':
        run_reference_arm(args, cfg)
    else:
        run_b200_arm(args, cfg)


if __name__ == '__main__':
    main()
