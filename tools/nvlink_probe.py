"""NVLink traffic of the stage-B peer stores, for ncu's nvlink counters (run under torchrun, 2+ ranks):
  python -m torch.distributed.run --nproc-per-node 2 --no-python ncu --metrics nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,gpu__time_duration.sum \
      -k regex:k_grid_finalize -c 12 --csv --log-file gpurun_out/nvl_%p.csv python tools/nvlink_probe.py
C2 shapes, a few views per rank, dense exchange (every tile is stored): each stage-B launch must send
4 B x G x (N-1)/N (+ halo rows) over NVLink."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from vissatsatellitestereo_b200 import distributed as D, engine as E, synthetic as S  # noqa: E402
from vissatsatellitestereo_b200.lib import latlon_utm_converter as geo  # noqa: E402


def nvlink_kib(index):
    """Cumulative NVLink payload counters of GPU `index` (NVML field values, KiB, all links): (tx, rx) or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        out = []
        for fid in (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX):
            v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, 0xffffffff)])[0]
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal))
        return tuple(out)
    except Exception as e:
        return None


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    V = int(os.environ.get('VISSAT_PROBE_VIEWS', '4'))
    cfg = S.SynthConfig(**S.CONFIGS['C2'].__dict__)
    cfg.n_views = V * world
    aoi = S.make_aoi(cfg, geo)
    terrain = S.Terrain(cfg, device=dev)
    mats = [S.make_camera(cfg, v, aoi['alt_min'])[0] for v in range(rank * V, (rank + 1) * V)]
    depths = [S.make_depth_map(cfg, rank * V + i, mats[i], terrain, device=dev) for i in range(V)]
    eng = E.DsmEngine(aoi, cfg.res, cfg.res, device=dev)
    eng.collect_stats = False
    stack = torch.empty((V, eng.n_size, eng.e_size), dtype=torch.float32, device=dev)
    xch = D.PeerExchange(eng, stack, [V] * world, sparse=False)
    steps = int(os.environ.get('VISSAT_PROBE_STEPS', '3'))
    G = eng.n_size * eng.e_size
    for _ in range(2):
        xch.begin_step()
        eng.views_to_dsm(depths, mats, stack)
        out, _ = xch.fuse_band()
    torch.cuda.synchronize()
    dist.barrier()
    c0 = nvlink_kib(local_rank)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        xch.begin_step()
        eng.views_to_dsm(depths, mats, stack)
        out, _ = xch.fuse_band()
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    c1 = nvlink_kib(local_rank)
    if c0 is not None and c1 is not None:
        tx, rx = (c1[0] - c0[0]) * 1024.0, (c1[1] - c0[1]) * 1024.0
        # every local view sends the rows of every other rank's band stack (its band + 1 halo row on each inner side), per step
        bands = D.row_bands(eng.n_size, world)
        rows_out = sum((b[1] - b[0]) + (1 if b[0] > 0 else 0) + (1 if b[1] < eng.n_size else 0) for j, b in enumerate(bands) if j != rank)
        expect = steps * V * 4.0 * eng.e_size * rows_out
        ms = t0.elapsed_time(t1)
        print('NVLINK_COUNTERS rank {} of {}: NVML tx {:.1f} MB rx {:.1f} MB over {} steps x {} views ({:.2f} ms); peer-store payload '
              'expected {:.1f} MB per direction -> tx/expected {:.3f}; {:.1f} GB/s tx while the steps ran'.format(
                  rank, world, tx / 1e6, rx / 1e6, steps, V, ms, expect / 1e6, tx / expect if expect else float('nan'),
                  tx / (ms * 1e-3) / 1e9), flush=True)
    else:
        print('NVLINK_COUNTERS rank {}: NVML nvlink throughput counters not available'.format(rank), flush=True)
    if rank == 0:
        print('NVLINK_PROBE world {} views/rank {} grid {}x{}: expected user bytes per stage-B launch and direction >= {:.1f} MB'.format(
            world, V, eng.n_size, eng.e_size, 4.0 * G * (world - 1) / world / 1e6), flush=True)
    xch.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
