"""PCIe ceiling for the e2e number: pinned H2D, D2H and both at once (16 MB chunks like the per-view transfers).
Alone: one GPU.  Under torchrun (python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py): all ranks copy at
the same time, which gives the HOST-side ceiling the multi-GPU e2e number runs into (rank 0 prints the aggregate)."""
import os
import time

import torch

rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('gloo')
n, chunk = 50, 2048 * 2048
h_in = [torch.empty(chunk, dtype=torch.float32).pin_memory() for _ in range(n)]
h_out = [torch.empty(chunk, dtype=torch.float32).pin_memory() for _ in range(n)]
d = [torch.empty(chunk, dtype=torch.float32, device='cuda') for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = time.perf_counter()
    for i in range(n):
        if h2d:
            with torch.cuda.stream(s1):
                d[i % 2].copy_(h_in[i], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[i].copy_(d[2 + i % 2], non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter() - t
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    return t


for name, a, b in (('H2D', 1, 0), ('D2H', 0, 1), ('both', 1, 1)):
    run(a, b)
    t = min(run(a, b) for _ in range(3))
    gb = n * chunk * 4 / 1e9
    if rank == 0:
        print('{:5s} {} GPU(s): {:.2f} ms  {:.1f} GB/s per direction per GPU, {:.1f} GB/s aggregate ({} direction(s))'.format(
            name, world, t * 1e3, gb / t, gb / t * world * (a + b), a + b), flush=True)
if world > 1:
    dist.destroy_process_group()
