"""PCIe ceiling for the e2e number: pinned H2D, D2H and both at once (16 MB chunks like the per-view transfers)."""
import time
import torch
n, chunk = 50, 2048 * 2048
h_in = [torch.empty(chunk, dtype=torch.float32).pin_memory() for _ in range(n)]
h_out = [torch.empty(chunk, dtype=torch.float32).pin_memory() for _ in range(n)]
d = [torch.empty(chunk, dtype=torch.float32, device='cuda') for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(n):
        if h2d:
            with torch.cuda.stream(s1):
                d[i % 2].copy_(h_in[i], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[i].copy_(d[2 + i % 2], non_blocking=True)
    torch.cuda.synchronize()
    return time.perf_counter() - t


for name, a, b in (('H2D', 1, 0), ('D2H', 0, 1), ('both', 1, 1)):
    run(a, b)
    t = min(run(a, b) for _ in range(3))
    gb = n * chunk * 4 / 1e9
    print('{:5s} {:.2f} ms  {:.1f} GB/s per direction'.format(name, t * 1e3, gb / t))
