"""Multi-GPU driver: one process per GPU (torchrun), NCCL over NVLink for the one exchange the path has.

SURVEY.md §8(e): stages A+B are independent per view -> the sorted view list is split into contiguous blocks,
one per rank (so the fused sum order stays the reference's sorted order, aggregate_2p5d.py:59).  Stage C is
independent per cell -> each rank owns a contiguous band of grid rows (+1 halo row each side for the final
3x3 blur).  The transpose between the two layouts is an all-to-all of per-view DSM row bands; it replaces the
file-system hand-off of aggregate_2p5d.py:57-66.  The result is independent of the number of ranks.

The host logic here runs on any backend (tests use gloo on CPU tensors); the compute calls need CUDA.
"""
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
