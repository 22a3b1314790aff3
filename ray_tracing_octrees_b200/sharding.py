"""Multi-GPU host logic of the path: the scene is replicated on every rank, work shards with no data-path collective
(pixels and frames are independent), and the only exchange is the framebuffer gather to rank 0.

Works with any torch.distributed backend: NCCL on the B200 box (device tensors over NVLink), gloo in the CPU tests.
"""
import numpy as np


def shard_frames(num_frames, world, rank):
    """Frames of a camera batch dealt round-robin (frame k -> rank k % world): neighbouring orbit angles cost about the same,
    so round-robin balances better than contiguous blocks.  Returns the list of global frame indices of `rank`."""
    return list(range(rank, num_frames, world))


def shard_rows(height, world, rank, band=8):
    """Interleaved scanline bands of ONE frame: band b (rows [b*band, (b+1)*band)) -> rank b % world.  Sky rows are cheap and
    skyline rows expensive, so interleaving keeps ranks balanced.  Returns [(y0, y1), ...] for `rank` (possibly empty)."""
    out = []
    for b, y0 in enumerate(range(0, height, band)):
        if b % world == rank:
            out.append((y0, min(height, y0 + band)))
    return out


def assemble_rows(height, width, world, parts, band=8, channels=None):
    """Inverse of shard_rows: parts[r] is rank r's planes concatenated in band order -> the full (height, width[, C]) plane."""
    first = next(p for p in parts if p is not None and len(p))
    shape = (height, width) if channels is None else (height, width, channels)
    full = np.empty(shape, dtype=np.asarray(first).dtype)
    for r in range(world):
        off = 0
        for (y0, y1) in shard_rows(height, world, r, band):
            n = y1 - y0
            full[y0:y1] = np.asarray(parts[r])[off:off + n]
            off += n
    return full


def gather_planes(plane, dst=0, group=None):
    """Framebuffer gather: every rank contributes a tensor of identical shape, rank `dst` gets the list (others None).
    One grouped NCCL/gloo gather; the caller may run it on a side stream to overlap the next batch."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return [plane]
    if dist.get_backend(group) == "nccl":
        outs = [torch.empty_like(plane) for _ in range(world)] if rank == dst else None
        dist.gather(plane, outs, dst=dst, group=group)
        return outs
    outs = [torch.empty_like(plane) for _ in range(world)]     # gloo: all_gather is the portable primitive
    dist.all_gather(outs, plane, group=group)
    return outs if rank == dst else None


def interleave_frames(per_rank, num_frames, world):
    """Inverse of shard_frames: per_rank[r][i] is rank r's i-th frame -> frames in global order."""
    out = [None] * num_frames
    for r in range(world):
        for i, k in enumerate(shard_frames(num_frames, world, r)):
            out[k] = per_rank[r][i]
    return out
