"""Multi-GPU plumbing of the denoising path (one process per GPU, torch.distributed).

The path shards over independent videos and over the cond / uncond CFG pair -- nothing is exchanged inside a
denoising step; the reference itself only runs independent ranks over disjoint prompts
(scripts/evaluation/inference.py:350-356).  The single collective is one all-gather of the final latents, so every
rank can decode any subset of the frames.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced split of `n_items` work units (videos) over `world` ranks -> (start, stop)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_gather_latents(local, counts=None, group=None):
    """local: [B_local, C, T, H, W] on this rank -> [sum(B_r), C, T, H, W], rank order.  `counts[r]` = B_r
    (defaults to equal shards).  Works on NCCL (GPU) and gloo (CPU tests)."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    if counts is None:
        counts = [local.shape[0]] * world
    bmax = max(counts)
    pad = local
    if local.shape[0] < bmax:
        pad = torch.cat([local, local.new_zeros((bmax - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def frame_shard(n_frames, rank, world):
    """Frames of the gathered latents this rank decodes with the VAE (decode is per frame)."""
    return shard_range(n_frames, rank, world)
