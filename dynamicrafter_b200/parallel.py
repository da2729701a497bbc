"""Multi-GPU plumbing of the denoising path (one process per GPU, torch.distributed).

The path shards over independent videos and over the cond / uncond CFG pair -- nothing is exchanged inside a
denoising step; the reference itself only runs independent ranks over disjoint prompts
(scripts/evaluation/inference.py:350-356).  The single collective is one all-gather of the final latents, so every
rank can decode any subset of the frames.
"""
import torch
import torch.distributed as dist


class LinkTraffic:
    """Bytes this rank hands to its peers (what crosses NVLink), counted on the host from the tensors given to each
    collective of this module: all-to-all = the off-diagonal blocks, all-gather = one copy per peer, all-reduce = the
    ring figure 2 (N-1)/N.  `bench.py` reports it per forward next to the algorithmic figure of SURVEY.md 8(e).  Counts
    calls as they are ENQUEUED (a replayed CUDA graph does not pass through here)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.bytes, self.calls = 0, 0

    def add(self, tensor_bytes, world, kind):
        if world <= 1:
            return
        if kind == "all_to_all":
            self.bytes += tensor_bytes * (world - 1) // world
        elif kind == "all_gather":
            self.bytes += tensor_bytes * (world - 1)
        else:
            self.bytes += 2 * tensor_bytes * (world - 1) // world
        self.calls += 1


LINK = LinkTraffic()


def _nbytes(t):
    return t.numel() * t.element_size()


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


# ------------------------------------------------------------------------------------------------------------
# Frame-sharded single video (576x1024 latency mode).  Per-frame work (ResBlock 2-D part,
# SpatialTransformer, up / down sampling) runs on the rank's own frames; the temporal blocks (TemporalConvBlock,
# TemporalTransformer: attention.py:365-412, openaimodel3d.py:272-279) need every frame of a pixel, so the residual
# stream is transposed frames <-> pixels with one all-to-all on each side of them, and the 5-D GroupNorms of those
# blocks all-reduce their 64 partial sums per video.
# ------------------------------------------------------------------------------------------------------------
class FrameShard:
    """Which frames / pixels of the video this rank owns.  T and every level's H*W must divide by the world size."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def frames(self, T):
        if T % self.world:
            raise ValueError(f"frame sharding needs T={T} divisible by the {self.world} ranks")
        n = T // self.world
        return self.rank * n, (self.rank + 1) * n


def frames_to_pixels(x, B, T_loc, HW, shard):
    """x: [B*T_loc*HW, C] rows (b, t_loc, p) of this rank's frames -> [B*T*HW_loc, C] rows (b, t, p_loc): every frame
    of this rank's pixel block.  One all-to-all of the whole tensor minus the diagonal block."""
    N, C = shard.world, x.shape[1]
    if HW % N:
        raise ValueError(f"frame sharding needs H*W={HW} divisible by the {N} ranks")
    send = x.reshape(B, T_loc, N, HW // N, C).permute(2, 0, 1, 3, 4).contiguous()        # [dst, b, t_loc, p, c]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=shard.group)                                 # [src, b, t_loc, p, c]
    LINK.add(_nbytes(send), N, "all_to_all")
    return recv.permute(1, 0, 2, 3, 4).reshape(B * N * T_loc * (HW // N), C)              # (b, (src t_loc), p_loc)


def pixels_to_frames(x, B, T, HW_loc, shard):
    """Inverse of frames_to_pixels: [B*T*HW_loc, C] rows (b, t, p_loc) -> [B*T_loc*HW, C] rows (b, t_loc, p)."""
    N, C = shard.world, x.shape[1]
    T_loc = T // N
    send = x.reshape(B, N, T_loc, HW_loc, C).permute(1, 0, 2, 3, 4).contiguous()          # [dst, b, t_loc, p_loc, c]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=shard.group)                                 # [src = pixel block, ...]
    LINK.add(_nbytes(send), N, "all_to_all")
    return recv.permute(1, 2, 0, 3, 4).reshape(B * T_loc * N * HW_loc, C)                 # (b, t_loc, (src p_loc))


# Version 2 of the transposes (round 2): the temporal blocks are split into "norm" (run on the frame side, statistics
# all-reduced) and "core" (run on the pixel side, producing the block's DELTA without the residual), so what crosses
# NVLink is the bf16 normalised operand one way and the bf16 delta the other way -- half the bytes of the fp32
# residual stream -- and the fp32 residual add happens on the frame side inside the unpack kernel.  One all-to-all per
# video: its receive buffer [src, t_loc, p_loc, c] IS the pixel-side row order (t, p_loc), and the pixel-side result
# [dst, t_loc, p_loc, c] IS the send buffer of the way back, so only the frame side re-orders rows
# (ops.block_transpose: one pack, one unpack fused with the residual add).
def frames_to_pixels_v2(y, B, T_loc, HW, shard):
    """y: [B*T_loc*HW, C] rows (b, t_loc, p) (bf16 operand; fp32 in the fp32 parity mode) -> [B*T*HW_loc, C] rows (b, t, p_loc)."""
    from . import ops
    N, C = shard.world, y.shape[1]
    if HW % N:
        raise ValueError(f"frame sharding needs H*W={HW} divisible by the {N} ranks")
    send = ops.block_transpose(y, B, T_loc, N)                       # [b][dst][t_loc][p_loc, c]
    recv = torch.empty_like(send)
    per = send.numel() // B
    sv, rv = send.view(-1), recv.view(-1)
    for b in range(B):
        dist.all_to_all_single(rv[b * per:(b + 1) * per], sv[b * per:(b + 1) * per], group=shard.group)
        LINK.add(per * send.element_size(), N, "all_to_all")
    return recv.view(B * N * T_loc * (HW // N), C)                   # [b][src][t_loc][p_loc] = rows (b, t, p_loc)


def pixels_to_frames_add(delta, x, B, T, HW_loc, shard):
    """delta: [B*T*HW_loc, C] rows (b, t, p_loc) of the block output WITHOUT its residual; x: fp32 [B*T_loc*HW, C] residual
    stream on the frame side -> x + delta in frame order (fp32)."""
    from . import ops
    N = shard.world
    T_loc = T // N
    delta = delta.contiguous()
    recv = torch.empty_like(delta)
    per = delta.numel() // B
    sv, rv = delta.view(-1), recv.view(-1)
    for b in range(B):
        dist.all_to_all_single(rv[b * per:(b + 1) * per], sv[b * per:(b + 1) * per], group=shard.group)
        LINK.add(per * delta.element_size(), N, "all_to_all")
    if delta.dtype == torch.bfloat16:
        return ops.block_transpose(recv, B, N, T_loc, residual=x)    # [b][src][t_loc][..] -> [b][t_loc][src][..] + x
    return x + ops.block_transpose(recv, B, N, T_loc).view_as(x)      # fp32 parity mode


def all_gather_frames(local, shard):
    """local: [B, C, T_loc, H, W] (this rank's frames) -> [B, C, T, H, W] on every rank."""
    bufs = [torch.empty_like(local) for _ in range(shard.world)]
    dist.all_gather(bufs, local.contiguous(), group=shard.group)
    LINK.add(_nbytes(local), shard.world, "all_gather")
    return torch.cat(bufs, dim=2)


def all_reduce_sums(sums, shard):
    """In-place SUM of the GroupNorm partial (sum, sum of squares) over the ranks holding the instance."""
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=shard.group)
    LINK.add(_nbytes(sums), shard.world, "all_reduce")
    return sums


# ------------------------------------------------------------------------------------------------------------
# Guidance branches over GPUs (option (i)): the cond / uncond (/ image-only) forwards of one
# video share nothing but x_t, so rank r evaluates branch r % n_branches and the [B, 4, T, h, w] predictions (655 KB
# each at the 512 config) are all-gathered every step; every rank then applies the fused update redundantly.
# ------------------------------------------------------------------------------------------------------------
class BranchShard:
    """One guidance branch per rank: the world size must equal the number of branches of the sampler (2, or 3 for the
    multi-condition sampler)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def gather(self, mine):
        """mine: this rank's branch output -> list of all branch outputs in branch order."""
        outs = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(outs, mine.contiguous(), group=self.group)
        LINK.add(_nbytes(mine), self.world, "all_gather")
        return outs
