"""world_size-2 gloo test (CPU) of the N>1 plumbing: video sharding + the single all-gather of final latents."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynamicrafter_b200.parallel import shard_range, all_gather_latents


def test_shard_range_covers_everything():
    for n in (1, 2, 7, 8, 9, 16):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_videos, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts = [shard_range(n_videos, r, world)[1] - shard_range(n_videos, r, world)[0] for r in range(world)]
    lo, hi = shard_range(n_videos, rank, world)
    full = torch.arange(n_videos * 4 * 2 * 3 * 3, dtype=torch.float32).reshape(n_videos, 4, 2, 3, 3)
    got = all_gather_latents(full[lo:hi].clone(), counts)
    ok = torch.equal(got, full)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(float(t))
    dist.destroy_process_group()


def test_all_gather_latents_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    for n_videos in (2, 3):
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_videos, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert q.get(timeout=10) == 1.0


def _transpose_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamicrafter_b200.parallel import FrameShard, frames_to_pixels, pixels_to_frames, all_gather_frames
    shard = FrameShard()
    B, T, HW, C = 2, 4, 6, 3
    full = torch.arange(B * T * HW * C, dtype=torch.float32).reshape(B, T, HW, C)
    t0, t1 = shard.frames(T)
    mine = full[:, t0:t1].reshape(-1, C).clone()                       # rows (b, t_loc, p)
    px = frames_to_pixels(mine, B, t1 - t0, HW, shard)                 # rows (b, t, p_loc)
    p0, p1 = rank * HW // world, (rank + 1) * HW // world
    ok = torch.equal(px, full[:, :, p0:p1].reshape(-1, C))
    back = pixels_to_frames(px, B, T, HW // world, shard)
    ok = ok and torch.equal(back, mine)
    nchw = full[:, t0:t1].permute(0, 3, 1, 2).reshape(B, C, t1 - t0, 2, 3).contiguous()
    ok = ok and torch.equal(all_gather_frames(nchw, shard), full.permute(0, 3, 1, 2).reshape(B, C, T, 2, 3))
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(float(t))
    dist.destroy_process_group()


def test_frame_pixel_transposes_world2():
    """frame-sharded mode: frames <-> pixels all-to-all is an exact transpose and its own inverse."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_transpose_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0
