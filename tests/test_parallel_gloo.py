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
    # host-side count of what was handed to the peer: two all-to-alls (half of each tensor) + one all-gather (all of it)
    from dynamicrafter_b200.parallel import LINK
    ok = ok and LINK.calls == 3 and LINK.bytes == 2 * (mine.numel() * 4 // 2) + nchw.numel() * 4
    LINK.reset()
    ok = ok and LINK.bytes == 0 and LINK.calls == 0
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


def _branch_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamicrafter_b200.parallel import BranchShard
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler

    class _Model:                                   # apply_model marks which conditioning it was given
        num_timesteps = 1000
        def apply_model(self, x, t, c, **kw):
            return x + c

    smp = DDIMSampler.__new__(DDIMSampler)
    smp.model, smp.batch_cfg = _Model(), True
    smp.branch_shard = BranchShard()
    x = torch.zeros(1, 4, 2, 2, 2)
    outs = smp._eval_branches(x, torch.zeros(1, dtype=torch.long), [torch.tensor([1.0]), torch.tensor([-2.0])], {})
    ok = len(outs) == 2 and float(outs[0].mean()) == 1.0 and float(outs[1].mean()) == -2.0
    # sharded single-video modes: every rank must apply the update with rank 0's draws, whatever its own seed was
    torch.manual_seed(100 + rank)
    noise = smp._sync_noise(torch.randn(1, 4, 2, 2, 2))
    torch.manual_seed(100)
    ok = ok and torch.equal(noise, torch.randn(1, 4, 2, 2, 2))
    try:                                            # three branches on two ranks must be refused
        smp._eval_branches(x, None, [torch.tensor([1.0])] * 3, {})
        ok = False
    except ValueError:
        pass
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(float(t))
    dist.destroy_process_group()


def test_branch_sharded_guidance_world2():
    """one guidance branch per rank: rank r evaluates conditioning r, every rank ends with both predictions."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_branch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0


def test_cat_cond_structures():
    """the guidance branches are batched only when their conditionings line up structurally."""
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    a, b = torch.ones(2, 3), torch.zeros(2, 3)
    assert DDIMSampler._cat_cond(a, b).shape == (4, 3)
    d1 = {"c_crossattn": [a], "c_concat": [torch.ones(2, 1)]}
    d2 = {"c_crossattn": [b], "c_concat": [torch.ones(2, 1)]}
    got = DDIMSampler._cat_cond(d1, d2, d2)
    assert got["c_crossattn"][0].shape == (6, 3) and got["c_concat"][0].shape == (6, 1)
    assert DDIMSampler._cat_cond(a, torch.zeros(2, 4)) is None
    assert DDIMSampler._cat_cond(d1, {"c_crossattn": [b]}) is None
    assert DDIMSampler._cat_cond(a, d1) is None


def test_precision_switch():
    from dynamicrafter_b200 import precision
    assert precision.get_precision() in ("bf16", "fp32")
    before = precision.get_precision()
    with precision.use_precision("fp32"):
        assert precision.fp32_mode()
        with precision.use_precision("bf16"):
            assert not precision.fp32_mode()
        assert precision.fp32_mode()
    assert precision.get_precision() == before
    import pytest
    with pytest.raises(ValueError):
        precision.set_precision("fp16")


def _grid_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamicrafter_b200.parallel import BranchShard, FrameShard
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    # the grid bench.py builds from 4 GPUs on: rank r = guidance branch r % 2 of frame shard r // 2
    fgroup = bgroup = None
    for br in range(2):
        g = dist.new_group([r for r in range(world) if r % 2 == br])
        if rank % 2 == br:
            fgroup = g
    for fi in range(world // 2):
        g = dist.new_group([2 * fi, 2 * fi + 1])
        if rank // 2 == fi:
            bgroup = g

    class _Unet:
        frame_shard = FrameShard(fgroup)

    class _Model:
        num_timesteps = 1000
        unet = _Unet()

    smp = DDIMSampler.__new__(DDIMSampler)
    smp.model = _Model()
    smp.branch_shard = BranchShard(bgroup)
    ok = smp.branch_shard.world == 2 and smp.branch_shard.rank == rank % 2
    ok = ok and _Unet.frame_shard.world == world // 2 and _Unet.frame_shard.rank == rank // 2
    ok = ok and _Unet.frame_shard.frames(16) == (rank // 2 * (32 // world), (rank // 2 + 1) * (32 // world))
    torch.manual_seed(100 + rank)
    noise = smp._sync_noise(torch.randn(1, 4, 2, 2, 2))
    torch.manual_seed(100)
    ok = ok and torch.equal(noise, torch.randn(1, 4, 2, 2, 2))
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(float(t))
    dist.destroy_process_group()


def test_branch_x_frame_grid_world4():
    """4 ranks as 2 guidance branches x 2 frame shards: group membership, frame ranges, and rank 0's random draws reach
    every rank through the two sub-group broadcasts."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grid_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0


def _block_transpose_torch(src, Z, A, B, residual=None):
    """torch statement of ops.block_transpose (the CUDA kernel itself is checked in tests/test_kernels_gpu.py)."""
    out = src.reshape(Z, A, B, -1).permute(0, 2, 1, 3).contiguous().reshape(src.shape)
    if residual is not None:
        return residual + out.float().view_as(residual)
    return out


def _v2_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamicrafter_b200 import ops, parallel
    ops.block_transpose = _block_transpose_torch            # host logic only: no GPU in this test
    shard = parallel.FrameShard()
    B, T, HW, C = 2, 4, 6, 8
    full = torch.arange(B * T * HW * C, dtype=torch.float32).reshape(B, T, HW, C)
    t0, t1 = shard.frames(T)
    p0, p1 = rank * HW // world, (rank + 1) * HW // world
    mine = full[:, t0:t1].reshape(-1, C).clone()                                   # rows (b, t_loc, p)
    ok = True
    for dt in (torch.float32, torch.bfloat16):
        parallel.LINK.reset()
        px = parallel.frames_to_pixels_v2(mine.to(dt), B, t1 - t0, HW, shard)      # rows (b, t, p_loc)
        ok = ok and torch.equal(px.float(), full[:, :, p0:p1].reshape(-1, C).to(dt).float())
        resid = torch.full_like(mine, 0.5)
        back = parallel.pixels_to_frames_add(px, resid, B, T, HW // world, shard)  # the way back + the fp32 residual add
        ok = ok and back.dtype == torch.float32 and torch.equal(back, mine.to(dt).float() + 0.5)
        # one all-to-all per video each way, half of every tensor leaves the rank
        ok = ok and parallel.LINK.calls == 2 * B and parallel.LINK.bytes == 2 * (mine.numel() * px.element_size() // 2)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(float(t))
    dist.destroy_process_group()


def test_frame_pixel_transposes_v2_world2():
    """round-2 flow of the frame-sharded mode: operand to the pixel side, delta back + residual add, one all-to-all per
    video each way (the block transposes replaced by their torch statement)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_v2_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1.0
