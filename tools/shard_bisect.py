"""Frame-sharded tiny UNet vs single GPU (bf16 mode), one line per run: used to bisect sharded-mode problems with the
environment switches (DCB_SHARD_V1, DCB_FUSED_GN_STATS, DCB_TEMPORAL_TC, ...).  torchrun --nproc-per-node N tools/shard_bisect.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if os.environ.get("POISON", "1") == "1":
    # fill the caching allocator with NaN patterns: any kernel that reads memory it (or its producer) never wrote shows up
    junk = [torch.full((256 << 20,), float("nan"), device=dev) for _ in range(8)]         # 8 GB
    junk += [torch.full((1 << 20,), float("nan"), device=dev) for _ in range(256)]
    torch.cuda.synchronize()
    del junk
res = bench.run_shard_check(dev, rank, world)
if rank == 0:
    print("SHARD", os.environ.get("TAG", ""), world, res["bf16"], res["fp32"], flush=True)
dist.destroy_process_group()
