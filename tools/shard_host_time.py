"""Frame-sharded 576x1024 forward: host enqueue time (Python + launch calls, no synchronisation) against the device time of
the same forwards -- is the eager sharded mode launch bound?  torchrun --nproc-per-node N tools/shard_host_time.py
BATCH=1 at N ranks is the per-rank work of the 2N-GPU frames x N x 2 guidance-branch mode.  GRAPH=1 also times the forward
replayed as a CUDA graph that holds the NCCL kernels (leaves with os._exit: process-group teardown hangs with such graphs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from dynamicrafter_b200.parallel import FrameShard
from dynamicrafter_b200.synth import synth_input

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
host, mp = bench.build_host("1024", dev, with_vae=False)
unet = host.unet
unet.use_cuda_graph = False
unet.frame_shard = FrameShard()
B, T, h, w = int(os.environ.get("BATCH", "2")), 16, 72, 128
x = synth_input("x", (B, 8, T, h, w), 3).to(dev)
ctx = synth_input("c", (B, 77 + 16 * T, 1024), 3).to(dev)
ts = torch.full((B,), 500, dtype=torch.long, device=dev)
fs = torch.full((B,), 10, dtype=torch.long, device=dev)
for _ in range(2):
    unet(x, ts, context=ctx, fs=fs)
torch.cuda.synchronize(); dist.barrier()
n = 4
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
host_s = 0.0
for _ in range(n):
    a = time.perf_counter()
    unet(x, ts, context=ctx, fs=fs)
    host_s += time.perf_counter() - a
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
if rank == 0:
    print(f"SHARDHOST world={world} batch={B}: host enqueue {host_s / n * 1e3:.1f} ms / forward, device {e0.elapsed_time(e1) / n:.1f} ms / forward, "
          f"wall {wall / n * 1e3:.1f} ms", flush=True)
if os.environ.get("GRAPH", "0") == "1":
    from dynamicrafter_b200.lvdm import _packing
    ref = unet(x, ts, context=ctx, fs=fs).clone()
    _packing.GRAPH_COLLECTIVES = True
    unet.use_cuda_graph = True
    for _ in range(2):
        out = unet(x, ts, context=ctx, fs=fs)
    torch.cuda.synchronize(); dist.barrier()
    diff = (out.float() - ref.float()).abs().max().item()
    n = 8
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        unet(x, ts, context=ctx, fs=fs)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if rank == 0:
        print(f"SHARDHOST world={world} batch={B} graph: device {e0.elapsed_time(e1) / n:.1f} ms / forward, wall {wall / n * 1e3:.1f} ms, "
              f"max |graph - eager| {diff:.3e}", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)
dist.destroy_process_group()
