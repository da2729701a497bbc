"""cProfile of the host side of ONE eager UNet forward (no CUDA graph): where the ~40 us per launch of Python go."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dynamicrafter_b200.synth import synth_input
dev = torch.device("cuda", 0)
host, mp = bench.build_host("512", dev, with_vae=False)
unet = host.unet
unet.use_cuda_graph = False
B, T, h, w = 1, 16, 40, 64
x = synth_input("x", (B, 8, T, h, w), 3).to(dev); ctx = synth_input("c", (B, 77 + 16 * T, 1024), 3).to(dev)
ts = torch.full((B,), 500, dtype=torch.long, device=dev); fs = torch.full((B,), 24, dtype=torch.long, device=dev)
for _ in range(3):
    unet(x, ts, context=ctx, fs=fs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    unet(x, ts, context=ctx, fs=fs)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {(t1 - t0) / 5 * 1e3:.1f} ms / forward, with sync {(t2 - t0) / 5 * 1e3:.1f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    unet(x, ts, context=ctx, fs=fs)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
