"""UNet forward (batch 2, 512 config) replayed as its CUDA graph back to back, without any sampler glue: the gap to
bench.py's ms_per_step is what the eager part of a DDIM step costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dynamicrafter_b200.configs import SAMPLER
dev = torch.device("cuda")
host, mp = bench.build_host("512", dev, with_vae=False)
host.unet.use_cuda_graph = True
inp = bench.make_inputs(mp, 1, 16, dev)
x = torch.cat([inp["x_T"], inp["concat"]], 1)
x2 = torch.cat([x, x], 0)
ctx = torch.cat([inp["cond"], inp["uncond"]], 0)
t = torch.tensor([601, 601], device=dev)
fs = torch.tensor([24, 24], device=dev)
for _ in range(3):
    y = host.unet(x2, t, context=ctx, fs=fs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = host.unet(x2, t, context=ctx, fs=fs)
e1.record()
torch.cuda.synchronize()
print(f"UNet forward (B=2) via graph incl. input copies: {e0.elapsed_time(e1) / 10:.3f} ms")
ent = list(host.unet.__dict__["_graphs"].values())[0]
e0.record()
for _ in range(10):
    ent[0].replay()
e1.record()
torch.cuda.synchronize()
print(f"bare graph replay: {e0.elapsed_time(e1) / 10:.3f} ms")
