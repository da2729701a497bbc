"""Single-shape GroupNorm driver for ncu: python tools/prof_gn.py G S C [bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
dev = torch.device("cuda")
G, S, C = (int(v) for v in sys.argv[1:4])
dt = torch.bfloat16 if "bf16" in sys.argv else torch.float32
src = torch.randn(G * S, C, device=dev).to(dt)
x = torch.empty_like(src)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(3):
    x.copy_(src)
    y = ops.group_norm(x, G, S, g, b, 1e-5)
torch.cuda.synchronize()
print("ok")
