"""Single-shape driver for ncu: python tools/prof_gemm.py M N K [f32] [res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
dev = torch.device("cuda")
M, N, K = (int(v) for v in sys.argv[1:4])
f32, res = "f32" in sys.argv, "res" in sys.argv
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
r = torch.randn(M, N, device=dev) if res else None
for _ in range(4):
    ops.linear(a, w, residual=r, out=out)
torch.cuda.synchronize()
print("ok")
