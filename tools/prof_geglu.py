"""Single-kernel driver for ncu: a few launches of one tap_gemm shape (default: the level-0 GEGLU GEMM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
dev = torch.device("cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else "geglu"
M, N, K = (81920, 2560, 320) if mode == "geglu" else (81920, 960, 320)
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
out = torch.empty(M, N // 2 if mode == "geglu" else N, dtype=torch.bfloat16, device=dev)
for _ in range(4):
    ops.linear(a, w, bias=bias, geglu=(mode == "geglu"), out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
