import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
dev = torch.device("cuda")
B, H, N = 32, 5, 2560
inner = H * 64
qkv = torch.randn(B * N, 3 * inner, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, inner, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], N, 1, 1.0)], B, H, N, 0.125, out=out)
torch.cuda.synchronize()
print("ok")
