"""Single-kernel driver for ncu: text + image cross-attention at the 40x64 level (32 frames x 5 heads x 2560 queries)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
dev = torch.device("cuda")
torch.manual_seed(0)
B, T, H, N = 2, 16, 5, 2560
inner, nb = H * 64, B * T
q = torch.randn(nb * N, inner, device=dev).to(torch.bfloat16)
kv_t = torch.randn(B * 77, 2 * inner, device=dev).to(torch.bfloat16)
kv_i = torch.randn(nb * 16, 2 * inner, device=dev).to(torch.bfloat16)
for _ in range(4):
    out = ops.flash_attention(q, [(kv_t[:, :inner], kv_t[:, inner:], 77, T, 1.0), (kv_i[:, :inner], kv_i[:, inner:], 16, 1, 1.0)],
                              nb, H, N, 0.125)
torch.cuda.synchronize()
print("ok")
