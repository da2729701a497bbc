"""Short-K GEMMs of the 40x64 level (CUDA events, 10 launches each).  DCB_GEMM_AS_MIN=0 disables the A-stationary mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
dev = torch.device("cuda")
torch.manual_seed(0)
M = 81920
for name, N, K, kw in [("to_out f32+res", 320, 320, dict(f32=True, bf16=False, res=True)), ("qkv bf16", 960, 320, dict()),
                       ("proj f32", 320, 320, dict(f32=True, bf16=False)), ("geglu", 2560, 320, dict(geglu=True)),
                       ("ff2 bf16+res K=1280", 320, 1280, dict(res=True))]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev) if kw.pop("res", False) else None
    geglu = kw.pop("geglu", False)
    run = lambda: ops.linear(a, w, bias=bias, residual=res, geglu=geglu, block_n=256 if geglu else 0, **kw)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{name:22s} {M}x{N}x{K}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:6.0f} TFLOP/s")
