"""Spatial self-attention micro-benchmark: CUDA events over repeated launches (qkv of the 40x64 level is 157 MB > L2).
python tools/bench_attn.py [B H N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
dev = torch.device("cuda")
shapes = [(32, 5, 2560), (32, 10, 640), (32, 5, 9216)] if len(sys.argv) < 4 else [tuple(int(v) for v in sys.argv[1:4])]
for B, H, N in shapes:
    inner = H * 64
    qkv = torch.randn(B * N, 3 * inner, device=dev).to(torch.bfloat16)
    out = torch.empty(B * N, inner, dtype=torch.bfloat16, device=dev)
    run = lambda: ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], N, 1, 1.0)], B, H, N, 0.125, out=out)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"B={B} H={H} N={N}: {us:8.1f} us  {4.0 * B * H * N * N * 64 / us / 1e6:7.1f} TFLOP/s")
