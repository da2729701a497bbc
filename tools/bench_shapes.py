"""tap_gemm over the step's main linear / conv shapes, CUDA-graph timed (no launch overhead): us and TFLOP/s per shape.
python tools/bench_shapes.py [tag]   (env switches, e.g. DCB_GEMM_RAGGED256=0, select the variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
NREP = 10


def graph_time(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(NREP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / NREP * 1e3


def rnd(*s, dt=torch.bfloat16):
    return torch.randn(*s, device=dev).to(dt)


# (kind, M or (nf,H,W) / (B,T,HW), N, K-per-tap C, f32 out, residual)
LIN = [(81920, 960, 320, False, False), (81920, 320, 320, True, True), (81920, 320, 1280, False, True),
       (20480, 1920, 640, False, False), (20480, 640, 640, True, True), (20480, 640, 2560, False, True),
       (5120, 3840, 1280, False, False), (5120, 1280, 1280, True, True), (5120, 1280, 5120, False, True),
       (1280, 3840, 1280, False, False), (1280, 1280, 1280, True, True), (1280, 1280, 5120, False, True)]
for M, N, K, f32, res in LIN:
    a, w = rnd(M, K), rnd(N, K) * K ** -0.5
    r = rnd(M, N, dt=torch.float32) if res else None
    out = torch.empty(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
    us = graph_time(lambda: ops.linear(a, w, residual=r, out=out))
    print(f"linear  M={M:6d} N={N:5d} K={K:5d} f32={int(f32)} res={int(res)}: {us:7.1f} us {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s")
B, T = 2, 16
for H, W, C in ((40, 64, 320), (20, 32, 640), (10, 16, 1280), (5, 8, 1280)):
    HW, M = H * W, B * T * H * W
    y = rnd(M, C)
    w3, w9 = rnd(C, 3 * C) * (3 * C) ** -0.5, rnd(C, 9 * C) * (9 * C) ** -0.5
    bias, res = torch.randn(C, device=dev), rnd(M, C, dt=torch.float32)
    with ops.stats_scope(dev):
        us = graph_time(lambda: ops.tconv3(y, B, T, HW, w3, bias=bias, f32=False, bf16=True, stats=True))
    print(f"tconv3  M={M:6d} C={C:5d} bf16+stats: {us:7.1f} us {2.0 * M * C * 3 * C / us / 1e6:7.1f} TFLOP/s")
    with ops.stats_scope(dev):
        us = graph_time(lambda: ops.conv3x3(y, B * T, H, W, w9, bias=bias, f32=False, bf16=True, stats=True))
    print(f"conv3x3 M={M:6d} C={C:5d} bf16+stats: {us:7.1f} us {2.0 * M * C * 9 * C / us / 1e6:7.1f} TFLOP/s")
    with ops.stats_scope(dev):
        us = graph_time(lambda: ops.conv3x3(y, B * T, H, W, w9, bias=bias, residual=res, stats=True))
    print(f"conv3x3 M={M:6d} C={C:5d} f32+res+stats: {us:7.1f} us {2.0 * M * C * 9 * C / us / 1e6:7.1f} TFLOP/s")
