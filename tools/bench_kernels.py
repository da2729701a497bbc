#!/usr/bin/env python
"""Micro-benchmarks of individual dcb200 kernels at the 512-config shapes (CUDA events, L2 flushed between
iterations by writing a 256 MB buffer).  Usage: python tools/bench_kernels.py [gn] [ln] [attn] [gemm] [tattn]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dynamicrafter_b200 import ops  # noqa: E402

dev = torch.device("cuda")
_flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timeit(fn, iters=10, flush=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us


def bench_gn():
    import ctypes
    from dynamicrafter_b200 import _lib
    L = _lib.lib()
    for G, S, C in [(2, 40960, 320), (2, 10240, 640), (2, 2560, 1280), (2, 640, 1280), (32, 2560, 320), (32, 640, 640),
                    (32, 160, 1280), (32, 40, 1280), (32, 2560, 640)]:
        x = torch.randn(G * S, C, device=dev)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ws, cnt = ops._gn_scratch(dev, G)
        stats = torch.empty((G, 32, 2), device=dev)
        y = torch.empty((G * S, C), dtype=torch.bfloat16, device=dev)
        st = ops._stream
        f_stats = lambda: L.dcb_group_norm_stats(ops.ptr(x), None, C, 0, G, S, ctypes.c_float(1e-5), ops.ptr(stats),
                                                 ops.ptr(ws), ops.ptr(cnt), st())
        f_apply = lambda: L.dcb_group_norm_apply(ops.ptr(x), None, C, 0, G, S, ops.ptr(stats), ops.ptr(g), ops.ptr(b), 1,
                                                 ops.ptr(y), None, st())
        t1, t2 = timeit(f_stats), timeit(f_apply)
        n = G * S * C
        print(f"gn G={G:3d} S={S:6d} C={C:5d}: stats {t1:7.1f} us {n * 4 / t1 / 1e3:7.0f} GB/s | apply {t2:7.1f} us "
              f"{n * 6 / t2 / 1e3:7.0f} GB/s")


def bench_ln():
    for rows, C in [(81920, 320), (20480, 640), (5120, 1280), (81920, 512)]:
        x = torch.randn(rows, C, device=dev)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        t = timeit(lambda: ops.layer_norm(x, g, b))
        print(f"ln rows={rows} C={C}: {t:7.1f} us {rows * C * 6 / t / 1e3:7.0f} GB/s")


def bench_attn():
    for B, H, N in [(32, 5, 2560), (32, 10, 640), (32, 20, 160), (32, 20, 40)]:
        inner = H * 64
        qkv = torch.randn(B * N, 3 * inner, device=dev).to(torch.bfloat16)
        out = torch.empty(B * N, inner, dtype=torch.bfloat16, device=dev)
        t = timeit(lambda: ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], N, 1, 1.0)],
                                               B, H, N, 0.125, out=out))
        print(f"flash self B={B} H={H} N={N}: {t:8.1f} us {4 * B * H * N * N * 64 / t / 1e6:7.1f} TFLOP/s")
    for B, H, N in [(32, 5, 2560), (32, 10, 640), (32, 20, 160)]:
        inner = H * 64
        q = torch.randn(B * N, inner, device=dev).to(torch.bfloat16)
        kt = torch.randn(2 * 77, 2 * inner, device=dev).to(torch.bfloat16)
        ki = torch.randn(B * 16, 2 * inner, device=dev).to(torch.bfloat16)
        t = timeit(lambda: ops.flash_attention(q, [(kt[:, :inner], kt[:, inner:], 77, 16, 1.0),
                                                   (ki[:, :inner], ki[:, inner:], 16, 1, 1.0)], B, H, N, 0.125))
        print(f"flash cross B={B} H={H} N={N}: {t:8.1f} us  ({B * N * inner * 4 / t / 1e3:6.0f} GB/s q+o)")


def bench_tattn():
    for B, HW, H in [(2, 2560, 5), (2, 2560, 8), (2, 640, 10), (2, 160, 20), (2, 40, 20)]:
        inner = H * 64
        qkv = torch.randn(B * 16 * HW, 3 * inner, device=dev).to(torch.bfloat16)
        t = timeit(lambda: ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], B, 16, HW,
                                                  H, 0.125))
        print(f"temporal B={B} HW={HW} H={H}: {t:7.1f} us {B * 16 * HW * inner * 8 / t / 1e3:7.0f} GB/s")


def bench_gemm():
    cases = [("linear f32+res", 81920, 320, 320, True, True, False), ("linear f32+res", 20480, 640, 640, True, True, False),
             ("linear f32+res", 5120, 1280, 1280, True, True, False), ("qkv bf16", 81920, 960, 320, False, False, False),
             ("geglu", 81920, 2560, 320, False, False, True), ("geglu", 20480, 5120, 640, False, False, True),
             ("geglu", 5120, 10240, 1280, False, False, True), ("ff2 bf16+res", 81920, 320, 1280, False, True, False),
             ("big bf16", 40960, 1280, 1280, False, False, False), ("big bf16", 8192, 8192, 8192, False, False, False)]
    for name, M, N, K, f32, res, geglu in cases:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        r = torch.randn(M, N // 2 if geglu else N, device=dev) if res else None
        out = torch.empty(M, N // 2 if geglu else N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
        t = timeit(lambda: ops.linear(a, w, bias=bias, residual=r, geglu=geglu, out=out))
        byt = M * K * 2 + N * K * 2 + out.numel() * out.element_size() + (r.numel() * 4 if res else 0)
        print(f"gemm {name:15s} M={M:6d} N={N:5d} K={K:5d}: {t:8.1f} us {2 * M * N * K / t / 1e6:7.1f} TFLOP/s "
              f"{byt / t / 1e3:6.0f} GB/s")
    for name, nf, H, W, C, N in [("conv3x3", 32, 40, 64, 320, 320), ("conv3x3", 32, 20, 32, 640, 640),
                                 ("conv3x3", 32, 10, 16, 1280, 1280), ("conv3x3", 32, 5, 8, 1280, 1280),
                                 ("conv3x3 vae", 8, 320, 512, 128, 128), ("conv3x3 vae", 8, 160, 256, 256, 256)]:
        a = torch.randn(nf * H * W, C, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, 9 * C, device=dev) * (9 * C) ** -0.5).to(torch.bfloat16)
        out = torch.empty(nf * H * W, N, device=dev)
        t = timeit(lambda: ops.conv3x3(a, nf, H, W, w, out=out))
        print(f"{name:12s} nf={nf} {H}x{W} C={C} N={N}: {t:8.1f} us {2 * nf * H * W * N * 9 * C / t / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["gn", "ln", "attn", "tattn", "gemm"]
    for w in which:
        globals()["bench_" + w]()
