"""CUDA-event timing of the short-key attention kernels at the 512 config's shapes (tcgen05 vs mma.sync via env)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamicrafter_b200 import ops


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


g = torch.Generator().manual_seed(1)
for H, N in ((5, 2560), (10, 640), (20, 160), (20, 40)):
    B, T, NT, NI = 2, 16, 77, 16
    inner, nb = H * 64, B * T
    q = torch.randn((nb * N, inner), generator=g).cuda().bfloat16()
    kt = torch.randn((B * NT, 2 * inner), generator=g).cuda().bfloat16()
    ki = torch.randn((nb * NI, 2 * inner), generator=g).cuda().bfloat16()
    us = timeit(lambda: ops.flash_attention(q, [(kt[:, :inner], kt[:, inner:], NT, T, 1.0), (ki[:, :inner], ki[:, inner:], NI, 1, 1.0)], nb, H, N, 0.125))
    gb = 2 * q.numel() * 2 / 1e9
    print(f"cross   H={H:2d} N={N:5d}: {us:7.1f} us  {gb / (us * 1e-6):7.0f} GB/s of Q+O")
for H, HW in ((5, 2560), (8, 2560), (10, 640), (20, 160), (20, 40)):
    B, L = 2, 16
    inner = H * 64
    qkv = torch.randn((B * L * HW, 3 * inner), generator=g).cuda().bfloat16()
    us = timeit(lambda: ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], B, L, HW, H, 0.125))
    gb = 4 * B * L * HW * inner * 2 / 1e9
    print(f"temporal H={H:2d} HW={HW:5d}: {us:7.1f} us  {gb / (us * 1e-6):7.0f} GB/s of q,k,v,o")
for F_, N, C in ((16, 2560, 512), (4, 9216, 512)):
    qkv = torch.randn((F_ * N, 3 * C), generator=g).cuda().bfloat16()
    us = timeit(lambda: ops.vae_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], F_, N, C, C ** -0.5), n=5)
    fl = 4.0 * F_ * N * N * C
    print(f"vae attn F={F_} N={N} C={C}: {us:8.1f} us  {fl / (us * 1e-6) / 1e12:6.1f} TFLOP/s algorithmic")
