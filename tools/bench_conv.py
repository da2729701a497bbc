import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
from tools.bench_kernels import timeit, dev
for name, nf, H, W, C, N in [("conv", 32, 40, 64, 320, 320), ("conv", 32, 20, 32, 640, 640), ("conv", 32, 10, 16, 1280, 1280),
                             ("conv", 32, 40, 64, 640, 640)]:
    a = torch.randn(nf * H * W, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, 9 * C, device=dev) * (9 * C) ** -0.5).to(torch.bfloat16)
    out = torch.empty(nf * H * W, N, device=dev)
    t = timeit(lambda: ops.conv3x3(a, nf, H, W, w, out=out))
    print(f"BN={os.environ.get('DCB_GEMM_BN')} PAIR={os.environ.get('DCB_GEMM_PAIR')} {name} {H}x{W} C={C} N={N}: {t:8.1f} us {2 * nf * H * W * N * 9 * C / t / 1e6:7.1f} TFLOP/s")
for M, N, K in [(81920, 640, 2560), (20480, 1280, 5120), (81920, 320, 1280)]:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: ops.linear(a, w, out=out))
    print(f"BN={os.environ.get('DCB_GEMM_BN')} PAIR={os.environ.get('DCB_GEMM_PAIR')} linear M={M} N={N} K={K}: {t:8.1f} us {2 * M * N * K / t / 1e6:7.1f} TFLOP/s")
