"""Single-kernel drivers for `ncu --set full` of the round-2 kernels:  python tools/prof_r02.py <which>
  cross     text + image cross-attention (attn_tc_kernel<false>) at the 40x64 level
  temporal  16-frame temporal attention (attn_tc_kernel<true>) at the 40x64 level
  vae       d = 512 VAE mid-block attention (vae_attn_tc_kernel), 8 frames of 40x64 tokens
  gn_apply  streaming GroupNorm apply from producer statistics (gn_stream_apply_kernel), 81920 x 320 fp32
  gemm_res  tap_gemm 81920x320x320 + fp32 residual + fused GroupNorm statistics (the worst GEMM shape of the step)
  gemm_conv tap_gemm 3x3 conv 81920x320x2880 + residual + statistics (a typical good shape)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
which = sys.argv[1]
BF = torch.bfloat16


def rnd(*s, dt=BF):
    return torch.randn(*s, device=dev).to(dt)


B, T = 2, 16
if which == "cross":
    H, N = 5, 2560
    inner, nb = H * 64, B * T
    q, kv_t, kv_i = rnd(nb * N, inner), rnd(B * 77, 2 * inner), rnd(nb * 16, 2 * inner)
    fn = lambda: ops.flash_attention(q, [(kv_t[:, :inner], kv_t[:, inner:], 77, T, 1.0),
                                         (kv_i[:, :inner], kv_i[:, inner:], 16, 1, 1.0)], nb, H, N, 0.125)
elif which == "temporal":
    H, HW = 5, 2560
    inner = H * 64
    qkv = rnd(B * T * HW, 3 * inner)
    fn = lambda: ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], B, T, HW, H, 0.125)
elif which == "vae":
    F, N, C = 8, 2560, 512
    qkv = rnd(F * N, 3 * C)
    fn = lambda: ops.vae_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], F, N, C, C ** -0.5)
elif which == "gn_apply":
    HW, C = 2560, 320
    x = rnd(B * T * HW, C, dt=torch.float32)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    acc = torch.ones((B * T, 32, 2), dtype=torch.int64, device=dev) * (1 << 30)
    cs = ops.ColSums(acc, HW, C, 0)
    fn = lambda: ops.group_norm(x, B * T, HW, g, b, 1e-5, silu=True, colsums=cs)
elif which in ("gemm_res", "gemm_conv"):
    HW, C = 2560, 320
    M = B * T * HW
    y, res, bias = rnd(M, C), rnd(M, C, dt=torch.float32), torch.randn(C, device=dev)
    if which == "gemm_res":
        w = rnd(C, C) * C ** -0.5
        def fn():
            with ops.stats_scope(dev):
                return ops.linear(y, w, bias=bias, residual=res, f32=True, bf16=False, stats_rows=HW)
    else:
        w = rnd(C, 9 * C) * (9 * C) ** -0.5
        def fn():
            with ops.stats_scope(dev):
                return ops.conv3x3(y, B * T, 40, 64, w, bias=bias, residual=res, stats=True)
else:
    raise SystemExit(__doc__)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok")
