"""tcgen05 implicit-GEMM kernel vs a plain torch fp32 computation on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _close(got, ref, tol=2e-3):
    err = (got.float() - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-6
    assert err / den < tol, (err, den)


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


# the last three rows: short-K, large-M launches of the 40x64 level (one of them ragged in M)
@pytest.mark.parametrize("M,K,N", [(300, 320, 320), (128, 64, 128), (1000, 1280, 512), (77, 1024, 640), (4096, 320, 960),
                                   (40000, 128, 1280), (20480, 640, 512), (20480, 320, 320), (19000, 320, 960),
                                   (40960, 256, 640),
                                   # 256-wide tiles with a narrow last tile (UMMA N = 64 / 128 / 192 / 96, fewer B rows)
                                   (19000, 320, 1920), (9600, 640, 640), (20480, 128, 352), (12000, 192, 448)])
def test_linear_bias_residual(M, K, N):
    from dynamicrafter_b200 import ops
    a, w = _mk((M, K), 1), _mk((N, K), 2, K ** -0.5)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    ref = a.float() @ w.float().t() + bias + res
    out = ops.linear(a, w, bias=bias, residual=res, f32=True, bf16=False)
    torch.cuda.synchronize()
    _close(out, ref, 1e-4)
    ob = ops.linear(a, w, bias=bias, residual=res, f32=False, bf16=True)
    _close(ob, ref, 1e-2)
    _close(ops.linear(a, w, f32=True, bf16=False), a.float() @ w.float().t(), 1e-4)
    _close(ops.linear(a, w, bias=bias), a.float() @ w.float().t() + bias, 1e-2)


def test_linear_strided_a_and_small_n():
    from dynamicrafter_b200 import ops
    big = _mk((500, 3 * 128), 3)
    a = big[:, 128:256]
    w = _mk((4, 128), 4, 0.1)
    ref = a.float() @ w.float().t()
    out = ops.linear(a, w, f32=True, bf16=False)
    _close(out, ref, 1e-4)


@pytest.mark.parametrize("K,inner,bn,M", [(128, 320, 128, 200), (64, 256, 128, 200), (320, 1280, 128, 200),
                                          (320, 1280, 256, 200), (64, 256, 256, 200), (320, 1280, 256, 20480)])
def test_geglu(K, inner, bn, M):
    from dynamicrafter_b200 import ops
    a, w = _mk((M, K), 5), _mk((2 * inner, K), 6, K ** -0.5)
    bias = torch.randn(2 * inner, device="cuda") * 0.1
    h = a.float() @ w.float().t() + bias
    ref = h[:, :inner] * F.gelu(h[:, inner:])
    half = bn // 2
    idx = torch.cat([torch.cat([torch.arange(j * half, (j + 1) * half), inner + torch.arange(j * half, (j + 1) * half)])
                     for j in range(inner // half)]).cuda()
    out = ops.linear(a, w[idx].contiguous(), bias=bias[idx].contiguous(), geglu=True, block_n=bn)
    _close(out, ref, 1e-2)


@pytest.mark.parametrize("nf,H,W,C,N", [(3, 10, 16, 64, 128), (2, 5, 8, 128, 320), (1, 40, 64, 64, 160), (4, 3, 2, 64, 64)])
def test_conv3x3(nf, H, W, C, N):
    from dynamicrafter_b200 import ops
    x = _mk((nf, H, W, C), 7)
    wt = _mk((N, C, 3, 3), 8, (9 * C) ** -0.5)
    bias = torch.randn(N, device="cuda")
    gb = torch.randn(nf, N, device="cuda")
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1) + gb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(nf * H * W, N)
    wp = wt.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    out = ops.conv3x3(x.reshape(nf * H * W, C), nf, H, W, wp, bias=bias, group_bias=gb, rows_per_group=H * W)
    _close(out, ref, 1e-4)


@pytest.mark.parametrize("B,T,HW,C,N", [(2, 16, 40, 128, 128), (1, 16, 160, 64, 320), (1, 16, 2, 64, 64)])
def test_tconv3(B, T, HW, C, N):
    from dynamicrafter_b200 import ops
    x = _mk((B, T, HW, C), 9)
    wt = _mk((N, C, 3, 1, 1), 10, (3 * C) ** -0.5)
    bias = torch.randn(N, device="cuda")
    x5 = x.float().permute(0, 3, 1, 2).unsqueeze(-1)           # [B, C, T, HW, 1]
    ref = F.conv3d(x5, wt.float(), bias, padding=(1, 0, 0)).squeeze(-1).permute(0, 2, 3, 1).reshape(B * T * HW, N)
    wp = wt.reshape(N, C, 3).permute(0, 2, 1).reshape(N, 3 * C).contiguous()
    res = torch.randn(B * T * HW, N, device="cuda")
    out = ops.tconv3(x.reshape(-1, C), B, T, HW, wp, bias=bias, residual=res)
    _close(out, ref + res, 1e-4)


def test_gemm_throughput_smoke():
    """Not a benchmark: just exercises a many-tile persistent launch and prints a rough rate."""
    from dynamicrafter_b200 import ops
    M, K, N = 40960, 1280, 1280
    a, w = _mk((M, K), 11), _mk((N, K), 12, K ** -0.5)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        ops.linear(a, w, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.linear(a, w, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"\ntap_gemm {M}x{N}x{K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")
    ref = a.float() @ w.float().t()
    _close(out, ref, 1e-2)


@pytest.mark.parametrize("nf,H,W,C,N,G", [(4, 40, 64, 64, 320, 4), (4, 10, 16, 64, 128, 4), (16, 5, 8, 64, 64, 16),
                                          (16, 5, 8, 64, 64, 1), (32, 4, 4, 64, 96, 2), (3, 20, 32, 64, 160, 3)])
def test_fused_groupnorm_statistics_conv(nf, H, W, C, N, G):
    """Column sums written by the conv epilogue -> (mean, rstd) identical to statistics of the stored output."""
    from dynamicrafter_b200 import ops
    x = _mk((nf, H, W, C), 21)
    wt = _mk((N, C, 3, 3), 22, (9 * C) ** -0.5)
    wp = wt.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    bias = torch.randn(N, device="cuda")
    for bf16 in (False, True):
        out, cs = ops.conv3x3(x.reshape(-1, C), nf, H, W, wp, bias=bias, f32=not bf16, bf16=bf16, stats=True)
        assert cs is not None
        S = nf * H * W // G
        stats = ops._stats_from_colsums(cs, G, S, N, 1e-5, out.device)
        o = out.float().reshape(G, S, 32, N // 32).permute(0, 2, 1, 3).reshape(G, 32, -1)
        mean, var = o.mean(-1), o.var(-1, unbiased=False)
        tol = 2e-3 if bf16 else 2e-5          # bf16 output is rounded after the statistics were taken
        assert (stats[:, :, 0] - mean).abs().max().item() < tol * (1 + mean.abs().max().item())
        assert ((stats[:, :, 1] - (var + 1e-5).rsqrt()).abs() / (var + 1e-5).rsqrt()).max().item() < (5e-3 if bf16 else 1e-4)
    # fp32 output with the TMA-prefetched residual (the slot is refilled by TMA: extra group barrier after the read-back)
    res = torch.randn(nf * H * W, N, device="cuda")
    out, cs = ops.conv3x3(x.reshape(-1, C), nf, H, W, wp, bias=bias, residual=res, stats=True)
    S = nf * H * W // G
    stats = ops._stats_from_colsums(cs, G, S, N, 1e-5, out.device)
    o = out.float().reshape(G, S, 32, N // 32).permute(0, 2, 1, 3).reshape(G, 32, -1)
    mean, var = o.mean(-1), o.var(-1, unbiased=False)
    assert (stats[:, :, 0] - mean).abs().max().item() < 2e-5 * (1 + mean.abs().max().item())
    assert ((stats[:, :, 1] - (var + 1e-5).rsqrt()).abs() / (var + 1e-5).rsqrt()).max().item() < 1e-4


@pytest.mark.parametrize("B,T,HW,C,N", [(2, 16, 40, 64, 128), (1, 16, 2560, 64, 64), (2, 16, 160, 64, 320)])
def test_fused_groupnorm_statistics_tconv(B, T, HW, C, N):
    from dynamicrafter_b200 import ops
    x = _mk((B * T * HW, C), 23)
    wp = _mk((N, 3 * C), 24, (3 * C) ** -0.5)
    out, cs = ops.tconv3(x, B, T, HW, wp, stats=True)
    assert cs is not None
    stats = ops._stats_from_colsums(cs, B, T * HW, N, 1e-5, out.device)
    o = out.reshape(B, T * HW, 32, N // 32).permute(0, 2, 1, 3).reshape(B, 32, -1)
    mean, var = o.mean(-1), o.var(-1, unbiased=False)
    assert (stats[:, :, 0] - mean).abs().max().item() < 2e-5
    assert ((stats[:, :, 1] - (var + 1e-5).rsqrt()).abs() / (var + 1e-5).rsqrt()).max().item() < 1e-4


@pytest.mark.parametrize("M,C,N,geglu", [(20480, 320, 960, False), (5000, 640, 640, False), (4096, 320, 2560, True),
                                         (300, 1280, 10240, True), (20480, 512, 1536, False), (1000, 320, 320, False)])
def test_layer_norm_folding(M, C, N, geglu):
    """LayerNorm folded into its neighbour GEMMs (attention.py:242-246): the producer GEMM leaves a bf16 copy + per-row
    (sum, sumsq), the consumer GEMM runs on gamma-scaled weights and normalises in its epilogue.  Checked against
    LayerNorm -> Linear (-> GEGLU) in fp32 torch; rows get a non-zero mean so the mean correction matters."""
    import torch.nn as nn
    from dynamicrafter_b200 import ops
    from dynamicrafter_b200.lvdm.modules.attention import fold_layer_norm, _fold_ln_weights
    from dynamicrafter_b200.lvdm._packing import pack_geglu
    torch.manual_seed(5)
    a, w0 = _mk((M, C), 31), _mk((C, C), 32, C ** -0.5)
    b0 = torch.randn(C, device="cuda") * 0.3 + 0.4
    res = torch.randn(M, C, device="cuda") + 0.25
    y_ref = a.float() @ w0.float().t() + b0 + res
    y, src = ops.linear(a, w0, bias=b0, residual=res, f32=True, bf16=False, ln_stats=True)
    _close(y, y_ref, 1e-4)
    _close(src.xb, y_ref, 1e-2)
    # per-row statistics
    acc = src.acc.view(M, 2).double()
    assert (acc[:, 0] / 2 ** 24 - y.double().sum(1)).abs().max().item() < 1e-3
    assert ((acc[:, 1] / 2 ** 28 - (y.double() ** 2).sum(1)).abs() / (y.double() ** 2).sum(1)).max().item() < 1e-5
    norm = nn.LayerNorm(C).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C) * 0.2 + 1.0)
        norm.bias.copy_(torch.randn(C) * 0.2)
    w1 = torch.randn(N, C, device="cuda") * C ** -0.5
    b1 = torch.randn(N, device="cuda") * 0.1
    with torch.no_grad():
        h = F.layer_norm(y_ref, (C,), norm.weight, norm.bias, norm.eps) @ w1.t() + b1
    if geglu:
        inner = N // 2
        ref = h[:, :inner] * F.gelu(h[:, inner:])
        wf, bf = _fold_ln_weights(norm, w1, b1)
        wq, bq, bn = pack_geglu(wf, bf)
        fold = ops.LnFold(wq, bq, wq.float().sum(1).contiguous(), norm.eps)
        out = ops.linear(src, None, geglu=True, block_n=bn, ln=fold)
    else:
        ref = h
        out = ops.linear(src, None, ln=fold_layer_norm(norm, w1, b1))
    torch.cuda.synchronize()
    _close(out, ref, 1.5e-2)
    # and against the unfolded kernel path (layer_norm kernel + plain GEMM): same precision class
    xn = ops.layer_norm(y, norm.weight, norm.bias, norm.eps)
    if not geglu:
        out2 = ops.linear(xn, w1.to(torch.bfloat16).contiguous(), bias=b1)
        e_fold = (out.float() - ref).norm() / ref.norm()
        e_plain = (out2.float() - ref).norm() / ref.norm()
        assert e_fold < 2.0 * e_plain + 1e-3, (float(e_fold), float(e_plain))


@pytest.mark.parametrize("nf,H,W,C,N", [(2, 5, 8, 128, 128), (3, 10, 16, 64, 320), (1, 40, 64, 64, 160), (2, 20, 128, 64, 64),
                                        (4, 3, 2, 64, 64)])
def test_conv3x3_upsample_folded(nf, H, W, C, N):
    """Nearest-x2 upsample + 3x3 conv (openaimodel3d.py:98-106, ae_modules.py:123-127) as four 2x2-tap phase GEMMs on the
    low-resolution tensor, against F.interpolate + F.conv2d; also the fused GroupNorm statistics of the scattered output."""
    from dynamicrafter_b200 import ops
    from dynamicrafter_b200.lvdm._packing import pack_conv3x3_up2
    x = _mk((nf, H, W, C), 41)
    wt = _mk((N, C, 3, 3), 42, (9 * C) ** -0.5)
    bias = torch.randn(N, device="cuda")
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, wt.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(nf * 4 * H * W, N)
    w4 = pack_conv3x3_up2(wt.float())
    out, cs = ops.conv3x3_up2(x.reshape(nf * H * W, C), nf, H, W, w4, bias=bias, stats=True)
    torch.cuda.synchronize()
    _close(out, ref, 4e-3)          # the pre-summed taps are rounded to bf16 once more than the 3x3 weights
    ob = ops.conv3x3_up2(x.reshape(nf * H * W, C), nf, H, W, w4, bias=bias, f32=False, bf16=True)
    _close(ob, ref, 1.5e-2)
    if cs is not None and N % 32 == 0:
        S = 4 * H * W
        stats = ops._stats_from_colsums(cs, nf, S, N, 1e-5, out.device)
        o = out.float().reshape(nf, S, 32, N // 32).permute(0, 2, 1, 3).reshape(nf, 32, -1)
        mean, var = o.mean(-1), o.var(-1, unbiased=False)
        assert (stats[:, :, 0] - mean).abs().max().item() < 2e-5 * (1 + mean.abs().max().item())
        assert ((stats[:, :, 1] - (var + 1e-5).rsqrt()).abs() / (var + 1e-5).rsqrt()).max().item() < 1e-4
