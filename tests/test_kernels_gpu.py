"""Each CUDA kernel (through the C ABI) against a plain torch fp32 computation of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _randn(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).cuda()


def _rel(got, ref):
    return ((got.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


# ------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("G,S,C1,C2,eps,silu", [(3, 40, 64, 0, 1e-5, True), (2, 640, 320, 0, 1e-6, False),
                                                (2, 200, 128, 64, 1e-5, True), (1, 4096, 1280, 640, 1e-5, True),
                                                (16, 1, 64, 0, 1e-5, True)])
def test_group_norm(G, S, C1, C2, eps, silu):
    from dynamicrafter_b200 import ops
    x1 = _randn((G * S, C1), 1, 2.0, 0.7)
    x2 = _randn((G * S, C2), 2, 0.5, -0.3) if C2 else None
    C = C1 + C2
    gamma, beta = _randn((C,), 3, 0.2, 1.0), _randn((C,), 4, 0.1)
    x = x1 if x2 is None else torch.cat([x1, x2], 1)
    ref = F.group_norm(x.reshape(G, S, C).permute(0, 2, 1), 32, gamma, beta, eps).permute(0, 2, 1).reshape(G * S, C)
    if silu:
        ref = F.silu(ref)
    y, raw = ops.group_norm(x1, G, S, gamma, beta, eps, x2=x2, silu=silu, want_raw=True)
    assert _rel(y, ref) < 4e-3                       # bf16 output rounding
    assert torch.equal(raw, x.to(torch.bfloat16))
    # second launch reuses the self-cleaning counters
    y2 = ops.group_norm(x1, G, S, gamma, beta, eps, x2=x2, silu=silu)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("rows,C", [(100, 64), (1000, 320), (77, 512), (640, 1280)])
def test_layer_norm(rows, C):
    from dynamicrafter_b200 import ops
    x = _randn((rows, C), 5, 1.5, 0.3)
    g, b = _randn((C,), 6, 0.2, 1.0), _randn((C,), 7, 0.1)
    ref = F.layer_norm(x, (C,), g, b, 1e-5)
    assert _rel(ops.layer_norm(x, g, b, 1e-5), ref) < 4e-3


# ------------------------------------------------------------------------------------------ attention
def _sdpa(q, k, v, scale):
    return torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1) @ v


@pytest.mark.parametrize("B,H,N", [(2, 2, 200), (4, 20, 160), (1, 3, 130), (1, 5, 2560), (3, 1, 40), (2, 3, 640), (1, 2, 16), (2, 2, 300), (1, 1, 257),
                                   (3, 2, 256), (2, 1, 1000)])
def test_flash_self_attention(B, H, N):
    from dynamicrafter_b200 import ops
    inner = H * 64
    qkv = _randn((B * N, 3 * inner), 8).to(torch.bfloat16)
    out = ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], N, 1, 1.0)], B, H, N, 0.125)
    f = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = _sdpa(f[0], f[1], f[2], 0.125).permute(0, 2, 1, 3).reshape(B * N, inner)
    assert _rel(out, ref) < 1e-2


@pytest.mark.parametrize("B,T,H,N,NT,NI", [(1, 16, 2, 160, 77, 16), (2, 4, 1, 70, 77, 16), (1, 2, 5, 2560, 77, 16),
                                           (1, 3, 2, 33, 77, 16), (2, 2, 2, 640, 50, 9), (1, 2, 2, 300, 100, 16)])
def test_flash_dual_cross_attention(B, T, H, N, NT, NI):
    """text keys shared by the T frames of a video + per-frame image keys, two independent softmaxes
    (resident-key kernel when the keys fit 80 + 16, tiled kernel otherwise)."""
    from dynamicrafter_b200 import ops
    inner, nb = H * 64, B * T
    q = _randn((nb * N, inner), 9).to(torch.bfloat16)
    kv_t = _randn((B * NT, 2 * inner), 10).to(torch.bfloat16)
    kv_i = _randn((nb * NI, 2 * inner), 11).to(torch.bfloat16)
    s_img = 1.37
    out = ops.flash_attention(q, [(kv_t[:, :inner], kv_t[:, inner:], NT, T, 1.0),
                                  (kv_i[:, :inner], kv_i[:, inner:], NI, 1, s_img)], nb, H, N, 0.125)
    qf = q.float().reshape(nb, N, H, 64).permute(0, 2, 1, 3)
    kt = kv_t.float().reshape(B, NT, 2, H, 64).permute(2, 0, 3, 1, 4).repeat_interleave(T, dim=1)
    ki = kv_i.float().reshape(nb, NI, 2, H, 64).permute(2, 0, 3, 1, 4)
    ref = _sdpa(qf, kt[0], kt[1], 0.125) + s_img * _sdpa(qf, ki[0], ki[1], 0.125)
    assert _rel(out, ref.permute(0, 2, 1, 3).reshape(nb * N, inner)) < 1e-2


@pytest.mark.parametrize("B,L,HW,H", [(2, 16, 40, 5), (1, 16, 7, 8), (1, 4, 33, 2), (2, 16, 2, 2), (2, 16, 8, 1), (2, 16, 1, 2),
                                      (2, 16, 4, 2)])
def test_temporal_attention(B, L, HW, H):
    from dynamicrafter_b200 import ops
    inner = H * 64
    qkv = _randn((B * L * HW, 3 * inner), 12).to(torch.bfloat16)
    out = ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], B, L, HW, H, 0.125)
    f = qkv.float().reshape(B, L, HW, 3, H, 64).permute(3, 0, 2, 4, 1, 5)       # [3, B, HW, H, L, 64]
    ref = _sdpa(f[0], f[1], f[2], 0.125).permute(0, 3, 1, 2, 4).reshape(B * L * HW, inner)
    assert _rel(out, ref) < 1e-2


def test_softmax_rows():
    from dynamicrafter_b200 import ops
    x = _randn((300, 777), 13, 3.0)
    assert _rel(ops.softmax_rows(x, 0.3), torch.softmax(x * 0.3, -1)) < 5e-3


# ------------------------------------------------------------------------------------------ layout
def test_layout_roundtrip_and_resample():
    from dynamicrafter_b200 import ops
    x = _randn((2, 8, 3, 5, 6), 14)
    rows = ops.nchw_to_rows(x, 64)
    ref = x.permute(0, 2, 3, 4, 1).reshape(-1, 8).to(torch.bfloat16)
    assert torch.equal(rows[:, :8], ref) and float(rows[:, 8:].abs().max()) == 0.0
    r32 = _randn((2 * 3 * 5 * 6, 12), 15)
    back = ops.rows_to_nchw(r32, 2, 4, 3, 5, 6)
    assert torch.equal(back, r32[:, :4].reshape(2, 3, 5, 6, 4).permute(0, 4, 1, 2, 3))
    f = _randn((3 * 5 * 6, 16), 16)
    up = ops.upsample2x(f, 3, 5, 6)
    ref_up = F.interpolate(f.reshape(3, 5, 6, 16).permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    assert torch.equal(up, ref_up.permute(0, 2, 3, 1).reshape(-1, 16).to(torch.bfloat16))


@pytest.mark.parametrize("nf,H,W,C,N", [(2, 10, 16, 64, 128), (3, 5, 7, 64, 64), (1, 40, 64, 128, 160)])
def test_stride2_conv_via_phase_planes(nf, H, W, C, N):
    from dynamicrafter_b200 import ops
    x = _randn((nf * H * W, C), 17)
    wt = _randn((N, C, 3, 3), 18, (9 * C) ** -0.5)
    bias = _randn((N,), 19)
    xb = x.to(torch.bfloat16).float().reshape(nf, H, W, C).permute(0, 3, 1, 2)
    ref = F.conv2d(xb, wt.to(torch.bfloat16).float(), bias, stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    wp = wt.permute(0, 2, 3, 1).reshape(N, 9 * C).to(torch.bfloat16).contiguous()
    out = ops.conv3x3_s2(ops.space_to_depth(x, nf, H, W), nf, H, W, wp, bias=bias)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 1e-4


def test_embeddings():
    from dynamicrafter_b200 import ops
    t = torch.tensor([999, 601, 1, 24], dtype=torch.long).cuda()
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).cuda()
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    got = ops.timestep_embedding(t, 320)
    assert (got - ref).abs().max().item() < 2e-4
    a, w, b = _randn((4, 320), 20), _randn((1280, 320), 21, 0.05), _randn((1280,), 22)
    out = ops.small_linear(a, w, b, silu_out=True)
    assert _rel(out, F.silu(a @ w.t() + b)) < 1e-5
    out2 = ops.small_linear(out, _randn((64, 1280), 23, 0.03), None, silu_in=True, out=torch.ones(4, 64).cuda(), accumulate=True)
    assert _rel(out2, 1 + F.silu(out) @ _randn((64, 1280), 23, 0.03).t()) < 1e-5


# ------------------------------------------------------------------------------------------ DDIM update
@pytest.mark.parametrize("is_v,gr,cfg", [(1, 0.0, True), (0, 0.0, True), (1, 0.7, True), (0, 0.0, False)])
def test_ddim_update_matches_reference_arithmetic(is_v, gr, cfg):
    from dynamicrafter_b200 import ops
    from dynamicrafter_b200._lib import DdimCoeffs
    shape = (2, 4, 16, 8, 8)
    x, ec, eu, nz = (_randn(shape, s) for s in (30, 31, 32, 33))
    c = DdimCoeffs()
    vals = dict(is_v=is_v, cfg_scale=7.5, guidance_rescale=gr, one_minus_guidance_rescale=1 - gr, sqrt_acp_t=0.6234,
                sqrt_1m_acp_t=0.7819, sqrt_a_t=0.6234, sqrt_1m_a_t=0.7819, sqrt_a_prev=0.71, dir_coef=0.55,
                sigma_t=0.33, temperature=1.0, x0_rescale=1.0123)
    for k, v in vals.items():
        setattr(c, k, v)
    f = {k: torch.tensor(getattr(c, k), dtype=torch.float32).cuda() for k in vals if k != "is_v"}
    xp, x0 = ops.ddim_update(x, ec, eu if cfg else None, nz, c)
    mo = ec
    if cfg:
        mo = eu + f["cfg_scale"] * (ec - eu)
        if gr > 0:
            dims = list(range(1, 5))
            fac = ec.std(dim=dims, keepdim=True) / mo.std(dim=dims, keepdim=True)
            mo = f["guidance_rescale"] * (mo * fac) + f["one_minus_guidance_rescale"] * mo
    if is_v:
        e_t = f["sqrt_acp_t"] * mo + f["sqrt_1m_acp_t"] * x
        p0 = f["sqrt_acp_t"] * x - f["sqrt_1m_acp_t"] * mo
    else:
        e_t = mo
        p0 = (x - f["sqrt_1m_a_t"] * e_t) / f["sqrt_a_t"]
    p0 = p0 * f["x0_rescale"]
    ref = f["sqrt_a_prev"] * p0 + f["dir_coef"] * e_t + f["sigma_t"] * nz * f["temperature"]
    if gr == 0:
        assert torch.equal(x0, p0) and torch.equal(xp, ref)         # bit exact: same fp32 ops in the same order
    else:
        assert _rel(x0, p0) < 1e-6 and _rel(xp, ref) < 1e-6


@pytest.mark.parametrize("G,S,C", [(3, 40, 64), (2, 640, 320), (1, 4096, 1280)])
def test_group_norm_bf16_source(G, S, C):
    from dynamicrafter_b200 import ops
    x = _randn((G * S, C), 41, 2.0, 0.7).to(torch.bfloat16)
    gamma, beta = _randn((C,), 3, 0.2, 1.0), _randn((C,), 4, 0.1)
    ref = F.silu(F.group_norm(x.float().reshape(G, S, C).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(G * S, C)
    y = ops.group_norm(x, G, S, gamma, beta, 1e-5, silu=True)
    assert _rel(y, ref) < 4e-3


def test_video_to_uint8_byte_exact():
    """output stage (inference.py:153-159): byte-exact against the reference's torch ops, including out-of-range
    values, the exact endpoints and values sitting on uint8 boundaries."""
    from dynamicrafter_b200 import ops
    from oracle import output_oracle
    x = _randn((2, 3, 5, 17, 23), 31, 0.9)
    x[0, 0, 0, 0, :8] = torch.tensor([-1.0, 1.0, -3.0, 7.0, 0.0, -0.999999, 0.999999, 0.00392157], device=x.device)
    k = torch.arange(0, 256, device=x.device, dtype=torch.float32)
    x[1, 1, 2].reshape(-1)[:256] = k / 255.0 * 2.0 - 1.0            # values that land on (or next to) integer levels
    got = ops.video_to_uint8(x)
    ref = output_oracle.video_to_uint8(x)
    assert got.dtype == torch.uint8 and got.shape == (2, 5, 17, 23, 3)
    assert torch.equal(got.cpu(), ref)


@pytest.mark.parametrize("F_,N,C", [(2, 96, 256), (1, 2560, 512), (3, 300, 512), (2, 129, 128), (1, 640, 384), (2, 40, 64)])
def test_vae_wide_head_attention(F_, N, C):
    """dcb_vae_attention: single head of width C per frame (ae_modules.py:61-74), ragged token counts, 1-2 output slices."""
    from dynamicrafter_b200 import ops
    qkv = _randn((F_ * N, 3 * C), 44, 0.7).to(torch.bfloat16)
    scale = C ** -0.5
    out = ops.vae_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], F_, N, C, scale)
    f = qkv.float().reshape(F_, N, 3, C).permute(2, 0, 1, 3)
    ref = _sdpa(f[0], f[1], f[2], scale).reshape(F_ * N, C)
    assert _rel(out, ref) < 1e-2


def test_temporal_attention_large_and_ragged():
    """tcgen05 temporal kernel: location counts that are not multiples of the 8-location tile, several CTAs' worth."""
    from dynamicrafter_b200 import ops
    for B, L, HW, H in [(2, 16, 2560, 5), (1, 16, 643, 2), (2, 9, 50, 3)]:
        inner = H * 64
        qkv = _randn((B * L * HW, 3 * inner), 12).to(torch.bfloat16)
        out = ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], B, L, HW, H, 0.125)
        f = qkv.float().reshape(B, L, HW, 3, H, 64).permute(3, 0, 2, 4, 1, 5)
        ref = _sdpa(f[0], f[1], f[2], 0.125).permute(0, 3, 1, 2, 4).reshape(B * L * HW, inner)
        assert _rel(out, ref) < 1e-2, (B, L, HW, H)


def test_cross_attention_bench_shape():
    """tcgen05 cross kernel at the 512 config's finest level: 32 frames x 5 heads x 2560 queries, 77 + 16 keys."""
    from dynamicrafter_b200 import ops
    B, T, H, N, NT, NI = 2, 16, 5, 2560, 77, 16
    inner, nb = H * 64, B * T
    q = _randn((nb * N, inner), 9).to(torch.bfloat16)
    kv_t = _randn((B * NT, 2 * inner), 10).to(torch.bfloat16)
    kv_i = _randn((nb * NI, 2 * inner), 11).to(torch.bfloat16)
    out = ops.flash_attention(q, [(kv_t[:, :inner], kv_t[:, inner:], NT, T, 1.0),
                                  (kv_i[:, :inner], kv_i[:, inner:], NI, 1, 0.8)], nb, H, N, 0.125)
    qf = q.float().reshape(nb, N, H, 64).permute(0, 2, 1, 3)
    kt = kv_t.float().reshape(B, NT, 2, H, 64).permute(2, 0, 3, 1, 4).repeat_interleave(T, dim=1)
    ki = kv_i.float().reshape(nb, NI, 2, H, 64).permute(2, 0, 3, 1, 4)
    ref = _sdpa(qf, kt[0], kt[1], 0.125) + 0.8 * _sdpa(qf, ki[0], ki[1], 0.125)
    assert _rel(out, ref.permute(0, 2, 1, 3).reshape(nb * N, inner)) < 1e-2


@pytest.mark.parametrize("Z,A,B,inner", [(2, 8, 2, 1152 * 320), (1, 4, 4, 40 * 64), (3, 2, 5, 8)])
def test_block_transpose(Z, A, B, inner):
    """[Z, A, B, inner] -> [Z, B, A, inner] (frame <-> pixel re-ordering of the frame-sharded mode), plain and fused with
    the fp32 residual add of a bf16 delta."""
    from dynamicrafter_b200 import ops
    torch.manual_seed(0)
    for dt in (torch.bfloat16, torch.float32):
        src = torch.randn(Z, A, B, inner, device="cuda").to(dt)
        out = ops.block_transpose(src, Z, A, B)
        assert torch.equal(out.view(Z, B, A, inner), src.permute(0, 2, 1, 3).contiguous())
    src = torch.randn(Z, A, B, inner, device="cuda").to(torch.bfloat16)
    res = torch.randn(Z, B, A, inner, device="cuda")
    out = ops.block_transpose(src, Z, A, B, residual=res)
    assert torch.equal(out, res + src.permute(0, 2, 1, 3).float())


@pytest.mark.parametrize("nf,H,W,C1,C2", [(4, 10, 16, 128, 128), (2, 40, 64, 320, 320), (3, 4, 8, 64, 64)])
def test_group_norm_concat_from_two_producers(nf, H, W, C1, C2):
    """Skip-concat GroupNorm (openaimodel3d.py:596-597) served by the fused statistics of BOTH producers (one streaming pass)
    vs the self-contained cooperative kernel on the same tensors; also the raw bf16 copy of the concatenation."""
    from dynamicrafter_b200 import ops
    torch.manual_seed(3)
    dev = "cuda"
    S = H * W
    outs = []
    for C in (C1, C2):
        a = torch.randn(nf * S, 64, device=dev).to(torch.bfloat16)
        w = (torch.randn(C, 9 * 64, device=dev) * (9 * 64) ** -0.5).to(torch.bfloat16)
        x, cs = ops.conv3x3(a, nf, H, W, w, bias=torch.randn(C, device=dev), stats=True)
        assert cs is not None
        outs.append((x, cs))
    (x1, cs1), (x2, cs2) = outs
    gamma, beta = torch.randn(C1 + C2, device=dev) * 0.2 + 1.0, torch.randn(C1 + C2, device=dev) * 0.2
    y_ref, raw_ref = ops.group_norm(x1, nf, S, gamma, beta, 1e-5, x2=x2, silu=True, want_raw=True)
    y, raw = ops.group_norm(x1, nf, S, gamma, beta, 1e-5, x2=x2, silu=True, want_raw=True, colsums=cs1, colsums2=cs2)
    torch.cuda.synchronize()
    assert torch.equal(raw, raw_ref)
    err = (y.float() - y_ref.float()).abs().max().item()
    assert err < 2e-2 * (1 + y_ref.float().abs().max().item()), err          # bf16 outputs of two statistics paths
    ref = F.group_norm(torch.cat([x1, x2], 1).reshape(nf, S, C1 + C2).permute(0, 2, 1), 32, gamma, beta, 1e-5)
    ref = F.silu(ref).permute(0, 2, 1).reshape(nf * S, C1 + C2)
    assert (y.float() - ref).abs().max().item() < 3e-2 * (1 + ref.abs().max().item())
