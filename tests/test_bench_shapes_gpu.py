"""Parity on the shapes bench.py actually times (VERDICT r01 "what's weak" #1).

  * C2: the full 1.44 B-parameter 512 config, B = 2 guidance pair as ONE forward, 16 frames, 40x64 latent, CUDA graph on
    -- each branch against a B = 1 forward of the CPU oracle (pinned to the reference, tests/test_oracle_golden.py);
  * C1: the 256 config at its 32x32 latent;
  * C4-shaped kernels: flash attention over 9216 / 2304 keys, the VAE mid-block attention at 72x128 (d = 512),
    GroupNorm over 147 456 rows.
The oracle forwards cost the host cores ~1-2 minutes in total on the GPU box.
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from dynamicrafter_b200.synth import synth_state_dict, synth_input

pytestmark = pytest.mark.gpu

# per-forward bars: 1.5x the values measured on B200 (printed by the tests), under north_star's 2e-2
C2_TOL = 1.7e-2
C1_TOL = 1.7e-2


def _full_unet(cfg_name, seed):
    from dynamicrafter_b200.configs import model_params
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    cfg = model_params(cfg_name)["unet_config"]["params"]
    m = UNetModel(**cfg)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict(sd)
    return m.cuda().eval(), sd, cfg


def test_unet_c2_full_shape_cfg_pair_graph_vs_oracle():
    """What bench.py times: x repeated for the cond / uncond pair, two different contexts, fs = 24, 16 f, 40x64."""
    from oracle import unet_oracle
    m, sd, cfg = _full_unet("512", 3)
    T, h, w = 16, 40, 64
    x1 = synth_input("x", (1, 8, T, h, w), 3)
    ctx = torch.cat([synth_input("cond", (1, 77 + 16 * T, 1024), 3), synth_input("uncond", (1, 77 + 16 * T, 1024), 3)], 0)
    x = torch.cat([x1, x1], 0)
    ts = torch.tensor([601, 601], dtype=torch.long)
    fs = torch.tensor([24, 24], dtype=torch.long)
    m.use_cuda_graph = True
    y = m(x.cuda(), ts.cuda(), context=ctx.cuda(), fs=fs.cuda()).cpu()
    y2 = m(x.cuda(), ts.cuda(), context=ctx.cuda(), fs=fs.cuda()).cpu()
    assert torch.equal(y, y2)                                   # graph replay is deterministic
    m.use_cuda_graph = False
    ye = m(x.cuda(), ts.cuda(), context=ctx.cuda(), fs=fs.cuda()).cpu()
    assert torch.equal(y, ye)                                   # graph replay == eager launches at the timed shape
    torch.set_num_threads(max(1, torch.get_num_threads()))
    errs = []
    for b in range(2):
        ref = unet_oracle.unet_forward(sd, cfg, x[b:b + 1], ts[b:b + 1], ctx[b:b + 1], fs[b:b + 1])
        errs.append(rel_l2(y[b:b + 1], ref))
    print(f"\nC2 full shape (B=2 pair, 16 f, 40x64, CUDA graph): per-branch rel-L2 vs oracle {errs[0]:.3e} / {errs[1]:.3e}")
    assert max(errs) < C2_TOL


def test_unet_c1_full_shape_vs_oracle():
    """BASELINE.json configs[0]: the 256 model (learnable image cross-attention scale, eps-param) at latent 32x32x16f."""
    from oracle import unet_oracle
    m, sd, cfg = _full_unet("256", 5)
    T, h, w = 16, 32, 32
    x = synth_input("x", (1, 8, T, h, w), 5)
    ctx = synth_input("context", (1, 77 + 16 * T, 1024), 5)
    ts = torch.tensor([901], dtype=torch.long)
    fs = torch.tensor([3], dtype=torch.long)
    m.use_cuda_graph = True
    y = m(x.cuda(), ts.cuda(), context=ctx.cuda(), fs=fs.cuda()).cpu()
    ref = unet_oracle.unet_forward(sd, cfg, x, ts, ctx, fs)
    err = rel_l2(y, ref)
    print(f"\nC1 full shape (256 config, 16 f, 32x32): rel-L2 vs oracle {err:.3e}")
    assert err < C1_TOL


# ------------------------------------------------------------------------------------------ C4-shaped kernels
def _randn(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale + shift).cuda()


def _rel(got, ref):
    return ((got.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,H,N", [(2, 5, 9216), (2, 10, 2304)])
def test_flash_self_attention_c4_lengths(B, H, N):
    """Spatial self-attention key counts of the 576x1024 config (72x128 and 36x64 latents)."""
    from dynamicrafter_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    inner = H * 64
    qkv = _randn((B * N, 3 * inner), 8).to(torch.bfloat16)
    out = ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], N, 1, 1.0)], B, H, N, 0.125)
    f = qkv.float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = torch.empty((B, H, N, 64), device="cuda")
    for b in range(B):                                           # one (batch) at a time: N x N fp32 scores per head
        ref[b] = torch.softmax(f[0][b] @ f[1][b].transpose(-1, -2) * 0.125, dim=-1) @ f[2][b]
    err = _rel(out, ref.permute(0, 2, 1, 3).reshape(B * N, inner))
    print(f"\nflash attention B={B} H={H} N={N}: rel-L2 vs torch fp32 {err:.3e}")
    assert err < 1e-2


def test_vae_attn_block_c4_shape_vs_torch():
    """ae_modules.py:53-78 at the 1024 config's latent (72x128 = 9216 tokens, single head, d = 512), two frames."""
    from dynamicrafter_b200.lvdm.modules.networks.ae_modules import AttnBlock
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    C, hh, ww, nf = 512, 72, 128, 2
    blk = AttnBlock(C)
    sd = synth_state_dict({k: tuple(v.shape) for k, v in blk.state_dict().items()}, 17)
    blk.load_state_dict(sd)
    blk = blk.cuda().eval()
    x = _randn((nf, C, hh, ww), 18, 1.5, 0.2)
    y = blk(x)
    # fp32 torch restatement of the reference forward, frame by frame
    ref = torch.empty_like(x)
    w = {k: v.cuda() for k, v in sd.items()}
    for f in range(nf):
        hn = F.group_norm(x[f:f + 1], 32, w["norm.weight"], w["norm.bias"], 1e-6)
        q = F.conv2d(hn, w["q.weight"], w["q.bias"]).reshape(C, hh * ww).t()
        k = F.conv2d(hn, w["k.weight"], w["k.bias"]).reshape(C, hh * ww)
        v = F.conv2d(hn, w["v.weight"], w["v.bias"]).reshape(C, hh * ww)
        p = torch.softmax((q @ k) * (C ** -0.5), dim=1)
        o = (v @ p.t()).reshape(1, C, hh, ww)
        ref[f:f + 1] = x[f:f + 1] + F.conv2d(o, w["proj_out.weight"], w["proj_out.bias"])
    err = _rel(y, ref)
    # the block's own contribution (y - x) is what the kernels compute: bar on that, not on the identity path
    err_delta = _rel(y - x, ref - x)
    print(f"\nVAE AttnBlock 72x128, d=512: rel-L2 {err:.3e} (attention branch alone {err_delta:.3e})")
    assert err < 5e-3 and err_delta < 2e-2


def test_group_norm_c4_rows():
    """5-D GroupNorm instance of the 1024 config: S = 16 * 72 * 128 = 147 456 rows, C = 320, fp32 and bf16 sources."""
    from dynamicrafter_b200 import ops
    G, S, C = 2, 147456, 320
    x = _randn((G * S, C), 21, 2.0, 0.7)
    gamma, beta = _randn((C,), 3, 0.2, 1.0), _randn((C,), 4, 0.1)
    ref = F.silu(F.group_norm(x.reshape(G, S, C).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(G * S, C)
    y = ops.group_norm(x, G, S, gamma, beta, 1e-5, silu=True)
    assert _rel(y, ref) < 4e-3
    xb = x.to(torch.bfloat16)
    refb = F.silu(F.group_norm(xb.float().reshape(G, S, C).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(G * S, C)
    assert _rel(ops.group_norm(xb, G, S, gamma, beta, 1e-5, silu=True), refb) < 4e-3
