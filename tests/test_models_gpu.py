"""Parity of the CUDA path (mirror classes -> C ABI -> sm_100a kernels) against the CPU oracle and the
reference-generated golden fixtures, on the seeded synthetic weights/inputs of oracle/make_golden.py.

Tolerances (BASELINE.json north_star): per-forward rel-L2 <= 2e-2 in bf16 mode (this round's only mode: bf16
tensor-core operands, fp32 accumulate / norms / softmax / residual stream); DDIM update bit-exact or 1e-6.
"""
import pytest
import torch

from conftest import load_golden, rel_l2
from dynamicrafter_b200.synth import synth_state_dict, synth_input

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
FP32_TOL = 1e-3


def _unet_case(name):
    g = load_golden(name)
    cfg, b, h, w, seed = g["cfg"], g["b"], g["h"], g["w"], g["seed"]
    T = cfg["temporal_length"]
    x = synth_input("x", (b, cfg["in_channels"], T, h, w), seed)
    ctx = synth_input("context", (b, 77 + 16 * T, cfg["context_dim"]), seed)
    ts = torch.tensor([g["t_val"] - 7 * i for i in range(b)], dtype=torch.long)
    fs = torch.tensor([cfg["default_fs"] + 2 * i for i in range(b)], dtype=torch.long) if g["with_fs"] else None
    return g, cfg, x, ts, ctx, fs


def _build_unet(cfg, shapes, seed):
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    m = UNetModel(**cfg)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    m.load_state_dict(synth_state_dict(shapes, seed), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name", ["unet_tiny512.pt", "unet_tiny512_b2.pt", "unet_tiny256.pt"])
def test_unet_forward_vs_reference_golden_and_oracle(name):
    from oracle import unet_oracle
    g, cfg, x, ts, ctx, fs = _unet_case(name)
    m = _build_unet(cfg, g["shapes"], g["seed"])
    y = m(x.cuda(), ts.cuda(), context=ctx.cuda(), fs=None if fs is None else fs.cuda(),
          cfg_img=None, unconditional_conditioning_img_nonetext=None).cpu()
    assert y.shape == g["y"].shape
    err_golden = rel_l2(y, g["y"])
    y_or = unet_oracle.unet_forward(synth_state_dict(g["shapes"], g["seed"]), cfg, x, ts, ctx, fs)
    err_oracle = rel_l2(y, y_or)
    print(f"\n{name}: rel-L2 vs reference golden {err_golden:.3e}, vs oracle {err_oracle:.3e}")
    assert err_golden < BF16_TOL and err_oracle < BF16_TOL
    # determinism: same inputs -> same bits
    y2 = m(x.cuda(), ts.cuda(), context=ctx.cuda(), fs=None if fs is None else fs.cuda()).cpu()
    assert torch.equal(y, y2)


@pytest.mark.parametrize("name", ["unet_tiny512.pt", "unet_tiny256.pt"])
def test_module_level_drop_ins(name):
    """ResBlock / SpatialTransformer / TemporalTransformer called with the reference's own signatures."""
    g, cfg, *_ = _unet_case(name)
    m = _build_unet(cfg, g["shapes"], g["seed"])
    b, h, w, seed, T, mc = g["b"], g["h"], g["w"], g["seed"], cfg["temporal_length"], cfg["model_channels"]
    emb = synth_input("emb", (b * T, 4 * mc), seed).cuda()
    hx = synth_input("h", (b * T, mc, h, w), seed).cuda()
    ctxf = synth_input("ctx_frames", (b * T, 77 + 16, cfg["context_dim"]), seed).cuda()
    blk = m.input_blocks[1]
    r = blk[0](hx, emb, batch_size=b).cpu()
    s = blk[1](hx, ctxf).cpu()
    tt = blk[2](hx.reshape(b, T, mc, h, w).permute(0, 2, 1, 3, 4)).cpu()
    errs = (rel_l2(r[:, ::4], g["res"]), rel_l2(s[:, ::4], g["st"]), rel_l2(tt[:, ::4], g["tt"]))
    print(f"\n{name}: ResBlock {errs[0]:.3e}  SpatialTransformer {errs[1]:.3e}  TemporalTransformer {errs[2]:.3e}")
    assert max(errs) < BF16_TOL
    # TimestepEmbedSequential dispatch in the reference layout
    seq = blk(hx, emb, context=ctxf, batch_size=b)
    assert seq.shape == hx.shape and torch.isfinite(seq).all()


def test_vae_decode_vs_reference_golden():
    from dynamicrafter_b200.lvdm.models.autoencoder import AutoencoderKL
    g = load_golden("vae_tiny.pt")
    ae = AutoencoderKL(ddconfig=g["dd"], embed_dim=4)
    sd = ae.state_dict()
    sd.update(synth_state_dict(g["shapes"], g["seed"]))
    ae.load_state_dict(sd)
    ae = ae.cuda().eval()
    z = synth_input("z", g["zshape"], g["seed"]).cuda()
    y = ae.decode(z).cpu()
    err = rel_l2(y, g["y"])
    mse = float(((y - g["y"]) ** 2).mean())
    psnr = 10 * torch.log10(torch.tensor(4.0 / mse)).item()        # peak-to-peak range 2.0 ([-1, 1] images)
    print(f"\nvae_tiny: rel-L2 {err:.3e}, PSNR {psnr:.1f} dB")
    assert err < BF16_TOL and psnr > 45.0


def test_vae_encode_vs_reference_golden():
    """AutoencoderKL.encode (Encoder with the pad-(0,1,0,1) stride-2 Downsample + quant_conv) against
    the reference's moments; the posterior object mirrors lvdm/distributions.py."""
    from dynamicrafter_b200.lvdm.models.autoencoder import AutoencoderKL
    g = load_golden("vae_tiny.pt")
    ae = AutoencoderKL(ddconfig=g["dd"], embed_dim=4)
    full = {k: tuple(v.shape) for k, v in ae.state_dict().items()}
    sd = synth_state_dict(full, 99)                                  # decoder side: anything
    sd.update(synth_state_dict(g["enc_shapes"], g["seed"]))          # encoder + quant_conv: the golden's weights
    ae.load_state_dict(sd)
    ae = ae.cuda().eval()
    img = synth_input("img", g["img_shape"], g["seed"])
    post = ae.encode(img.cuda())
    m = post.parameters.cpu()
    err = rel_l2(m, g["moments"])
    print(f"\nvae_tiny encode: moments rel-L2 vs reference golden {err:.3e}")
    assert m.shape == g["moments"].shape and err < BF16_TOL
    assert torch.equal(post.mode(), post.mean) and post.sample(torch.zeros_like(post.mean).cpu()).shape == post.mean.shape
    # Encoder called directly with the reference's signature
    h = ae.encoder(img.cuda())
    assert h.shape == (2, 2 * g["dd"]["z_channels"], 8, 12)


def test_resampler_vs_reference_golden():
    """drop-in Resampler (image_proj_model) -- same state_dict keys as the reference, output within the
    bf16 bar of the reference golden."""
    from dynamicrafter_b200.lvdm.modules.encoders.resampler import Resampler
    g = load_golden("resampler_tiny.pt")
    m = Resampler(**g["cfg"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v) for k, v in g["shapes"].items()}
    m.load_state_dict(synth_state_dict(g["shapes"], g["seed"]), strict=True)
    m = m.cuda().eval()
    x = synth_input("clip_tokens", g["xshape"], g["seed"])
    y = m(x.cuda()).cpu()
    err = rel_l2(y, g["y"])
    print(f"\nresampler_tiny: rel-L2 vs reference golden {err:.3e}")
    assert y.shape == g["y"].shape and err < BF16_TOL


def test_vae_decode_full_size_vs_oracle():
    """The real decoder (ch 128, mult 1-2-4-4, mid attention over 2560 tokens) on one 40x64 latent frame of the 512
    config against the CPU oracle: the tiny fixture above never reaches these widths / the 2560-token AttnBlock."""
    from oracle import vae_oracle
    from dynamicrafter_b200.configs import model_params
    from dynamicrafter_b200.lvdm.models.autoencoder import AutoencoderKL
    fsc = model_params("512")["first_stage_config"]["params"]
    ae = AutoencoderKL(ddconfig=fsc["ddconfig"], embed_dim=fsc["embed_dim"])
    sd = synth_state_dict({k: tuple(v.shape) for k, v in ae.state_dict().items()}, 11)
    ae.load_state_dict(sd)
    ae = ae.cuda().eval()
    z = synth_input("z", (1, 4, 40, 64), 11) * 4.0
    ref = vae_oracle.decode(sd, fsc["ddconfig"], z)
    y = ae.decode(z.cuda()).cpu()
    mse = float(((y - ref) ** 2).mean())
    psnr = 10 * torch.log10(torch.tensor(4.0 / mse)).item()
    print(f"\nfull-size VAE, 320x512 frame: rel-L2 {rel_l2(y, ref):.3e}, PSNR {psnr:.1f} dB")
    assert y.shape == (1, 3, 320, 512) and rel_l2(y, ref) < BF16_TOL and psnr > 45.0


class _MockDiffusion:
    """Duck-typed stand-in for LatentVisualDiffusion (what DDIMSampler reads)."""

    def __init__(self, tables, apply_model, device):
        self.num_timesteps = tables.num_timesteps
        self.parameterization = tables.parameterization
        self.use_dynamic_rescale = tables.use_dynamic_rescale
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            setattr(self, k, getattr(tables, k).to(device))
        if tables.use_dynamic_rescale:
            self.scale_arr = tables.scale_arr.to(device)
        self.apply_model = apply_model
        self.device = device


SCHED = {
    "c1": dict(kw=dict(parameterization="eps"), S=10, spacing="uniform"),
    "c2": dict(kw=dict(parameterization="v", rescale_betas_zero_snr=True, use_dynamic_rescale=True, base_scale=0.7),
               S=50, spacing="uniform_trailing"),
}


@pytest.mark.parametrize("tag", ["c1", "c2"])
def test_sampler_schedule_tables_bit_exact(tag):
    from oracle import ddim_oracle
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    g = load_golden("ddim.pt")["tables_" + tag]
    tab = ddim_oracle.DiffusionTables(**SCHED[tag]["kw"])
    smp = DDIMSampler(_MockDiffusion(tab, None, torch.device("cuda")))
    smp.make_schedule(SCHED[tag]["S"], ddim_discretize=SCHED[tag]["spacing"], ddim_eta=1.0, verbose=False)
    assert (smp.ddim_timesteps == g["timesteps"].numpy()).all()
    for k in ("alphas", "alphas_prev", "sigmas", "sqrt_one_minus_alphas"):
        got = torch.as_tensor(getattr(smp, "ddim_" + k))
        assert got.dtype == g[k].dtype and torch.equal(got.cpu(), g[k]), k


@pytest.mark.parametrize("tag,eta", [("mock_c2", 0.0), ("mock_c1", 0.0), ("mock_c2", 1.0)])
def test_sampler_with_mock_denoiser_teacher_forced(tag, eta):
    """Per-step parity with the oracle: feed the oracle's x_t and noise into p_sample_ddim each step."""
    from oracle import ddim_oracle
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    c = SCHED["c2" if tag.endswith("c2") else "c1"]
    gr = 0.7 if tag.endswith("c2") else 0.0
    tab = ddim_oracle.DiffusionTables(**c["kw"])
    fn = lambda x, t, cnd, **kw: torch.tanh(0.3 * x + cnd)
    dev = torch.device("cuda")
    smp = DDIMSampler(_MockDiffusion(tab, fn, dev))
    smp.batch_cfg = False
    smp.make_schedule(c["S"], ddim_discretize=c["spacing"], ddim_eta=eta, verbose=False)
    dt = ddim_oracle.ddim_tables(tab, c["S"], c["spacing"], eta)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn((2, 4, 16, 8, 8), generator=gen)
    worst = 0.0
    import numpy as np
    rng_state = None
    for i, step in enumerate(np.flip(dt["timesteps"])):
        index = c["S"] - i - 1
        ts = torch.full((2,), int(step), dtype=torch.long)
        noise = torch.randn(x.shape, generator=gen)
        xo, p0 = ddim_oracle.p_sample_ddim(tab, dt, lambda a, b_, cc: fn(a, b_, cc), x, torch.tensor([0.1]), ts, index,
                                           ucg_scale=7.5, uncond=torch.tensor([-0.1]), guidance_rescale=gr, noise=noise)
        # our sampler draws its own noise; sigma*noise is removed and the oracle's added to compare the rest
        smp._step_hint = int(step)
        coeffs = smp._coeffs(index, int(step), 7.5, gr, 1.0)
        from dynamicrafter_b200 import ops
        xc = x.cuda()
        e_c, e_u = fn(xc, None, torch.tensor([0.1], device=dev)), fn(xc, None, torch.tensor([-0.1], device=dev))
        xg, p0g = ops.ddim_update(xc, e_c.contiguous(), e_u.contiguous(), noise.cuda(), coeffs)
        worst = max(worst, rel_l2(xg.cpu(), xo), rel_l2(p0g.cpu(), p0))
        x = xo
    print(f"\n{tag} eta={eta}: worst per-step rel-L2 {worst:.3e}")
    assert worst < 2e-6     # tanh on GPU vs CPU differs by ~1 ulp; the update itself adds no error


@pytest.mark.parametrize("tag", ["mockmc_c2", "mockmc_c1"])
def test_multicond_sampler_vs_reference_golden(tag):
    """drop-in for ddim_multiplecond.DDIMSampler (three guidance branches as one batch, fused
    3-term combine + update) against the reference run with the mock denoiser, eta = 0, free running over all steps."""
    from oracle import ddim_oracle
    from dynamicrafter_b200.lvdm.models.samplers.ddim_multiplecond import DDIMSampler
    g = load_golden("ddim.pt")[tag + "_eta0"]
    c = SCHED["c2" if tag.endswith("c2") else "c1"]
    gr = 0.7 if tag.endswith("c2") else 0.0
    tab = ddim_oracle.DiffusionTables(**c["kw"])
    dev = torch.device("cuda")
    smp = DDIMSampler(_MockDiffusion(tab, lambda x, t, cnd, **kw: torch.tanh(0.3 * x + cnd.reshape(-1, 1, 1, 1, 1)), dev))
    torch.manual_seed(0)
    x_T = torch.randn((2, 4, 16, 8, 8))                       # the reference drew x_T first from the same CPU stream
    one = lambda v: torch.full((2,), v, device=dev)           # per-sample conditioning so the branches batch
    samples, inter = smp.sample(S=c["S"], batch_size=2, shape=(4, 16, 8, 8), conditioning=one(0.1),
                                unconditional_conditioning=one(-0.1), unconditional_guidance_scale=7.5, eta=0.0,
                                verbose=False, x_T=x_T.cuda(), timestep_spacing=c["spacing"], guidance_rescale=gr,
                                cfg_img=2.0, unconditional_conditioning_img_nonetext=one(0.05))
    err = rel_l2(samples.cpu(), g["samples"])
    print(f"\n{tag}: {c['S']}-step multi-condition sampler vs reference golden rel-L2 {err:.3e}")
    assert len(inter["x_inter"]) == g["n_inter"]
    assert err < 1e-4      # tanh differs by ~1 ulp between CPU and GPU; errors compound over up to 50 steps


@pytest.mark.parametrize("tag,mode,tol_x,tol_p", [("e2e_c2", "bf16", 0.1, 0.15), ("e2e_c1", "bf16", 0.1, 0.15),
                                                 ("e2e_c2", "fp32", 1e-3, 1e-3), ("e2e_c1", "fp32", 1e-3, 1e-3)])
def test_real_unet_teacher_forced_against_reference_trajectory(tag, mode, tol_x, tol_p):
    from dynamicrafter_b200.precision import use_precision
    with use_precision(mode):
        _teacher_forced(tag, tol_x, tol_p)


def _teacher_forced(tag, tol_x, tol_p):
    """per-step parity with teacher forcing -- feed the REFERENCE's x_t of every step of its eta = 0
    run (real tiny UNet, CFG 7.5, guidance rescale) into p_sample_ddim and compare pred_x0 / x_{t-1}."""
    from oracle import ddim_oracle
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    import numpy as np
    g = load_golden("ddim.pt")[tag]
    c = SCHED["c2" if tag.endswith("c2") else "c1"]
    tab = ddim_oracle.DiffusionTables(**c["kw"])
    ucfg, seed = g["ucfg"], g["seed"]
    unet = _build_unet(ucfg, g["unet_shapes"], seed)
    dev = torch.device("cuda")
    B, T, h, w = 1, 16, 8, 8
    cond = {"c_crossattn": [synth_input("cond", (B, 77 + 16 * T, 1024), seed).cuda()],
            "c_concat": [synth_input("concat", (B, 4, T, h, w), seed).cuda()]}
    uc = {"c_crossattn": [synth_input("uncond", (B, 77 + 16 * T, 1024), seed).cuda()], "c_concat": cond["c_concat"]}
    fs = torch.tensor([ucfg["default_fs"]] * B, dtype=torch.long, device=dev)

    def apply_model(x, t, cnd, **kw):
        return unet(torch.cat([x] + cnd["c_concat"], dim=1), t, context=torch.cat(cnd["c_crossattn"], 1), **kw)

    smp = DDIMSampler(_MockDiffusion(tab, apply_model, dev))
    smp.make_schedule(g["S"], ddim_discretize=g["spacing"], ddim_eta=0.0, verbose=False)
    steps = np.flip(smp.ddim_timesteps)
    assert g["traj_x"].shape[0] == g["S"] + 1
    worst_x, worst_p = 0.0, 0.0
    for i, step in enumerate(steps):
        index = g["S"] - i - 1
        ts = torch.full((B,), int(step), device=dev, dtype=torch.long)
        xp, p0 = smp.p_sample_ddim(g["traj_x"][i].cuda(), cond, ts, index=index, unconditional_guidance_scale=7.5,
                                   unconditional_conditioning=uc, fs=fs, guidance_rescale=g["gr"])
        worst_x = max(worst_x, rel_l2(xp.cpu(), g["traj_x"][i + 1]))
        worst_p = max(worst_p, rel_l2(p0.cpu(), g["traj_pred_x0"][i + 1]))
    print(f"\n{tag}: teacher-forced worst per-step rel-L2  x_prev {worst_x:.3e}  pred_x0 {worst_p:.3e}")
    assert worst_x < tol_x and worst_p < tol_p   # CFG 7.5 amplifies the per-forward error (bf16 bar 2e-2) ~5x


@pytest.mark.parametrize("tag,mode,tol", [("e2e_c2", "bf16", 0.1), ("e2e_c1", "bf16", 0.1), ("e2e_c2", "fp32", 5e-3)])
def test_end_to_end_sample_and_decode(tag, mode, tol):
    """DDIMSampler.sample over the real (tiny) UNet + VAE decode vs the reference golden, free running; in bf16 mode
    and in the fp32 parity mode (UNet in fp32 mode, VAE decode always bf16)."""
    from dynamicrafter_b200.precision import use_precision
    with use_precision(mode):
        _end_to_end(tag, tol)


def _end_to_end(tag, tol):
    from oracle import ddim_oracle
    from oracle.make_golden import TINY_VAE
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler
    from dynamicrafter_b200.lvdm.models.autoencoder import AutoencoderKL
    g = load_golden("ddim.pt")[tag]
    c = SCHED["c2" if tag.endswith("c2") else "c1"]
    tab = ddim_oracle.DiffusionTables(**c["kw"])
    ucfg, seed = g["ucfg"], g["seed"]
    unet = _build_unet(ucfg, g["unet_shapes"], seed)
    dev = torch.device("cuda")
    B, T, h, w = 1, 16, 8, 8
    cond = {"c_crossattn": [synth_input("cond", (B, 77 + 16 * T, 1024), seed).cuda()],
            "c_concat": [synth_input("concat", (B, 4, T, h, w), seed).cuda()]}
    uc = {"c_crossattn": [synth_input("uncond", (B, 77 + 16 * T, 1024), seed).cuda()], "c_concat": cond["c_concat"]}
    x_T = synth_input("x_T", (B, 4, T, h, w), seed).cuda()
    fs = torch.tensor([ucfg["default_fs"]] * B, dtype=torch.long, device=dev)

    def apply_model(x, t, cnd, **kw):           # ddpm3d.py:723-733 + DiffusionWrapper hybrid :1254-1258
        xc = torch.cat([x] + cnd["c_concat"], dim=1)
        return unet(xc, t, context=torch.cat(cnd["c_crossattn"], 1), **kw)

    smp = DDIMSampler(_MockDiffusion(tab, apply_model, dev))
    # eta = 0 variant of the golden run is not stored; with eta=1 the CUDA Philox noise differs from the CPU
    # generator the golden used, so compare the deterministic part: run with the golden's settings but eta=0
    # against the oracle run with eta=0 on the same x_T.
    usd = synth_state_dict(g["unet_shapes"], seed)
    from oracle import unet_oracle, vae_oracle
    concat_cpu = synth_input("concat", (B, 4, T, h, w), seed)
    fs_cpu = fs.cpu()

    def apply_cpu(x, t, ctx):
        return unet_oracle.unet_forward(usd, ucfg, torch.cat([x, concat_cpu], 1), t, ctx, fs_cpu)

    ref, _ = ddim_oracle.sample(tab, apply_cpu, g["S"], (B, 4, T, h, w), synth_input("cond", (B, 77 + 16 * T, 1024), seed),
                                uncond=synth_input("uncond", (B, 77 + 16 * T, 1024), seed), ucg_scale=7.5, eta=0.0,
                                timestep_spacing=g["spacing"], guidance_rescale=g["gr"], x_T=x_T.cpu())
    samples, inter = smp.sample(S=g["S"], batch_size=B, shape=(4, T, h, w), conditioning=cond,
                                unconditional_guidance_scale=7.5, unconditional_conditioning=uc, eta=0.0,
                                verbose=False, x_T=x_T, fs=fs, timestep_spacing=g["spacing"],
                                guidance_rescale=g["gr"], cfg_img=None, unconditional_conditioning_img_nonetext=None)
    err = rel_l2(samples.cpu(), ref)
    print(f"\n{tag}: free-running {g['S']}-step latents rel-L2 vs oracle {err:.3e}")
    assert err < tol      # CFG 7.5 amplifies the per-forward bf16 error ~5x per step
    # decode the reference's latents with our VAE and compare with the reference's video
    ae = AutoencoderKL(ddconfig=TINY_VAE, embed_dim=4)
    ae.load_state_dict(synth_state_dict(g["vae_shapes"], seed))
    ae = ae.cuda().eval()
    z = g["samples"].permute(0, 2, 1, 3, 4).reshape(B * T, 4, h, w).cuda() * (1.0 / 0.18215)
    video = ae.decode(z).cpu().reshape(B, T, 3, 8 * h, 8 * w).permute(0, 2, 1, 3, 4)
    mse = float(((video - g["video"]) ** 2).mean())
    psnr = 10 * torch.log10(torch.tensor(4.0 / mse)).item()
    print(f"{tag}: decoded frames rel-L2 {rel_l2(video, g['video']):.3e}  PSNR {psnr:.1f} dB")
    assert psnr > 45.0


def test_unet_full_512_config_vs_oracle_reduced_frames():
    """The real 1.44 B-parameter 512 config (all channel widths, BN=160 GEMM tiles, 5/10/20 heads) on 2 frames of
    a 16x24 latent, against the CPU oracle (a full 16-frame 40x64 forward takes the oracle ~40 s per call)."""
    from oracle import unet_oracle
    from dynamicrafter_b200.configs import model_params
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    cfg = model_params("512")["unet_config"]["params"]
    m = UNetModel(**cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd = synth_state_dict(shapes, 3)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    b, T, h, w = 1, 2, 16, 24
    x = synth_input("x", (b, 8, T, h, w), 3)
    ctx = synth_input("context", (b, 77 + 16 * T, 1024), 3)
    ts = torch.tensor([601], dtype=torch.long)
    y = m(x.cuda(), ts.cuda(), context=ctx.cuda()).cpu()
    m.use_cuda_graph = True
    yg = m(x.cuda(), ts.cuda(), context=ctx.cuda()).cpu()
    yg2 = m(x.cuda(), ts.cuda(), context=ctx.cuda()).cpu()
    assert torch.equal(y, yg) and torch.equal(yg, yg2)          # CUDA-graph replay == eager launches
    ref = unet_oracle.unet_forward(sd, cfg, x, ts, ctx, None)
    err = rel_l2(y, ref)
    print(f"\nfull 512-config UNet (2 frames, 16x24 latent): rel-L2 vs oracle {err:.3e}")
    assert err < BF16_TOL
    # the same forward in the fp32 parity mode (weights re-packed as hi/lo splits)
    from dynamicrafter_b200.precision import use_precision
    m.use_cuda_graph = False
    with use_precision("fp32"):
        y32 = m(x.cuda(), ts.cuda(), context=ctx.cuda()).cpu()
    err32 = rel_l2(y32, ref)
    print(f"full 512-config UNet, fp32 mode: rel-L2 vs oracle {err32:.3e}")
    assert err32 < FP32_TOL


@pytest.mark.parametrize("script", ["frame_shard_check.py", "branch_shard_check.py"])
def test_sharded_single_video_modes_match_single_gpu(script):
    """frames sharded over 2 GPUs (all-to-all around the temporal blocks) and guidance branches sharded
    over 2 GPUs (predictions all-gathered every step); both need >= 2 GPUs."""
    import os, socket, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, script)],
                       capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0



@pytest.mark.parametrize("name", ["unet_tiny512.pt", "unet_tiny512_b2.pt", "unet_tiny256.pt"])
def test_unet_forward_fp32_mode(name):
    """north_star: per-step eps within rel-L2 <= 1e-3 of the fp32 reference in fp32 mode (hi/lo-split bf16 operands on
    the tensor cores, everything else fp32; csrc/fp32_mode.cu)."""
    from dynamicrafter_b200.precision import set_precision
    g, cfg, x, ts, ctx, fs = _unet_case(name)
    set_precision("fp32")
    try:
        m = _build_unet(cfg, g["shapes"], g["seed"])
        y = m(x.cuda(), ts.cuda(), context=ctx.cuda(), fs=None if fs is None else fs.cuda()).cpu()
    finally:
        set_precision("bf16")
    err = rel_l2(y, g["y"])
    print(f"\n{name}: fp32 mode rel-L2 vs reference golden {err:.3e}")
    assert err < FP32_TOL


def test_multicond_branches_batched_equals_one_by_one():
    """The three guidance branches as ONE batch-3 UNet forward vs three separate forwards (dict conditioning, real tiny
    UNet).  Run in the fp32 parity mode, where batching may only change the result at the 1e-4 level."""
    from oracle import ddim_oracle
    from dynamicrafter_b200.precision import use_precision
    from dynamicrafter_b200.lvdm.models.samplers.ddim_multiplecond import DDIMSampler
    g = load_golden("ddim.pt")["e2e_c2"]
    ucfg, seed = g["ucfg"], g["seed"]
    tab = ddim_oracle.DiffusionTables(**SCHED["c2"]["kw"])
    dev = torch.device("cuda")
    B, T, h, w = 1, 16, 8, 8
    mk = lambda name: {"c_crossattn": [synth_input(name, (B, 77 + 16 * T, 1024), seed).cuda()],
                       "c_concat": [synth_input("concat", (B, 4, T, h, w), seed).cuda()]}
    cond, uc, uc_img = mk("cond"), mk("uncond"), mk("uncond_img")
    x = synth_input("x_T", (B, 4, T, h, w), seed).cuda()
    fs = torch.tensor([ucfg["default_fs"]] * B, dtype=torch.long, device=dev)
    with use_precision("fp32"):
        unet = _build_unet(ucfg, g["unet_shapes"], seed)

        def apply_model(xx, t, cnd, **kw):
            return unet(torch.cat([xx] + cnd["c_concat"], dim=1), t, context=torch.cat(cnd["c_crossattn"], 1), **kw)

        smp = DDIMSampler(_MockDiffusion(tab, apply_model, dev))
        smp.make_schedule(4, ddim_discretize="uniform_trailing", ddim_eta=0.0, verbose=False)
        ts = torch.full((B,), int(smp.ddim_timesteps[-1]), device=dev, dtype=torch.long)
        outs = []
        for batched in (True, False):
            smp.batch_cfg = batched
            outs.append(smp.p_sample_ddim(x, cond, ts, index=3, unconditional_guidance_scale=7.5,
                                          unconditional_conditioning=uc, fs=fs, guidance_rescale=0.7, cfg_img=2.0,
                                          unconditional_conditioning_img_nonetext=uc_img))
    err = max(rel_l2(outs[0][0].cpu(), outs[1][0].cpu()), rel_l2(outs[0][1].cpu(), outs[1][1].cpu()))
    print(f"\nmulti-condition step, batch-3 forward vs 3 forwards (fp32 mode): rel-L2 {err:.3e}")
    assert err < 5e-4       # fp32-mode forwards agree to ~2e-5; the 3-term guidance (7.5 / 2.0) amplifies that ~5x


def test_cross_kv_cache_follows_the_context_tensor():
    """Cross-attention K/V are projected once per context TENSOR (same object, same version) and reused over the steps
    (attention.py:86-94 recomputes them every forward).  A new tensor, or an in-place change of the cached one, must
    re-project: compare with a model that never caches (eager and CUDA-graph paths)."""
    from dynamicrafter_b200.lvdm import _packing
    from dynamicrafter_b200.lvdm.modules.networks.openaimodel3d import UNetModel
    g = load_golden("unet_tiny512.pt")
    cfg, h, w, seed = g["cfg"], g["h"], g["w"], g["seed"]
    T = cfg["temporal_length"]
    m = UNetModel(**cfg)
    m.load_state_dict(synth_state_dict(g["shapes"], seed), strict=True)
    m = m.cuda().eval()
    x = synth_input("x", (1, cfg["in_channels"], T, h, w), seed).cuda()
    ts = torch.tensor([g["t_val"]], dtype=torch.long).cuda()
    ctx_a = synth_input("context", (1, 77 + 16 * T, cfg["context_dim"]), seed).cuda()
    ctx_b = synth_input("context", (1, 77 + 16 * T, cfg["context_dim"]), seed + 1).cuda()

    def fresh(ctx):
        old = _packing.CACHE_CROSS_KV
        _packing.CACHE_CROSS_KV = False
        try:
            m.use_cuda_graph = False
            return m(x, ts, context=ctx.clone())
        finally:
            _packing.CACHE_CROSS_KV = old

    ref_a, ref_b = fresh(ctx_a), fresh(ctx_b)
    assert not torch.equal(ref_a, ref_b)
    for graph in (False, True):
        m.use_cuda_graph = graph
        m.__dict__.pop("_ctx_cache", None)
        assert torch.equal(m(x, ts, context=ctx_a), ref_a)
        assert torch.equal(m(x, ts, context=ctx_a), ref_a)            # cache hit
        assert torch.equal(m(x, ts, context=ctx_b), ref_b)            # another tensor
        assert torch.equal(m(x, ts, context=ctx_a), ref_a)
        c = ctx_a.clone()
        assert torch.equal(m(x, ts, context=c), ref_a)
        c.copy_(ctx_b)                                                # same object, new version
        assert torch.equal(m(x, ts, context=c), ref_b)
