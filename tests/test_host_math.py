"""Host-side math of the round-2 restructurings, checked on the CPU in fp32 (no kernel involved): the weight folding
that lets the CUDA path skip a tensor or a pass must be an identity of the reference computation."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F


def test_upsample_fold_is_the_reference_upsample_conv():
    """nearest-x2 + 3x3 conv (openaimodel3d.py:98-106, ae_modules.py:123-127) == four 2x2-tap phase convs of the original
    image with the pre-summed weights of pack_conv3x3_up2 (what ops.conv3x3_up2 runs as phase GEMMs)."""
    from dynamicrafter_b200.lvdm import _packing
    from dynamicrafter_b200 import ops
    torch.manual_seed(0)
    N, C, H, W = 5, 64, 6, 7
    x = torch.randn(2, C, H, W)
    w = torch.randn(N, C, 3, 3) * 0.1
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    # emulate the packing in fp32 (the packer rounds to bf16 at the end: take the values before the cast)
    sets = (((0,), (1, 2)), ((0, 1), (2,)))
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for py in range(2):
        for px in range(2):
            acc = 0
            for a in range(2):
                for b in range(2):
                    wk = sum(w[:, :, ky, kx] for ky in sets[py][a] for kx in sets[px][b])          # [N, C]
                    dy, dx = ops._UP2_OFFS[py][a], ops._UP2_OFFS[px][b]
                    patch = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]                       # x[j + dy, i + dx], zero padded
                    acc = acc + torch.einsum("bchw,nc->bnhw", patch, wk)
            out[:, :, py::2, px::2] = acc
    assert (out - ref).abs().max().item() < 1e-4
    # and the packer's layout: [N, (a, b, c)] with c padded to 64
    w4 = _packing.pack_conv3x3_up2(w)
    assert len(w4) == 4 and w4[0].shape == (N, 4 * 64) and w4[0].dtype == torch.bfloat16
    wk00 = w[:, :, 0, 0]                                           # phase (0,0), tap (a=0, b=0) = kernel tap (0,0) alone
    assert torch.equal(w4[0][:, :C], wk00.to(torch.bfloat16))


def test_layer_norm_fold_is_layernorm_then_linear():
    """fold_layer_norm: LN(x) W^T + b == rstd * (x W'^T) - rstd * mean * rowsum(W') + b' with W' = W gamma, b' = b + W beta
    (attention.py:242-246; what the LayerNorm-folding GEMM mode evaluates in its epilogue)."""
    from dynamicrafter_b200.lvdm.modules.attention import _fold_ln_weights
    torch.manual_seed(1)
    C, N, M = 96, 40, 33
    norm = nn.LayerNorm(C)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(C) * 0.3 + 1.0)
        norm.bias.copy_(torch.randn(C) * 0.3)
    w, b = torch.randn(N, C) * C ** -0.5, torch.randn(N)
    x = torch.randn(M, C) * 2.0 + 0.7
    with torch.no_grad():
        ref = F.linear(norm(x), w, b)
        wf, bf = _fold_ln_weights(norm, w, b)
        mean = x.mean(1, keepdim=True)
        rstd = (x.var(1, unbiased=False, keepdim=True) + norm.eps).rsqrt()
        got = rstd * (x @ wf.t()) - rstd * mean * wf.sum(1)[None, :] + bf
    assert (got - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("C1,C2,ok", [(1280, 1280, True), (640, 640, True), (320, 320, True), (1280, 640, False),
                                      (640, 320, False), (64, 64, True), (96, 32, False)])
def test_concat_groupnorm_group_alignment(C1, C2, ok):
    """A group of cat([x1, x2]) must be a union of whole producer groups of ONE source for the two-source apply."""
    from dynamicrafter_b200 import ops
    assert ops._concat_groups_align(C1, C2) == ok
    if ok:
        cg = (C1 + C2) // 32
        for g in range(32):
            lo, hi = g * cg, (g + 1) * cg
            assert hi <= C1 or lo >= C1                                     # never straddles the two sources
            src_cg = (C1 if hi <= C1 else C2) // 32
            assert (lo if hi <= C1 else lo - C1) % src_cg == 0 and cg % src_cg == 0


@pytest.mark.filterwarnings("ignore")                  # 0 / 0 in the sigma formula at alpha_prev = alpha (S = 1000 ...)
@pytest.mark.parametrize("method", ["uniform", "uniform_trailing", "quad"])
def test_ddim_schedule_is_the_reference_schedule(method):
    """Host math of S2 (make_schedule): timesteps and the sigma / alpha tables, every discretisation the reference offers
    (utils_diffusion.py:56-92; its configs only use two of them), both parameterisations' beta tables, eta 0 and 1 --
    bit-exact against the unmodified reference functions (oracle/_ref) and against the oracle's restatement."""
    import numpy as np
    from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler, make_ddim_timesteps
    from oracle.ddim_oracle import DiffusionTables, ddim_tables
    ref_host = pytest.importorskip("oracle.ref_host")
    ref = None
    if ref_host.ref_root() is not None:
        ref_host.install_shims()
        import lvdm.models.utils_diffusion as ref
    for zero_snr in (False, True):
        tab = DiffusionTables(rescale_betas_zero_snr=zero_snr, parameterization="v" if zero_snr else "eps",
                              use_dynamic_rescale=zero_snr)
        for S in (1, 2, 10, 25, 50, 200, 1000):
            for eta in (0.0, 1.0):
                smp = DDIMSampler(tab)
                if method == "uniform" and S == 1000:        # steps 1..1000 index a 1000-entry table: the reference's
                    with pytest.raises(IndexError):           # own behaviour (utils_diffusion.py:59-61,82)
                        smp.make_schedule(S, ddim_discretize=method, ddim_eta=eta, verbose=False)
                    if ref is not None:
                        with pytest.raises(IndexError):
                            ref.make_ddim_sampling_parameters(tab.alphas_cumprod.cpu(), ref.make_ddim_timesteps(
                                method, S, tab.num_timesteps, verbose=False), eta, verbose=False)
                    continue
                smp.make_schedule(S, ddim_discretize=method, ddim_eta=eta, verbose=False)
                want = ddim_tables(tab, S, method, eta)
                assert np.array_equal(smp.ddim_timesteps, want["timesteps"])
                assert np.array_equal(np.asarray(smp.ddim_sigmas), np.asarray(want["sigmas"]), equal_nan=True)
                assert np.array_equal(np.asarray(smp.ddim_alphas), np.asarray(want["alphas"]))
                assert np.array_equal(np.asarray(smp.ddim_alphas_prev), np.asarray(want["alphas_prev"]))
                if ref is not None:
                    ts = ref.make_ddim_timesteps(method, S, tab.num_timesteps, verbose=False)
                    assert np.array_equal(make_ddim_timesteps(method, S, tab.num_timesteps), ts)
                    sig, al, alp = ref.make_ddim_sampling_parameters(tab.alphas_cumprod.cpu(), ts, eta, verbose=False)
                    assert np.array_equal(np.asarray(smp.ddim_sigmas), np.asarray(sig), equal_nan=True)
                    assert np.array_equal(np.asarray(smp.ddim_alphas), np.asarray(al))
                    assert np.array_equal(np.asarray(smp.ddim_alphas_prev), np.asarray(alp))


def _ref_host_or_skip():
    ref_host = pytest.importorskip("oracle.ref_host")
    if ref_host.ref_root() is None:
        pytest.skip("reference sources not staged (oracle/build_ref.py)")
    return ref_host


@pytest.mark.parametrize("name", ["256", "512", "1024"])
def test_config_tables_are_the_reference_yaml(name):
    """configs.MODEL_PARAMS restates configs/inference_{256,512,1024}_v1.0.yaml: every hyper-parameter it carries equals
    the YAML's value (UNet kwargs, VAE ddconfig, schedule, parameterisation, rescale switches)."""
    ref_host = _ref_host_or_skip()
    from dynamicrafter_b200.configs import MODEL_PARAMS
    mp = ref_host.load_config(f"inference_{name}_v1.0.yaml").model.params
    ours = MODEL_PARAMS[name]
    for k, v in ours.items():
        if k in ("unet_config", "first_stage_config"):
            continue
        if k == "parameterization" and k not in mp:  # the 256 YAML leaves the DDPM default (ddpm3d.py:58)
            assert v == "eps"
            continue
        assert k in mp and (list(mp[k]) == list(v) if isinstance(v, list) else mp[k] == v), k
    for k, v in ours["unet_config"]["params"].items():
        if k == "use_checkpoint":                    # YAML: True (training); scripts/evaluation/inference.py:322 sets False
            assert v is False
            continue
        assert mp.unet_config.params[k] == v, k
    assert set(mp.unet_config.params) == set(ours["unet_config"]["params"])
    vae = ours["first_stage_config"]["params"]
    assert mp.first_stage_config.params.embed_dim == vae["embed_dim"]
    assert dict(mp.first_stage_config.params.ddconfig) == vae["ddconfig"]


@pytest.mark.parametrize("name", ["256", "512", "1024"])
def test_diffusion_tables_are_the_reference_tables(name):
    """DiffusionHost's schedule buffers (betas, alphas_cumprod, the v-parameterisation tables, the dynamic-rescale array)
    equal those the reference's LatentVisualDiffusion registers (ddpm3d.py:135-189,520-527), bit for bit in fp32."""
    ref_host = _ref_host_or_skip()
    from oracle.make_golden import TINY_UNET, TINY_VAE
    from dynamicrafter_b200.configs import model_params
    from dynamicrafter_b200.host import DiffusionHost
    ref, _ = ref_host.build_lvd(f"inference_{name}_v1.0.yaml", unet_params=TINY_UNET, vae_ddconfig=TINY_VAE)
    mp = model_params(name)
    mp["unet_config"]["params"] = dict(TINY_UNET)
    mp.pop("first_stage_config")
    ours = DiffusionHost.from_model_params(mp)
    assert ours.num_timesteps == ref.num_timesteps and ours.parameterization == ref.parameterization
    assert bool(ours.use_dynamic_rescale) == bool(ref.use_dynamic_rescale)
    names = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"]
    if ref.use_dynamic_rescale:
        names.append("scale_arr")
    for n in names:
        a, b = getattr(ours, n), getattr(ref, n)
        assert a.dtype == b.dtype == torch.float32 and torch.equal(a.cpu(), b.cpu()), n


@pytest.mark.parametrize("name,h,w", [("256", 32, 32), ("512", 40, 64), ("1024", 72, 128)])
def test_roofline_numerators_are_the_reference_flop_counts(name, h, w):
    """bench.py's roofline divides configs.UNET_TFLOP / VAE_TFLOP by measured time.  Those constants are what torch's FLOP
    counter sees when the UNMODIFIED reference UNetModel / Decoder run one forward at the benchmarked shape on the meta
    device (conv + matmul, 2 FLOPs per MAC; B = 1, 16 frames) -- so the roofline's algorithmic work is the reference's."""
    ref_host = _ref_host_or_skip()
    ref_host.install_shims()
    from torch.utils.flop_counter import FlopCounterMode
    from lvdm.modules.networks.openaimodel3d import UNetModel
    from lvdm.modules.networks.ae_modules import Decoder
    from dynamicrafter_b200.configs import model_params, UNET_TFLOP, VAE_TFLOP
    mp = model_params(name)
    with torch.device("meta"), torch.no_grad():
        unet = UNetModel(**mp["unet_config"]["params"]).eval()
        x, ctx = torch.empty(1, 8, 16, h, w), torch.empty(1, 77 + 16 * 16, 1024)
        t = torch.empty(1, dtype=torch.long)
        with FlopCounterMode(display=False) as fc:
            unet(x, t, context=ctx, fs=t)
        assert abs(fc.get_total_flops() / 1e12 - UNET_TFLOP[name]) < 1e-3
        dec = Decoder(**mp["first_stage_config"]["params"]["ddconfig"]).eval()
        with FlopCounterMode(display=False) as fc:
            dec(torch.empty(16, 4, h, w))
        assert abs(fc.get_total_flops() / 1e12 - VAE_TFLOP[name]) < 1e-3
