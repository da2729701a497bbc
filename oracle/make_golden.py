#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py            # writes tests/golden/*.pt, *.json

The reference has no tests of its own, so these fixtures are what pins
the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Weights and inputs
come from dynamicrafter_b200.synth (keyed, order-independent generators) so that the GPU box
can regenerate bit-identical weights/inputs without the reference; only OUTPUTS are stored.

Shims: pytorch_lightning and omegaconf are absent -> stub modules;
cond-stage encoders -> torch.nn.Identity; DDIMSampler.register_buffer hard-codes CUDA
(ddim.py:18-22) -> plain setattr on CPU.
"""
import gzip
import json
import os
import sys
import types

import torch
import yaml

REF = os.environ.get("DCB_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from dynamicrafter_b200.synth import synth_state_dict, synth_input  # noqa: E402


# ----------------------------------------------------------------------------- shims
def install_shims():
    import torch.nn as nn
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.seed_everything = lambda s: torch.manual_seed(s)
    util = types.ModuleType("pytorch_lightning.utilities")
    util.rank_zero_only = lambda f: f
    util.rank_zero_info = print
    pl.utilities = util
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = util


class AttrDict(dict):
    """Stand-in for OmegaConf nodes (ddpm3d.py:81 reads unet_config.params.temporal_length)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(o):
    if isinstance(o, dict):
        return AttrDict({k: to_attr(v) for k, v in o.items()})
    if isinstance(o, list):
        return [to_attr(v) for v in o]
    return o


# ----------------------------------------------------------------------------- configs
TINY_UNET = dict(in_channels=8, out_channels=4, model_channels=64, attention_resolutions=[4, 2, 1],
                 num_res_blocks=2, channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64,
                 transformer_depth=1, context_dim=1024, use_linear=True, use_checkpoint=False,
                 temporal_conv=True, temporal_attention=True, temporal_selfatt_only=True,
                 use_relative_position=False, use_causal_attention=False, temporal_length=16,
                 addition_attention=True, image_cross_attention=True, default_fs=24, fs_condition=True)
TINY_UNET_256 = dict(TINY_UNET, image_cross_attention_scale_learnable=True, default_fs=3)
TINY_VAE = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64,
                ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def shapes_of(m):
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def build_ref_unet(cfg, seed):
    from lvdm.modules.networks.openaimodel3d import UNetModel
    m = UNetModel(**cfg).eval()
    m.load_state_dict(synth_state_dict(shapes_of(m), seed), strict=True)
    return m


def unet_inputs(cfg, b, h, w, seed, t_val, with_fs=True):
    T = cfg["temporal_length"]
    x = synth_input("x", (b, cfg["in_channels"], T, h, w), seed)
    ctx = synth_input("context", (b, 77 + 16 * T, cfg["context_dim"]), seed)
    ts = torch.tensor([t_val - 7 * i for i in range(b)], dtype=torch.long)
    fs = torch.tensor([cfg["default_fs"] + 2 * i for i in range(b)], dtype=torch.long) if with_fs else None
    return x, ts, ctx, fs


def golden_unet():
    cases = [("unet_tiny512", TINY_UNET, 1, 8, 16, 3, 601, True),
             ("unet_tiny512_b2", TINY_UNET, 2, 8, 8, 4, 981, False),
             ("unet_tiny256", TINY_UNET_256, 1, 8, 8, 5, 301, True)]
    for name, cfg, b, h, w, seed, t_val, with_fs in cases:
        m = build_ref_unet(cfg, seed)
        x, ts, ctx, fs = unet_inputs(cfg, b, h, w, seed, t_val, with_fs)
        with torch.no_grad():
            # extra kwargs are swallowed by the reference (openaimodel3d.py:548)
            y = m(x, ts, context=ctx, fs=fs, cfg_img=None)
            # module-boundary goldens (ResBlock / SpatialTransformer / TemporalTransformer)
            T = cfg["temporal_length"]
            emb = synth_input("emb", (b * T, 4 * cfg["model_channels"]), seed)
            hx = synth_input("h", (b * T, cfg["model_channels"], h, w), seed)
            blk = m.input_blocks[1]
            ctxf = synth_input("ctx_frames", (b * T, 77 + 16, cfg["context_dim"]), seed)
            r = blk[0](hx, emb, batch_size=b)
            s = blk[1](hx, ctxf)
            tt = blk[2](hx.reshape(b, T, -1, h, w).permute(0, 2, 1, 3, 4))
        torch.save(dict(cfg=cfg, b=b, h=h, w=w, seed=seed, t_val=t_val, with_fs=with_fs,
                        y=y.clone(), shapes=shapes_of(m),
                        res=r[:, ::4].clone(), st=s[:, ::4].clone(), tt=tt[:, ::4].clone(),   # every 4th channel
                        x_sum=float(x.double().sum()), ctx_sum=float(ctx.double().sum())),
                   os.path.join(OUT, name + ".pt"))
        print(name, "y std", float(y.std()), "absmax", float(y.abs().max()))


def golden_vae():
    from lvdm.modules.networks.ae_modules import Decoder
    dec = Decoder(**TINY_VAE).eval()
    pq = torch.nn.Conv2d(4, TINY_VAE["z_channels"], 1)          # autoencoder.py:35
    shapes = {"decoder." + k: v for k, v in shapes_of(dec).items()}
    shapes.update({"post_quant_conv." + k: v for k, v in shapes_of(pq).items()})
    sd = synth_state_dict(shapes, 11)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})
    pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in sd.items() if k.startswith("post_quant")})
    z = synth_input("z", (2, 4, 8, 12), 11)
    with torch.no_grad():
        y = dec(pq(z))                                           # autoencoder.py:104-107
    # encode side: Encoder + quant_conv -> moments of the posterior (autoencoder.py:97-102)
    from lvdm.modules.networks.ae_modules import Encoder
    enc = Encoder(**TINY_VAE).eval()
    qc = torch.nn.Conv2d(2 * TINY_VAE["z_channels"], 2 * 4, 1)  # autoencoder.py:34
    eshapes = {"encoder." + k: v for k, v in shapes_of(enc).items()}
    eshapes.update({"quant_conv." + k: v for k, v in shapes_of(qc).items()})
    esd = synth_state_dict(eshapes, 11)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in esd.items() if k.startswith("encoder.")})
    qc.load_state_dict({k[len("quant_conv."):]: v for k, v in esd.items() if k.startswith("quant_conv.")})
    img = synth_input("img", (2, 3, 64, 96), 11)
    with torch.no_grad():
        moments = qc(enc(img))
    torch.save(dict(dd=TINY_VAE, seed=11, zshape=(2, 4, 8, 12), y=y.clone(), shapes=shapes, enc_shapes=eshapes,
                    img_shape=(2, 3, 64, 96), moments=moments.clone()), os.path.join(OUT, "vae_tiny.pt"))
    print("vae_tiny y std", float(y.std()), "moments std", float(moments.std()))


def build_lvd(unet_cfg, vae_cfg, base_yaml, seed):
    """Full LatentVisualDiffusion from the reference YAML with the tiny UNet/VAE swapped in."""
    from utils.utils import instantiate_from_config
    with open(os.path.join(REF, "configs", base_yaml)) as f:
        cfg = to_attr(yaml.safe_load(f))
    mp = cfg.model.params
    mp.unet_config.params = to_attr(dict(unet_cfg))
    mp.first_stage_config.params.ddconfig = to_attr(dict(vae_cfg))
    mp.cond_stage_config = to_attr({"target": "torch.nn.Identity"})
    mp.img_cond_stage_config = to_attr({"target": "torch.nn.Identity"})
    mp.image_proj_stage_config = to_attr({"target": "torch.nn.Identity"})
    model = instantiate_from_config(cfg.model).eval()
    unet = model.model.diffusion_model
    unet.load_state_dict(synth_state_dict(shapes_of(unet), seed), strict=True)
    fsm = model.first_stage_model
    fsm.load_state_dict(synth_state_dict(shapes_of(fsm), seed), strict=True)
    return model, cfg


def golden_ddim():
    from lvdm.models.samplers.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)

    out = {}
    # (i) schedule tables for the two real settings (C1: 256 cfg, C2: 512 cfg)
    for tag, yml, S, spacing, eta in (("c1", "inference_256_v1.0.yaml", 10, "uniform", 1.0),
                                      ("c2", "inference_512_v1.0.yaml", 50, "uniform_trailing", 1.0),
                                      ("c2_eta0", "inference_512_v1.0.yaml", 50, "uniform_trailing", 0.0)):
        ucfg = TINY_UNET_256 if "256" in yml else TINY_UNET
        model, _ = build_lvd(ucfg, TINY_VAE, yml, 21)
        smp = DDIMSampler(model)
        smp.make_schedule(S, ddim_discretize=spacing, ddim_eta=eta, verbose=False)
        tab = dict(timesteps=torch.as_tensor(smp.ddim_timesteps.copy()),
                   alphas=torch.as_tensor(smp.ddim_alphas).clone(),
                   alphas_prev=torch.as_tensor(smp.ddim_alphas_prev).clone(),
                   sigmas=torch.as_tensor(smp.ddim_sigmas).clone(),
                   sqrt_one_minus_alphas=torch.as_tensor(smp.ddim_sqrt_one_minus_alphas).clone(),
                   alphas_cumprod=model.alphas_cumprod.clone(),
                   sqrt_alphas_cumprod=model.sqrt_alphas_cumprod.clone(),
                   sqrt_one_minus_alphas_cumprod=model.sqrt_one_minus_alphas_cumprod.clone())
        if model.use_dynamic_rescale:
            tab["scale"] = smp.ddim_scale_arr.clone()
            tab["scale_prev"] = smp.ddim_scale_arr_prev.clone()
            tab["scale_arr"] = model.scale_arr.clone()
        out["tables_" + tag] = tab

    # (ii) sampler with a deterministic mock denoiser (SURVEY.md section 8(c) pin (ii))
    for tag, yml, S, spacing, gr in (("mock_c2", "inference_512_v1.0.yaml", 50, "uniform_trailing", 0.7),
                                     ("mock_c1", "inference_256_v1.0.yaml", 10, "uniform", 0.0)):
        model, _ = build_lvd(TINY_UNET_256 if "256" in yml else TINY_UNET, TINY_VAE, yml, 21)
        model.apply_model = lambda x, t, c, **kw: torch.tanh(0.3 * x + c)
        for eta in (0.0, 1.0):
            smp = DDIMSampler(model)
            torch.manual_seed(0)
            samples, inter = smp.sample(S=S, batch_size=2, shape=(4, 16, 8, 8),
                                        conditioning=torch.tensor([0.1]),
                                        unconditional_conditioning=torch.tensor([-0.1]),
                                        unconditional_guidance_scale=7.5, eta=eta, verbose=False,
                                        timestep_spacing=spacing, guidance_rescale=gr)
            out[f"{tag}_eta{int(eta)}"] = dict(samples=samples.clone(),
                                               n_inter=len(inter["x_inter"]),
                                               last_pred_x0=inter["pred_x0"][-1].clone())
            print(tag, eta, float(samples.mean()), float(samples.std()))

    # (ii-b) the multi-condition sampler (ddim_multiplecond.py: cond / uncond / image-only uncond, cfg_img) with the
    # same mock denoiser
    from lvdm.models.samplers.ddim_multiplecond import DDIMSampler as DDIMSamplerMC
    DDIMSamplerMC.register_buffer = lambda self, n, a: setattr(self, n, a)
    for tag, yml, S, spacing, gr in (("mockmc_c2", "inference_512_v1.0.yaml", 50, "uniform_trailing", 0.7),
                                     ("mockmc_c1", "inference_256_v1.0.yaml", 10, "uniform", 0.0)):
        model, _ = build_lvd(TINY_UNET_256 if "256" in yml else TINY_UNET, TINY_VAE, yml, 21)
        model.apply_model = lambda x, t, c, **kw: torch.tanh(0.3 * x + c)
        for eta in (0.0, 1.0):
            smp = DDIMSamplerMC(model)
            torch.manual_seed(0)
            samples, inter = smp.sample(S=S, batch_size=2, shape=(4, 16, 8, 8),
                                        conditioning=torch.tensor([0.1]),
                                        unconditional_conditioning=torch.tensor([-0.1]),
                                        unconditional_guidance_scale=7.5, eta=eta, verbose=False,
                                        timestep_spacing=spacing, guidance_rescale=gr, cfg_img=2.0,
                                        unconditional_conditioning_img_nonetext=torch.tensor([0.05]))
            out[f"{tag}_eta{int(eta)}"] = dict(samples=samples.clone(), n_inter=len(inter["x_inter"]),
                                               last_pred_x0=inter["pred_x0"][-1].clone())
            print(tag, eta, float(samples.mean()), float(samples.std()))

    # (iii) end-to-end: real (tiny) UNet through apply_model + decode_first_stage
    for tag, yml, S, spacing, gr, ucfg in (
            ("e2e_c2", "inference_512_v1.0.yaml", 4, "uniform_trailing", 0.7, TINY_UNET),
            ("e2e_c1", "inference_256_v1.0.yaml", 4, "uniform", 0.0, TINY_UNET_256)):
        model, _ = build_lvd(ucfg, TINY_VAE, yml, 31)
        model.perframe_ae = True
        B, T, h, w = 1, 16, 8, 8
        cond = {"c_crossattn": [synth_input("cond", (B, 77 + 16 * T, 1024), 31)],
                "c_concat": [synth_input("concat", (B, 4, T, h, w), 31)]}
        uc = {"c_crossattn": [synth_input("uncond", (B, 77 + 16 * T, 1024), 31)],
              "c_concat": cond["c_concat"]}
        x_T = synth_input("x_T", (B, 4, T, h, w), 31)
        fs = torch.tensor([ucfg["default_fs"]] * B, dtype=torch.long)
        smp = DDIMSampler(model)
        torch.manual_seed(123)
        samples, inter = smp.sample(S=S, batch_size=B, shape=(4, T, h, w), conditioning=cond,
                                    unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                    eta=1.0, verbose=False, x_T=x_T, fs=fs, timestep_spacing=spacing,
                                    guidance_rescale=gr, cfg_img=None,
                                    unconditional_conditioning_img_nonetext=None)
        with torch.no_grad():
            video = model.decode_first_stage(samples)
        # the same run without noise (eta = 0), every step logged: teacher-forcing trajectory for the CUDA path
        smp0 = DDIMSampler(model)
        _, traj = smp0.sample(S=S, batch_size=B, shape=(4, T, h, w), conditioning=cond,
                              unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                              eta=0.0, verbose=False, x_T=x_T, fs=fs, timestep_spacing=spacing,
                              guidance_rescale=gr, log_every_t=1, cfg_img=None,
                              unconditional_conditioning_img_nonetext=None)
        out[tag] = dict(samples=samples.clone(), video=video.clone(), S=S, spacing=spacing, gr=gr,
                        traj_x=torch.stack(traj["x_inter"]).clone(), traj_pred_x0=torch.stack(traj["pred_x0"]).clone(),
                        unet_shapes=shapes_of(model.model.diffusion_model),
                        vae_shapes=shapes_of(model.first_stage_model),
                        ucfg=ucfg, seed=31, yaml=yml)
        print(tag, "samples std", float(samples.std()), "video std", float(video.std()))
    torch.save(out, os.path.join(OUT, "ddim.pt"))


def golden_mask():
    """Sampler mask / x0 blend path (ddim.py:174-180) with the mock denoiser: clean_cond=True (deterministic) and
    clean_cond=False (q_sample noise from the CPU generator), eta = 0; plus one stochastic_encode / decode round."""
    from lvdm.models.samplers.ddim import DDIMSampler
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)
    out = {}
    shape = (2, 4, 16, 8, 8)
    mask = (synth_input("mask", (2, 1, 16, 8, 8), 61) > 0).float()
    x0 = synth_input("x0", shape, 61)
    x_T = synth_input("x_T", shape, 61)
    for tag, yml, S, spacing, gr in (("c2", "inference_512_v1.0.yaml", 50, "uniform_trailing", 0.7),
                                     ("c1", "inference_256_v1.0.yaml", 10, "uniform", 0.0)):
        model, _ = build_lvd(TINY_UNET_256 if "256" in yml else TINY_UNET, TINY_VAE, yml, 21)
        model.apply_model = lambda x, t, c, **kw: torch.tanh(0.3 * x + c)
        for clean in (True, False):
            smp = DDIMSampler(model)
            torch.manual_seed(0)
            samples, _ = smp.sample(S=S, batch_size=2, shape=shape[1:], conditioning=torch.tensor([0.1]),
                                    unconditional_conditioning=torch.tensor([-0.1]), unconditional_guidance_scale=7.5,
                                    eta=0.0, verbose=False, timestep_spacing=spacing, guidance_rescale=gr, x_T=x_T,
                                    mask=mask, x0=x0, clean_cond=clean)
            out[f"{tag}_clean{int(clean)}"] = samples.clone()
            print("mask", tag, clean, float(samples.mean()), float(samples.std()))
        # stochastic_encode to DDIM index 30 (C1: 6) with given noise, then decode back (ddim.py:282-319)
        smp = DDIMSampler(model)
        smp.make_schedule(S, ddim_discretize=spacing, ddim_eta=0.0, verbose=False)
        t_enc = 30 if S == 50 else 6
        noise = synth_input("enc_noise", shape, 61)
        z = smp.stochastic_encode(x0, torch.tensor([t_enc] * 2), noise=noise)
        dec = smp.decode(z, torch.tensor([0.1]), t_enc, unconditional_guidance_scale=7.5,
                         unconditional_conditioning=torch.tensor([-0.1]))
        out[f"{tag}_enc"] = z.clone()
        out[f"{tag}_dec"] = dec.clone()
        out[f"{tag}_t_enc"] = t_enc
    torch.save(out, os.path.join(OUT, "ddim_mask.pt"))


TINY_RESAMPLER = dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=4, embedding_dim=192, output_dim=256,
                      ff_mult=4, video_length=4)


def golden_resampler():
    """image_proj_model: the reference Resampler on synthetic CLIP tokens."""
    from lvdm.modules.encoders.resampler import Resampler
    m = Resampler(**TINY_RESAMPLER).eval()
    shapes = shapes_of(m)
    m.load_state_dict(synth_state_dict(shapes, 51), strict=True)
    x = synth_input("clip_tokens", (2, 37, TINY_RESAMPLER["embedding_dim"]), 51)
    with torch.no_grad():
        y = m(x)
    torch.save(dict(cfg=TINY_RESAMPLER, seed=51, xshape=tuple(x.shape), shapes=shapes, y=y.clone()),
               os.path.join(OUT, "resampler_tiny.pt"))
    print("resampler_tiny y std", float(y.std()), tuple(y.shape))


def golden_sds():
    """SDS guidance core: the unmodified DynamiCrafterGuidancePipeline._sds_loss / _add_noise /
    _sample_timestep (guidance_pipeline.py:273-424) on a deterministic mock denoiser.  The module imports omegaconf,
    matplotlib and (through scripts.evaluation.funcs) decord at top level: absent here -> empty stub modules; the object
    is created without __init__ (which would download a checkpoint) and given exactly the attributes the methods read."""
    for name in ("omegaconf", "matplotlib", "matplotlib.pyplot", "decord"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["omegaconf"].OmegaConf = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["decord"].VideoReader = object
    sys.modules["decord"].cpu = lambda *a, **k: None
    import torch.nn.functional as F
    import guidance_pipeline as gp
    out = {}
    shape = (2, 4, 16, 8, 8)
    for tag, yml, res in (("c2", "inference_512_v1.0.yaml", (320, 512)), ("c1", "inference_256_v1.0.yaml", (256, 256))):
        model, _ = build_lvd(TINY_UNET_256 if "256" in yml else TINY_UNET, TINY_VAE, yml, 21)
        model.apply_model = lambda x, t, c, **kw: torch.tanh(0.3 * x + c)
        p = object.__new__(gp.DynamiCrafterGuidancePipeline)
        p.device, p.resolution, p.model, p.debug_enabled = torch.device("cpu"), res, model, False
        torch.manual_seed(5)
        ts = [p._sample_timestep(2) for _ in range(3)]
        out[f"{tag}_timesteps_seed5"] = torch.stack(ts)
        for wt in ("t", "ada", "uniform"):
            latents = synth_input("sds_latents_" + wt, shape, 71).requires_grad_(True)
            t_fixed = torch.tensor([439, 79] if tag == "c2" else [501, 101], dtype=torch.long)
            p._sample_timestep = lambda b, lo=0.02, hi=0.98, _t=t_fixed: _t
            rec = {}
            real_add, real_mse = p._add_noise, F.mse_loss

            def add_noise(x, n, t, _r=rec, _f=real_add):
                _r["noise"], _r["t"] = n.clone(), t.clone()
                return _f(x, n, t)

            def mse(a, b, reduction="mean", _r=rec, _f=real_mse):
                _r["target"] = b.clone()
                return _f(a, b, reduction=reduction)

            p._add_noise = add_noise
            F.mse_loss = mse
            try:
                torch.manual_seed(9)
                loss = p._sds_loss(latents, {"cond": torch.tensor([0.1]), "uc": torch.tensor([-0.1]), "fs": None},
                                   cfg_scale=7.5, weight_type=wt)
            finally:
                F.mse_loss = real_mse
                del p._add_noise
            loss.backward()
            out[f"{tag}_{wt}"] = dict(noise=rec["noise"], t=rec["t"], target=rec["target"], loss=loss.detach().clone(),
                                      dloss=latents.grad.clone())
            print("sds", tag, wt, float(loss), float(rec["target"].std()))
        out[f"{tag}_alphas_cumprod"] = model.alphas_cumprod.clone()
        out[f"{tag}_resolution"] = res
    torch.save(out, os.path.join(OUT, "sds.pt"))


def golden_keys():
    """state_dict key -> shape of the full-size 512 / 256 models (drop-in contract)."""
    from lvdm.modules.networks.openaimodel3d import UNetModel
    from lvdm.modules.networks.ae_modules import Decoder
    res = {}
    for tag, yml in (("512", "inference_512_v1.0.yaml"), ("256", "inference_256_v1.0.yaml")):
        with open(os.path.join(REF, "configs", yml)) as f:
            cfg = yaml.safe_load(f)["model"]["params"]
        with torch.device("meta"):
            u = UNetModel(**cfg["unet_config"]["params"])
            d = Decoder(**cfg["first_stage_config"]["params"]["ddconfig"])
        from lvdm.models.autoencoder import AutoencoderKL
        with torch.device("meta"):
            ae = AutoencoderKL(ddconfig=cfg["first_stage_config"]["params"]["ddconfig"],
                               lossconfig={"target": "torch.nn.Identity"}, embed_dim=4)
        res["autoencoder_" + tag] = {k: list(v.shape) for k, v in ae.state_dict().items()}
        res["unet_" + tag] = {k: list(v.shape) for k, v in u.state_dict().items()}
        res["decoder_" + tag] = {k: list(v.shape) for k, v in d.state_dict().items()}
        res["unet_cfg_" + tag] = cfg["unet_config"]["params"]
        res["ddconfig_" + tag] = cfg["first_stage_config"]["params"]["ddconfig"]
        res["model_cfg_" + tag] = {k: v for k, v in cfg.items() if not isinstance(v, dict)}
    with gzip.open(os.path.join(OUT, "state_dict_keys.json.gz"), "wt") as f:
        json.dump(res, f)
    print({k: len(v) for k, v in res.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    install_shims()
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["unet", "vae", "ddim", "keys"]
    for w in which:
        globals()["golden_" + w]()
