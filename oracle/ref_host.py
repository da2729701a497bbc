"""Build the UNMODIFIED reference host (``LatentVisualDiffusion`` + ``DDIMSampler``) from its own YAML.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference sources are imported from
``oracle/_ref`` (staged by ``oracle/build_ref.py``; travels to the GPU box) or, in the build container, straight
from ``/root/reference``.  Shims for what is absent offline: a stub ``pytorch_lightning``
(``LightningModule = nn.Module``), an attr-dict in place of OmegaConf nodes, ``torch.nn.Identity`` for the CLIP
cond stages (conditioning is synthetic), and a device-agnostic ``DDIMSampler.register_buffer`` (the reference
hard-codes ``.to("cuda")``, ddim.py:18-22).

Used for
  * ``bench.py --impl reference`` / ``cpu_baseline``: the reference's own ``DDIMSampler.p_sample_ddim`` on the host
    cores (kind = "reference");
  * ``tests/test_reference_host*.py``: the drop-in check -- the reference's ``instantiate_from_config`` builds its
    ``LatentVisualDiffusion`` with the YAML ``target:`` strings pointed at the dcb200 mirror classes, its own
    ``DDIMSampler`` / ``decode_first_stage`` then drive them.
"""
import os
import sys
import types

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")

MIRROR_TARGETS = {
    "unet": "dynamicrafter_b200.lvdm.modules.networks.openaimodel3d.UNetModel",
    "first_stage": "dynamicrafter_b200.lvdm.models.autoencoder.AutoencoderKL",
}


def ref_root():
    """Directory holding the reference's ``lvdm`` / ``utils`` / ``configs`` (None if neither copy exists)."""
    for d in (STAGED, os.environ.get("DCB_REFERENCE", "/root/reference")):
        if os.path.isdir(os.path.join(d, "lvdm")) and os.path.isdir(os.path.join(d, "utils")):
            return d
    return None


def install_shims():
    """Stub pytorch_lightning (absent offline) and put the reference on sys.path.  Idempotent."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference sources not found: run `python oracle/build_ref.py` in the build container")
    if root not in sys.path:
        sys.path.insert(0, root)
    if "pytorch_lightning" not in sys.modules:
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
    return root


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


def load_config(yaml_name):
    root = install_shims()
    with open(os.path.join(root, "configs", yaml_name)) as f:
        return to_attr(yaml.safe_load(f))


def build_lvd(yaml_name, unet_params=None, vae_ddconfig=None, mirror=False, device=None):
    """``instantiate_from_config(cfg.model)`` of the reference YAML -> the reference's LatentVisualDiffusion.

    unet_params / vae_ddconfig: optional replacement (tiny test networks).  mirror=True edits ONLY the two ``target:``
    strings the way INTEGRATION.md tells a maintainer to, so the reference host builds the dcb200 classes.
    """
    install_shims()
    from utils.utils import instantiate_from_config
    cfg = load_config(yaml_name)
    mp = cfg.model.params
    if unet_params is not None:
        mp.unet_config.params = to_attr(dict(unet_params))
    mp.unet_config.params["use_checkpoint"] = False             # as scripts/evaluation/inference.py:322
    if vae_ddconfig is not None:
        mp.first_stage_config.params.ddconfig = to_attr(dict(vae_ddconfig))
    for k in ("cond_stage_config", "img_cond_stage_config", "image_proj_stage_config"):
        mp[k] = to_attr({"target": "torch.nn.Identity"})
    if mirror:
        mp.unet_config.target = MIRROR_TARGETS["unet"]
        mp.first_stage_config.target = MIRROR_TARGETS["first_stage"]
    if device is not None and torch.device(device).type == "cuda":
        with torch.device(device):
            model = instantiate_from_config(cfg.model)
    else:
        model = instantiate_from_config(cfg.model)
    return model.eval(), cfg


def ddim_sampler_cls(multicond=False):
    """The reference's DDIMSampler with a device-agnostic register_buffer (ddim.py:18-22 forces 'cuda')."""
    install_shims()
    if multicond:
        from lvdm.models.samplers.ddim_multiplecond import DDIMSampler
    else:
        from lvdm.models.samplers.ddim import DDIMSampler

    def register_buffer(self, name, attr):
        if isinstance(attr, torch.Tensor):
            attr = attr.to(self.model.betas.device)
        setattr(self, name, attr)

    DDIMSampler.register_buffer = register_buffer
    return DDIMSampler


# ------------------------------------------------------------------------------------------ CPU baseline
SAMPLER_512 = dict(S=50, spacing="uniform_trailing", eta=1.0, cfg=7.5, guidance_rescale=0.7, fs=24)


def reference_step_runner(yaml_name="inference_512_v1.0.yaml", threads=None, seed=7, frames=16):
    """-> (step_fn, info).  step_fn() runs ONE classifier-free-guided DDIM step of the unmodified reference on the
    CPU through its own public API: DDIMSampler.p_sample_ddim (two B=1 UNetModel forwards -- cond, uncond -- plus
    its pointwise update, ddim.py:206-279) at the full workload shape (latent [1,4,16,h,w], context 77+256 tokens)."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    model, cfg = build_lvd(yaml_name)
    mp = cfg.model.params
    h, w = mp.image_size
    T = frames
    g = torch.Generator().manual_seed(seed)
    # the reference zero-initialises its output projections; timing does not depend on values, but keep the
    # activations finite and non-trivial: re-draw all-zero tensors
    with torch.no_grad():
        for p in model.model.diffusion_model.parameters():
            if p.abs().max() == 0:
                if p.dim() >= 2:
                    p.uniform_(-1, 1, generator=g).mul_(1.0 / (p[0].numel() ** 0.5))
                else:
                    p.uniform_(-0.02, 0.02, generator=g)
    x = torch.randn((1, 4, T, h, w), generator=g)
    concat = torch.randn((1, 4, T, h, w), generator=g)
    cond = {"c_crossattn": [torch.randn((1, 77 + 16 * T, 1024), generator=g)], "c_concat": [concat]}
    uc = {"c_crossattn": [torch.randn((1, 77 + 16 * T, 1024), generator=g)], "c_concat": [concat]}
    sm = SAMPLER_512
    smp = ddim_sampler_cls()(model)
    smp.make_schedule(sm["S"], ddim_discretize=sm["spacing"], ddim_eta=sm["eta"], verbose=False)
    total = len(smp.ddim_timesteps)
    fs = torch.tensor([sm["fs"]], dtype=torch.long)
    state = {"x": x, "i": 0}

    def step():
        i = state["i"] % total
        t = int(smp.ddim_timesteps[total - i - 1])
        ts = torch.full((1,), t, dtype=torch.long)
        with torch.no_grad():
            x_prev, _ = smp.p_sample_ddim(state["x"], cond, ts, index=total - i - 1,
                                          unconditional_guidance_scale=sm["cfg"], unconditional_conditioning=uc,
                                          fs=fs, guidance_rescale=sm["guidance_rescale"])
        state["x"], state["i"] = x_prev, state["i"] + 1
        return x_prev

    info = {"threads": threads, "latent": (h, w), "frames": T, "root": ref_root(),
            "api": "lvdm.models.samplers.ddim.DDIMSampler.p_sample_ddim -> LatentVisualDiffusion.apply_model -> "
                   "lvdm.modules.networks.openaimodel3d.UNetModel.forward (fp32 eager torch CPU)"}
    return step, info
