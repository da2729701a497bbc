"""Drop-in for lvdm/models/samplers/ddim_multiplecond.py: the DDIM sampler with a third guidance
branch -- cond, uncond and "image only, no text" -- combined as

    e_uncond + cfg_img * (e_uncond_img - e_uncond) + unconditional_guidance_scale * (e_cond - e_uncond_img)

(ddim_multiplecond.py:229-236).  The reference file differs from ddim.py only in p_sample_ddim; here the three branches
run as ONE batch-3B UNet forward and the combine + guidance rescale + v/eps conversion + x_{t-1} update stay one fused
kernel (dcb_ddim_update_multicond, arithmetic in the reference's order)."""
import torch

from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm.models.samplers.ddim import DDIMSampler as _DDIMSampler


class DDIMSampler(_DDIMSampler):
    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, uc_type=None, cfg_img=None,
                      mask=None, x0=None, guidance_rescale=0.0, **kwargs):
        if use_original_steps or quantize_denoised or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError("use_original_steps / quantize_denoised / score_corrector / noise_dropout "
                                      "are not used on the DynamiCrafter inference path")
        if not x.is_cuda:
            raise RuntimeError("dcb200 DDIMSampler runs on CUDA tensors only (no CPU path)")
        b, device = x.shape[0], x.device
        x = x.float().contiguous()
        if cfg_img is None:
            cfg_img = unconditional_guidance_scale
        uc_img = kwargs['unconditional_conditioning_img_nonetext']        # KeyError like the reference (:223)
        e_u = e_ui = None
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_c = self.model.apply_model(x, t, c, **kwargs)
        else:
            e_c, e_u, e_ui = self._eval_branches(x, t, [c, unconditional_conditioning, uc_img], kwargs)
            e_u, e_ui = e_u.float().contiguous(), e_ui.float().contiguous()
        e_c = e_c.float().contiguous()
        hint = getattr(self, "_step_hint", None)
        step = hint if hint is not None else int(t.reshape(-1)[0])
        coeffs = self._coeffs(index, step, unconditional_guidance_scale, guidance_rescale if e_u is not None else 0.0,
                              temperature)
        if repeat_noise:
            noise = torch.randn((1, *x.shape[1:]), device=device).repeat(b, *((1,) * (x.dim() - 1)))
        else:
            noise = torch.randn(x.shape, device=device)
        return ops.ddim_update(x, e_c, e_u, noise, coeffs, e_uncond_img=e_ui, cfg_img=float(cfg_img))
