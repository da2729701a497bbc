"""Drop-in mirror of lvdm/models/samplers/ddim.py::DDIMSampler on the dcb200 kernels.

Same public surface as the reference (DDIMSampler(model), make_schedule ddim.py:24, sample :60,
ddim_sampling :135, p_sample_ddim :206) and the same duck-typed ``model`` (num_timesteps, alphas_cumprod,
betas, use_dynamic_rescale/scale_arr, parameterization, apply_model, sqrt_*_alphas_cumprod).

What changes:
  * the ~25 pointwise launches + 2 std reductions after the two UNet calls of a step (ddim.py:226-277) are
    ONE fused kernel (dcb_ddim_update); its scalar coefficients are computed on the host with the reference's
    own fp32 torch expressions and order (SURVEY.md 8(a) gotcha 7), so results match bit for bit when
    guidance_rescale == 0 and to fp32 round-off otherwise;
  * the cond / uncond UNet evaluations are issued as one batch-2B ``apply_model`` call (identical numerics per
    sample, twice the rows per GEMM) when the two conditionings have the same structure;
  * schedule tables stay on the host (the reference forces them onto "cuda" and reads them back one scalar at a
    time, ddim.py:18-22,251-254); the per-step noise still comes from ``torch.randn`` on the sampling device,
    i.e. the same Philox stream the reference consumes (ddim.py:273), and is drawn even when sigma == 0.
"""
import numpy as np
import torch

from dynamicrafter_b200 import ops
from dynamicrafter_b200._lib import DdimCoeffs


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=False):
    """utils_diffusion.py:56-76."""
    if ddim_discr_method == 'uniform':
        c = num_ddpm_timesteps // num_ddim_timesteps
        steps = np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1
    elif ddim_discr_method == 'uniform_trailing':
        c = num_ddpm_timesteps / num_ddim_timesteps
        steps = np.flip(np.round(np.arange(num_ddpm_timesteps, 0, -c))).astype(np.int64) - 1
    elif ddim_discr_method == 'quad':
        steps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int) + 1
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    if verbose:
        print(f'Selected timesteps for ddim sampler: {steps}')
    return steps


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.counter = 0
        self.batch_cfg = True          # run cond + uncond as one 2B batch when possible

    def register_buffer(self, name, attr):
        setattr(self, name, attr)

    # ------------------------------------------------------------------------------------ schedule
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose)
        acp = self.model.alphas_cumprod
        assert acp.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        acp_cpu = acp.detach().to(torch.float32).cpu()
        ts = self.ddim_timesteps
        if getattr(self.model, "use_dynamic_rescale", False):
            self.ddim_scale_arr = self.model.scale_arr.detach().cpu()[ts]
            self.ddim_scale_arr_prev = torch.cat([self.ddim_scale_arr[0:1], self.ddim_scale_arr[:-1]])
        self.register_buffer('alphas_cumprod', acp_cpu.clone())
        # same operand types as utils_diffusion.py:82-86: fp32 tensor, fp64 ndarray -> fp64 tensor sigmas
        alphas = acp_cpu[ts]
        alphas_prev = np.asarray([acp_cpu[0]] + acp_cpu[ts[:-1]].tolist())
        sigmas = ddim_eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
        self.register_buffer('ddim_sigmas', sigmas)
        self.register_buffer('ddim_alphas', alphas)
        self.register_buffer('ddim_alphas_prev', alphas_prev)
        self.register_buffer('ddim_sqrt_one_minus_alphas', np.sqrt(1. - alphas))
        # 1000-entry tables used by the v-parameterisation (ddpm3d.py:239-251), host copies
        self._sqrt_acp = self.model.sqrt_alphas_cumprod.detach().to(torch.float32).cpu()
        self._sqrt_1m_acp = self.model.sqrt_one_minus_alphas_cumprod.detach().to(torch.float32).cpu()

    def _coeffs(self, index, step, cfg_scale, guidance_rescale, temperature):
        """Scalar coefficients of one step, each produced by the reference's fp32 expression (ddim.py:251-277)."""
        one = (1,)
        a_t = torch.full(one, self.ddim_alphas[index])
        a_prev = torch.full(one, self.ddim_alphas_prev[index])
        sigma_t = torch.full(one, self.ddim_sigmas[index])
        s1m_at = torch.full(one, self.ddim_sqrt_one_minus_alphas[index])
        c = DdimCoeffs()
        c.is_v = 1 if self.model.parameterization == "v" else 0
        c.cfg_scale = float(cfg_scale)
        c.guidance_rescale = float(guidance_rescale)
        c.one_minus_guidance_rescale = float(1 - guidance_rescale)
        c.sqrt_acp_t = float(self._sqrt_acp[step])
        c.sqrt_1m_acp_t = float(self._sqrt_1m_acp[step])
        c.sqrt_a_t = float(a_t.sqrt())
        c.sqrt_1m_a_t = float(s1m_at)
        c.sqrt_a_prev = float(a_prev.sqrt())
        c.dir_coef = float((1. - a_prev - sigma_t ** 2).sqrt())
        c.sigma_t = float(sigma_t)
        c.temperature = float(temperature)
        if getattr(self.model, "use_dynamic_rescale", False):
            c.x0_rescale = float(torch.full(one, self.ddim_scale_arr_prev[index]) / torch.full(one, self.ddim_scale_arr[index]))
        else:
            c.x0_rescale = 1.0
        return c

    # ------------------------------------------------------------------------------------ sampling
    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1.,
               noise_dropout=0., score_corrector=None, corrector_kwargs=None, verbose=True,
               schedule_verbose=False, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, precision=None, fs=None, timestep_spacing='uniform',
               guidance_rescale=0.0, **kwargs):
        if conditioning is not None:
            first = conditioning[list(conditioning.keys())[0]] if isinstance(conditioning, dict) else conditioning
            cbs = first[0].shape[0] if isinstance(first, (list, tuple)) else first.shape[0]
            if cbs != batch_size:
                print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_discretize=timestep_spacing, ddim_eta=eta, verbose=schedule_verbose)
        if len(shape) == 3:
            size = (batch_size, *shape)
        elif len(shape) == 4:
            size = (batch_size, *shape)
        else:
            raise ValueError(f"shape must be (C,H,W) or (C,T,H,W), got {shape}")
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t, unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, verbose=verbose,
                                  precision=precision, fs=fs, guidance_rescale=guidance_rescale, **kwargs)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, verbose=True, precision=None,
                      fs=None, guidance_rescale=0.0, **kwargs):
        if ddim_use_original_steps:
            raise NotImplementedError("ddim_use_original_steps is not used on the DynamiCrafter inference path")
        device = self.model.betas.device
        b = shape[0]
        img = self._sync_noise(torch.randn(shape, device=device)) if x_T is None else x_T
        if precision is not None and precision == 16:
            raise NotImplementedError("precision=16 latents: dcb200 keeps the DDIM state in fp32")
        if timesteps is None:
            timesteps = self.ddim_timesteps
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        clean_cond = kwargs.pop("clean_cond", False)
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img_orig = x0 if clean_cond else self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            self._step_hint = int(step)
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, quantize_denoised=quantize_denoised,
                                              temperature=temperature, noise_dropout=noise_dropout,
                                              score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning,
                                              mask=mask, x0=x0, fs=fs, guidance_rescale=guidance_rescale, **kwargs)
            self._step_hint = None
            if callback:
                callback(i)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def decode(self, x_latent, cond, t_start, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
               use_original_steps=False, callback=None, **kwargs):
        """ddim.py:282-301: denoise `x_latent` from DDIM index t_start - 1 down to 0 (img2img-style partial runs).
        Extra kwargs (fs, guidance_rescale, ...) are forwarded to p_sample_ddim."""
        if use_original_steps:
            raise NotImplementedError("use_original_steps is not used on the DynamiCrafter inference path")
        timesteps = self.ddim_timesteps[:t_start]
        time_range = np.flip(timesteps)
        total_steps = timesteps.shape[0]
        x_dec = x_latent
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((x_latent.shape[0],), int(step), device=x_latent.device, dtype=torch.long)
            self._step_hint = int(step)
            x_dec, _ = self.p_sample_ddim(x_dec, cond, ts, index=index,
                                          unconditional_guidance_scale=unconditional_guidance_scale,
                                          unconditional_conditioning=unconditional_conditioning, **kwargs)
            self._step_hint = None
            if callback:
                callback(i)
        return x_dec

    @torch.no_grad()
    def stochastic_encode(self, x0, t, use_original_steps=False, noise=None):
        """ddim.py:304-319: x_t = sqrt(a_t) x0 + sqrt(1 - a_t) noise with the DDIM tables gathered at index t."""
        if use_original_steps:
            sa, s1m = self._sqrt_acp, self._sqrt_1m_acp
        else:
            sa = torch.sqrt(torch.as_tensor(self.ddim_alphas, dtype=torch.float32))
            s1m = torch.as_tensor(self.ddim_sqrt_one_minus_alphas, dtype=torch.float32)
        if noise is None:
            noise = torch.randn_like(x0)
        ti = t.detach().to("cpu", torch.long).reshape(-1)
        return ops.q_sample(x0, noise, sa[ti], s1m[ti])

    # cond / uncond (/ image-only uncond) as one batch -----------------------------------------------
    @staticmethod
    def _cat_cond(*cs):
        """Batch-concatenate structurally identical conditionings (tensors, or dicts of tensors / tensor lists);
        None if they do not line up (the branches are then evaluated one by one)."""
        def same(ts):
            return all(isinstance(t, torch.Tensor) for t in ts) and all(t.shape == ts[0].shape for t in ts)
        if same(cs):
            return torch.cat(cs, 0)
        if all(isinstance(c, dict) for c in cs) and all(c.keys() == cs[0].keys() for c in cs):
            out = {}
            for k in cs[0]:
                vals = [c[k] for c in cs]
                if all(isinstance(v, (list, tuple)) for v in vals) and all(len(v) == len(vals[0]) for v in vals) and \
                        all(same([v[j] for v in vals]) for j in range(len(vals[0]))):
                    out[k] = [torch.cat([v[j] for v in vals], 0) for j in range(len(vals[0]))]
                elif same(vals):
                    out[k] = torch.cat(vals, 0)
                else:
                    return None
            return out
        return None

    def _sync_noise(self, noise):
        """Single-video sharded modes (branch_shard / UNetModel.frame_shard): every rank applies the DDIM update to the
        whole latent, so all ranks must use rank 0's random draws -- broadcast them instead of relying on the callers
        having seeded every rank identically (with per-rank seeds the ranks would silently diverge)."""
        groups = []
        bs = getattr(self, "branch_shard", None)
        if bs is not None:
            groups.append(bs.group)
        fs = getattr(getattr(self.model, "unet", None), "frame_shard", None)
        if fs is not None:
            groups.append(fs.group)
        if groups:
            import torch.distributed as dist
            for g in groups:                     # branch x frame grids: the two broadcasts reach every rank from rank 0
                dist.broadcast(noise, src=dist.get_global_rank(g, 0) if g is not None else 0, group=g)
        return noise

    @staticmethod
    def _cond_leaves(c):
        if isinstance(c, torch.Tensor):
            return [c]
        if isinstance(c, dict):
            out = []
            for k in sorted(c):
                v = c[k]
                out += list(v) if isinstance(v, (list, tuple)) else [v]
            return out
        return [c]

    def _cat_cond_cached(self, conds):
        """_cat_cond, remembered while the caller keeps passing the SAME conditioning tensors (same objects, unchanged
        versions) -- every step of a sampling loop does.  The batched conditioning is then the same tensor object from
        step to step, which lets UNetModel reuse the cross-attention keys / values it projected from it."""
        leaves = [t for c in conds for t in self._cond_leaves(c)]
        if not all(isinstance(t, torch.Tensor) for t in leaves):
            return self._cat_cond(*conds)
        ent = getattr(self, "_cat_cache", None)
        vers = [t._version for t in leaves]
        if ent is not None and len(ent[0]) == len(leaves) and all(a is b for a, b in zip(ent[0], leaves)) and ent[1] == vers:
            return ent[2]
        allc = self._cat_cond(*conds)
        self._cat_cache = (leaves, vers, allc)           # holds the leaves: their storage cannot be recycled meanwhile
        return allc

    def _eval_branches(self, x, t, conds, kwargs):
        """One UNet forward over all guidance branches (same x, t; different conditioning) -> list of outputs."""
        b, n = x.shape[0], len(conds)
        shard = getattr(self, "branch_shard", None)      # parallel.BranchShard: one guidance branch per GPU
        if shard is not None:
            if shard.world != n:
                raise ValueError(f"branch sharding needs {n} ranks (one per guidance branch), got {shard.world}")
            return shard.gather(self.model.apply_model(x, t, conds[shard.rank], **kwargs))
        allc = self._cat_cond_cached(conds) if self.batch_cfg else None
        if allc is None:
            return [self.model.apply_model(x, t, c, **kwargs) for c in conds]
        kwn = {k: (torch.cat([v] * n, 0) if isinstance(v, torch.Tensor) and v.dim() >= 1 and v.shape[0] == b else v)
               for k, v in kwargs.items()}
        out = self.model.apply_model(torch.cat([x] * n, 0), torch.cat([t] * n, 0), allc, **kwn)
        return [out[i * b:(i + 1) * b].contiguous() for i in range(n)]

    def _eval_pair(self, x, t, c, uc, kwargs):
        return tuple(self._eval_branches(x, t, [c, uc], kwargs))

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, uc_type=None,
                      conditional_guidance_scale_temporal=None, mask=None, x0=None, guidance_rescale=0.0, **kwargs):
        if use_original_steps or quantize_denoised or score_corrector is not None or noise_dropout > 0.:
            raise NotImplementedError("use_original_steps / quantize_denoised / score_corrector / noise_dropout "
                                      "are not used on the DynamiCrafter inference path")
        if not x.is_cuda:
            raise RuntimeError("dcb200 DDIMSampler runs on CUDA tensors only (no CPU path)")
        b, device = x.shape[0], x.device
        x = x.float().contiguous()
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_c, e_u = self.model.apply_model(x, t, c, **kwargs), None
        else:
            if not (isinstance(c, torch.Tensor) or isinstance(c, dict)):
                raise NotImplementedError
            e_c, e_u = self._eval_pair(x, t, c, unconditional_conditioning, kwargs)
            e_u = e_u.float().contiguous()
        e_c = e_c.float().contiguous()
        # the DDPM timestep is known on the host inside ddim_sampling; only external callers pay a D2H read.
        # (All samples of a batch share one timestep during sampling, ddim.py:171.)
        hint = getattr(self, "_step_hint", None)
        step = hint if hint is not None else int(t.reshape(-1)[0])
        coeffs = self._coeffs(index, step, unconditional_guidance_scale, guidance_rescale if e_u is not None else 0.0,
                              temperature)
        if repeat_noise:
            noise = torch.randn((1, *x.shape[1:]), device=device).repeat(b, *((1,) * (x.dim() - 1)))
        else:
            noise = torch.randn(x.shape, device=device)              # drawn every step, like ddim.py:273
        if coeffs.sigma_t != 0.0:
            noise = self._sync_noise(noise)
        return ops.ddim_update(x, e_c, e_u, noise, coeffs)
