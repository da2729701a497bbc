"""Functional fp32 restatement of the image-context Resampler (TEST INFRASTRUCTURE): lvdm/modules/encoders/resampler.py
:48-145 (PerceiverAttention :48-94, FeedForward :27-34, Resampler.forward :132-145)."""
import math

import torch
import torch.nn.functional as F


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _heads(x, heads):
    b, n, w = x.shape
    return x.view(b, n, heads, -1).transpose(1, 2)


def perceiver_attention(sd, p, x, latents, heads, dim_head):
    x = _ln(sd, p + ".norm1", x)
    latents = _ln(sd, p + ".norm2", latents)
    b, l, _ = latents.shape
    q = latents @ sd[p + ".to_q.weight"].t()
    kv = torch.cat((x, latents), dim=-2) @ sd[p + ".to_kv.weight"].t()
    k, v = kv.chunk(2, dim=-1)
    q, k, v = _heads(q, heads), _heads(k, heads), _heads(v, heads)
    scale = 1 / math.sqrt(math.sqrt(dim_head))
    w = torch.softmax(((q * scale) @ (k * scale).transpose(-2, -1)).float(), dim=-1)
    out = (w @ v).permute(0, 2, 1, 3).reshape(b, l, -1)
    return out @ sd[p + ".to_out.weight"].t()


def feed_forward(sd, p, x):
    h = _ln(sd, p + ".0", x) @ sd[p + ".1.weight"].t()
    return F.gelu(h) @ sd[p + ".3.weight"].t()


@torch.no_grad()
def resampler_forward(sd, cfg, x):
    """x: [b, n_tokens, embedding_dim] CLIP image tokens -> [b, num_queries * video_length, output_dim]."""
    latents = sd["latents"].repeat(x.shape[0], 1, 1)
    x = x.float() @ sd["proj_in.weight"].t() + sd["proj_in.bias"]
    for i in range(cfg["depth"]):
        latents = perceiver_attention(sd, f"layers.{i}.0", x, latents, cfg["heads"], cfg["dim_head"]) + latents
        latents = feed_forward(sd, f"layers.{i}.1", latents) + latents
    latents = latents @ sd["proj_out.weight"].t() + sd["proj_out.bias"]
    return _ln(sd, "norm_out", latents)
