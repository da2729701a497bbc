"""Functional fp32 restatement of the lvdm 3D UNet forward (TEST INFRASTRUCTURE).

Everything is a pure function of ``(sd, cfg, inputs)`` where ``sd`` is a flat
``state_dict`` with the reference's key names and ``cfg`` the ``unet_config.params``
mapping of configs/inference_*.yaml.  Citations are into /root/reference.
"""
import math

import torch
import torch.nn.functional as F

TEXT_LEN = 77          # attention.py:45 text_context_len
IMG_TOKENS = 16        # openaimodel3d.py:556 "77 + t*16"


# --------------------------------------------------------------------------- structure
def unet_layout(cfg):
    """Enumerate the blocks UNetModel.__init__ builds (openaimodel3d.py:383-546).

    Returns dict(input=[[layer,...],...], middle=[...], output=[[...],...]) where a
    layer is a tuple: ("conv_in", cin, cout) | ("res", cin, cout) | ("st", ch, heads, dhead)
    | ("tt", ch, heads, dhead) | ("down", ch) | ("up", ch).
    """
    mc = cfg["model_channels"]
    mult = list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    attn_res = list(cfg["attention_resolutions"])
    dh = cfg["num_head_channels"]
    temporal = cfg.get("temporal_attention", True)

    def attn_layers(ch):
        heads = ch // dh
        out = [("st", ch, heads, dh)]
        if temporal:
            out.append(("tt", ch, heads, dh))
        return out

    inp = [[("conv_in", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers += attn_layers(ch)
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch)] + attn_layers(ch) + [("res", ch, ch)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, m * mc)]
            ch = m * mc
            if ds in attn_res:
                layers += attn_layers(ch)
            if level and i == nrb:
                layers.append(("up", ch))
                ds //= 2
            out.append(layers)
    return dict(input=inp, middle=middle, output=out)


# --------------------------------------------------------------------------- primitives
def sinusoid(t, dim, max_period=10000):
    """utils_diffusion.py:8-28 (cos || sin, fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _lin(sd, p, x, bias=True):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias") if bias else None)


def _gn(sd, p, x, eps):
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _mlp2(sd, p, x):
    """time_embed / fps_embedding: Linear-SiLU-Linear (openaimodel3d.py:370-380)."""
    return _lin(sd, p + ".2", F.silu(_lin(sd, p + ".0", x)))


def _heads(t, h):
    b, n, c = t.shape
    return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)


def _unheads(t, h):
    bh, n, d = t.shape
    return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)


def _sdpa(q, k, v, scale):
    sim = torch.einsum("bid,bjd->bij", q, k) * scale
    return torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)


def cross_attention(sd, p, x, context, heads, dhead, image_xattn, learnable_scale):
    """attention.py:81-144 (eager path; masks / relative positions unused by the configs)."""
    q = _lin(sd, p + ".to_q", x, bias=False)
    self_attn = context is None
    ctx = x if self_attn else context
    k_ip = v_ip = None
    if image_xattn and not self_attn:
        ctx, ctx_img = ctx[:, :TEXT_LEN], ctx[:, TEXT_LEN:]
        k_ip = _lin(sd, p + ".to_k_ip", ctx_img, bias=False)
        v_ip = _lin(sd, p + ".to_v_ip", ctx_img, bias=False)
    elif not self_attn:
        ctx = ctx[:, :TEXT_LEN]
    k = _lin(sd, p + ".to_k", ctx, bias=False)
    v = _lin(sd, p + ".to_v", ctx, bias=False)
    scale = dhead ** -0.5
    qh = _heads(q, heads)
    out = _unheads(_sdpa(qh, _heads(k, heads), _heads(v, heads), scale), heads)
    if k_ip is not None:
        out_ip = _unheads(_sdpa(qh, _heads(k_ip, heads), _heads(v_ip, heads), scale), heads)
        if learnable_scale:
            out = out + 1.0 * out_ip * (torch.tanh(sd[p + ".alpha"]) + 1)   # attention.py:139-140
        else:
            out = out + 1.0 * out_ip
    return _lin(sd, p + ".to_out.0", out)


def transformer_block(sd, p, x, context, heads, dhead, image_xattn, learnable_scale):
    """attention.py:242-246 + GEGLU feed-forward :415-442."""
    def ln(name, t):
        return F.layer_norm(t, (t.shape[-1],), sd[f"{p}.{name}.weight"], sd[f"{p}.{name}.bias"], 1e-5)
    x = cross_attention(sd, p + ".attn1", ln("norm1", x), None, heads, dhead, False, False) + x
    x = cross_attention(sd, p + ".attn2", ln("norm2", x), context, heads, dhead,
                        image_xattn, learnable_scale) + x
    hgate = _lin(sd, p + ".ff.net.0.proj", ln("norm3", x))
    a, gate = hgate.chunk(2, dim=-1)
    x = _lin(sd, p + ".ff.net.2", a * F.gelu(gate)) + x
    return x


def spatial_transformer(sd, p, x, context, heads, dhead, learnable_scale, image_xattn=True):
    """attention.py:294-310, use_linear=True.  x: [(b t), c, h, w]."""
    n, c, h, w = x.shape
    t = _gn(sd, p + ".norm", x, 1e-6)
    t = t.permute(0, 2, 3, 1).reshape(n, h * w, c)
    t = _lin(sd, p + ".proj_in", t)
    t = transformer_block(sd, p + ".transformer_blocks.0", t, context, heads, dhead,
                          image_xattn, learnable_scale)
    t = _lin(sd, p + ".proj_out", t)
    return t.reshape(n, h, w, c).permute(0, 3, 1, 2) + x


def temporal_transformer(sd, p, x5, heads, dhead):
    """attention.py:365-412 with only_self_att=True.  x5: [b, c, t, h, w].

    proj_in / proj_out are nn.Linear when use_linear, else kernel-1 Conv1d
    (weight [out, in, 1]); both are per-token affine maps.
    """
    b, c, t, h, w = x5.shape
    y = _gn(sd, p + ".norm", x5, 1e-6)
    y = y.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)

    def proj(name, z):
        wt = sd[f"{p}.{name}.weight"]
        return F.linear(z, wt.reshape(wt.shape[0], wt.shape[1]), sd[f"{p}.{name}.bias"])
    y = proj("proj_in", y)
    y = transformer_block(sd, p + ".transformer_blocks.0", y, None, heads, dhead, False, False)
    y = proj("proj_out", y)
    y = y.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)
    return y + x5


def temporal_conv_block(sd, p, x5):
    """openaimodel3d.py:272-279: 4 x [GN32(eps 1e-5) - SiLU - Conv3d(3,1,1)] + identity."""
    y = x5
    for i, norm_idx, conv_idx in ((1, 0, 2), (2, 0, 3), (3, 0, 3), (4, 0, 3)):
        q = f"{p}.conv{i}"
        y = F.silu(_gn(sd, f"{q}.{norm_idx}", y, 1e-5))
        y = F.conv3d(y, sd[f"{q}.{conv_idx}.weight"], sd[f"{q}.{conv_idx}.bias"], padding=(1, 0, 0))
    return x5 + y


def res_block(sd, p, x, emb, batch, temporal_conv=True):
    """openaimodel3d.py:210-236 (no up/down, no scale-shift norm).  x: [(b t), c, h, w]."""
    h = F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5))
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = _lin(sd, p + ".emb_layers.1", F.silu(emb))
    h = h + e[:, :, None, None]
    h = F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    h = x + h
    if temporal_conv and (p + ".temopral_conv.conv1.0.weight") in sd:
        n, c, hh, ww = h.shape
        h5 = h.reshape(batch, n // batch, c, hh, ww).permute(0, 2, 1, 3, 4)
        h5 = temporal_conv_block(sd, p + ".temopral_conv", h5)
        h = h5.permute(0, 2, 1, 3, 4).reshape(n, c, hh, ww)
    return h


def _run_layers(sd, prefix, layers, h, emb, context, batch, learnable_scale):
    for j, layer in enumerate(layers):
        p = f"{prefix}.{j}"
        kind = layer[0]
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = res_block(sd, p, h, emb, batch)
        elif kind == "st":
            h = spatial_transformer(sd, p, h, context, layer[2], layer[3], learnable_scale)
        elif kind == "tt":
            n, c, hh, ww = h.shape
            h5 = h.reshape(batch, n // batch, c, hh, ww).permute(0, 2, 1, 3, 4)
            h5 = temporal_transformer(sd, p, h5, layer[2], layer[3])
            h = h5.permute(0, 2, 1, 3, 4).reshape(n, c, hh, ww)
        elif kind == "down":
            h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        else:
            raise ValueError(kind)
    return h


def split_context(context, t):
    """openaimodel3d.py:555-562: per-frame context [(b t), 77+16, C] (or repeat)."""
    if context.shape[1] == TEXT_LEN + t * IMG_TOKENS:
        text, img = context[:, :TEXT_LEN], context[:, TEXT_LEN:]
        text = text.repeat_interleave(t, dim=0)
        img = img.reshape(img.shape[0] * t, IMG_TOKENS, img.shape[-1])
        return torch.cat([text, img], dim=1)
    return context.repeat_interleave(t, dim=0)


@torch.no_grad()
def unet_forward(sd, cfg, x, timesteps, context, fs=None):
    """openaimodel3d.py:548-603.  x: [b, c, t, h, w] fp32 -> [b, out_c, t, h, w]."""
    b, _, t, hh, ww = x.shape
    mc = cfg["model_channels"]
    learnable = bool(cfg.get("image_cross_attention_scale_learnable", False))
    emb = _mlp2(sd, "time_embed", sinusoid(timesteps, mc))
    ctx = split_context(context.float(), t)
    emb = emb.repeat_interleave(t, dim=0)
    if cfg.get("fs_condition", False):
        if fs is None:
            fs = torch.tensor([cfg.get("default_fs", 4)] * b, dtype=torch.long)
        emb = emb + _mlp2(sd, "fps_embedding", sinusoid(fs, mc)).repeat_interleave(t, dim=0)
    h = x.permute(0, 2, 1, 3, 4).reshape(b * t, x.shape[1], hh, ww).float()
    lay = unet_layout(cfg)
    hs = []
    for i, layers in enumerate(lay["input"]):
        h = _run_layers(sd, f"input_blocks.{i}", layers, h, emb, ctx, b, learnable)
        if i == 0 and cfg.get("addition_attention", False):
            n, c, _, _ = h.shape
            h5 = h.reshape(b, t, c, hh, ww).permute(0, 2, 1, 3, 4)
            h5 = temporal_transformer(sd, "init_attn.0", h5, 8, cfg["num_head_channels"])
            h = h5.permute(0, 2, 1, 3, 4).reshape(n, c, hh, ww)
        hs.append(h)
    h = _run_layers(sd, "middle_block", lay["middle"], h, emb, ctx, b, learnable)
    for i, layers in enumerate(lay["output"]):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_layers(sd, f"output_blocks.{i}", layers, h, emb, ctx, b, learnable)
    y = F.silu(_gn(sd, "out.0", h, 1e-5))
    y = F.conv2d(y, sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return y.reshape(b, t, y.shape[1], hh, ww).permute(0, 2, 1, 3, 4).contiguous()
