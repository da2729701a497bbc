"""Functional fp32 restatement of AutoencoderKL.decode and .encode (TEST INFRASTRUCTURE).

``sd`` is the ``first_stage_model`` state_dict restricted to ``post_quant_conv.*`` and
``decoder.*`` keys; ``dd`` is ``first_stage_config.params.ddconfig``.
Follows lvdm/models/autoencoder.py:104-107 and lvdm/modules/networks/ae_modules.py
(Decoder.forward :539-578, ResnetBlock.forward :190-210, AttnBlock.forward :53-78,
Upsample.forward :123-127, Normalize :15-16 -> GroupNorm(32, eps=1e-6)); encode: autoencoder.py:97-102,
Encoder.forward :438-463, Downsample.forward :103-109 (pad (0,1,0,1) + stride-2 conv).
"""
import torch
import torch.nn.functional as F


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def _swish(x):
    return x * torch.sigmoid(x)


def resnet_block(sd, p, x):
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), 1)
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd, p, x):
    n, c, h, w = x.shape
    y = _gn(sd, p + ".norm", x)
    q = _conv(sd, p + ".q", y, 0).reshape(n, c, h * w).permute(0, 2, 1)
    k = _conv(sd, p + ".k", y, 0).reshape(n, c, h * w)
    v = _conv(sd, p + ".v", y, 0).reshape(n, c, h * w)
    wgt = torch.bmm(q, k) * (int(c) ** -0.5)
    wgt = F.softmax(wgt, dim=2)
    o = torch.bmm(v, wgt.permute(0, 2, 1)).reshape(n, c, h, w)
    return x + _conv(sd, p + ".proj_out", o, 0)


def decoder_layout(dd):
    """(level -> [(cin, cout), ...], has_upsample) in execution order (ae_modules.py:510-525)."""
    ch, mult, nrb = dd["ch"], list(dd["ch_mult"]), dd["num_res_blocks"]
    block_in = ch * mult[-1]
    levels = []
    for lvl in reversed(range(len(mult))):
        out = ch * mult[lvl]
        blocks = []
        for _ in range(nrb + 1):
            blocks.append((block_in, out))
            block_in = out
        levels.append((lvl, blocks, lvl != 0))
    return ch * mult[-1], levels, block_in


@torch.no_grad()
def decode(sd, dd, z):
    """z: [n, z_channels, h, w] (already divided by scale_factor) -> [n, out_ch, 8h, 8w]."""
    z = _conv(sd, "post_quant_conv", z.float(), 0)
    h = _conv(sd, "decoder.conv_in", z, 1)
    h = resnet_block(sd, "decoder.mid.block_1", h)
    h = attn_block(sd, "decoder.mid.attn_1", h)
    h = resnet_block(sd, "decoder.mid.block_2", h)
    _, levels, _ = decoder_layout(dd)
    for lvl, blocks, up in levels:
        for i in range(len(blocks)):
            h = resnet_block(sd, f"decoder.up.{lvl}.block.{i}", h)
        if up:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", h, 1)
    h = _swish(_gn(sd, "decoder.norm_out", h))
    return _conv(sd, "decoder.conv_out", h, 1)


@torch.no_grad()
def decode_first_stage(sd, dd, z5, scale_factor=0.18215):
    """ddpm3d.py:646-667: [b, c, t, h, w] latents -> [b, 3, t, 8h, 8w], frame by frame."""
    b, c, t, h, w = z5.shape
    z = z5.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    frames = [decode(sd, dd, 1.0 / scale_factor * z[i:i + 1]) for i in range(z.shape[0])]
    out = torch.cat(frames, dim=0)
    return out.reshape(b, t, out.shape[1], out.shape[2], out.shape[3]).permute(0, 2, 1, 3, 4).contiguous()


@torch.no_grad()
def encode(sd, dd, x):
    """x: [n, 3, H, W] image in [-1, 1] -> moments [n, 2*embed_dim, H/8, W/8] (mean | logvar), i.e. the
    ``parameters`` of the reference's DiagonalGaussianDistribution (autoencoder.py:97-102)."""
    ch, mult, nrb = dd["ch"], list(dd["ch_mult"]), dd["num_res_blocks"]
    h = _conv(sd, "encoder.conv_in", x.float(), 1)
    for lvl in range(len(mult)):
        for blk in range(nrb):
            h = resnet_block(sd, f"encoder.down.{lvl}.block.{blk}", h)
        if lvl != len(mult) - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = F.conv2d(h, sd[f"encoder.down.{lvl}.downsample.conv.weight"], sd[f"encoder.down.{lvl}.downsample.conv.bias"],
                         stride=2, padding=0)
    h = resnet_block(sd, "encoder.mid.block_1", h)
    h = attn_block(sd, "encoder.mid.attn_1", h)
    h = resnet_block(sd, "encoder.mid.block_2", h)
    h = _conv(sd, "encoder.conv_out", _swish(_gn(sd, "encoder.norm_out", h)), 1)
    return _conv(sd, "quant_conv", h, 0)
