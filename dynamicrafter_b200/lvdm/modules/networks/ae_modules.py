"""Drop-in mirror of the VAE Decoder path of lvdm/modules/networks/ae_modules.py on the dcb200 kernels.

Decoder (ae_modules.py:466-578), ResnetBlock (:151-210), AttnBlock (:26-78), Upsample (:111-127) keep the
reference constructor arguments and ``state_dict`` keys.  The convolutions are the same tcgen05 implicit-GEMM
kernel as the UNet; the single-head d=C mid-block attention is a flash-style tcgen05 kernel (dcb_vae_attention: one
launch for all frames, the N x N scores never reach HBM).

``Encoder`` is a parameter holder only (checkpoint compatibility); encoding runs once per video and is outside
the hot path.
"""
import numpy as np
import torch
import torch.nn as nn

from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm import _packing
from dynamicrafter_b200.lvdm._packing import PackedModule, Act, pack_linear, pack_conv3x3, pack_conv3x3_up2


def _f32(t):
    return t.detach().float().contiguous()


def Normalize(in_channels, num_groups=32):
    return torch.nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def _gn(norm, h, silu):
    return ops.group_norm(h.x, h.B * h.T, h.H * h.W, norm.weight, norm.bias, norm.eps, silu=silu, colsums=h.cs)


class Upsample(PackedModule):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("resamp_with_conv=False is not used by DynamiCrafter")
        self.with_conv, self.in_channels = with_conv, in_channels
        self.conv = torch.nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def _pack(self):
        if _packing.FOLD_UPSAMPLE:
            return {"w4": pack_conv3x3_up2(self.conv.weight), "b": _f32(self.conv.bias)}
        return {"w": pack_conv3x3(self.conv.weight), "b": _f32(self.conv.bias)}

    def _fwd(self, h):
        p = self.packed()
        nf = h.B * h.T
        if "w4" in p:            # nearest x2 folded into the conv (ae_modules.py:123-127): four 2x2-tap phase GEMMs
            xb = h.x if h.x.dtype == torch.bfloat16 else ops.cast_bf16(h.x)
            out, cs = ops.conv3x3_up2(xb, nf, h.H, h.W, p["w4"], bias=p["b"], stats=True)
            return h.like(out, 2 * h.H, 2 * h.W, cs=cs)
        up = ops.upsample2x(h.x, nf, h.H, h.W)
        out, cs = ops.conv3x3(up, nf, 2 * h.H, 2 * h.W, p["w"], bias=p["b"], stats=True)
        return h.like(out, 2 * h.H, 2 * h.W, cs=cs)


class Downsample(PackedModule):
    """ae_modules.py:90-109: pad (0,1,0,1) then 3x3 stride-2 conv = implicit GEMM over the four 2x2 phase planes
    with tap k reading pixel 2i + k (TMA zero fill supplies the right / bottom padding)."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("avg-pool downsampling is not used by DynamiCrafter")
        self.with_conv, self.in_channels = with_conv, in_channels
        self.conv = torch.nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def _pack(self):
        return {"w": pack_conv3x3(self.conv.weight), "b": _f32(self.conv.bias)}

    def _fwd(self, h):
        p = self.packed()
        nf = h.B * h.T
        planes = ops.space_to_depth(h.x, nf, h.H, h.W)
        o, cs = ops.conv3x3_s2(planes, nf, h.H, h.W, p["w"], bias=p["b"], asym=True, stats=True)
        return h.like(o, H=(h.H - 2) // 2 + 1, W=(h.W - 2) // 2 + 1, cs=cs)


class ResnetBlock(PackedModule):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        if conv_shortcut:
            raise NotImplementedError("conv_shortcut is not used by DynamiCrafter")
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = torch.nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if temb_channels > 0:
            self.temb_proj = torch.nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = torch.nn.Dropout(dropout)
        self.conv2 = torch.nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if self.in_channels != self.out_channels:
            self.nin_shortcut = torch.nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def _pack(self):
        p = {"w1": pack_conv3x3(self.conv1.weight), "b1": _f32(self.conv1.bias),
             "w2": pack_conv3x3(self.conv2.weight), "b2": _f32(self.conv2.bias)}
        if self.in_channels != self.out_channels:
            p["ws"] = pack_linear(self.nin_shortcut.weight)
            p["bs"] = _f32(self.nin_shortcut.bias)
        return p

    def _fwd(self, h):
        p = self.packed()
        nf, S = h.B * h.T, h.H * h.W
        sc = "ws" in p
        res = ops.group_norm(h.x, nf, S, self.norm1.weight, self.norm1.bias, self.norm1.eps, silu=True, want_raw=sc,
                             colsums=h.cs)
        y, raw = res if sc else (res, None)
        h1, cs1 = ops.conv3x3(y, nf, h.H, h.W, p["w1"], bias=p["b1"], f32=not _packing.BF16_VAE,
                              bf16=_packing.BF16_VAE, stats=True)
        y2 = ops.group_norm(h1, nf, S, self.norm2.weight, self.norm2.bias, self.norm2.eps, silu=True, colsums=cs1)
        xs = ops.linear(raw, p["ws"], bias=p["bs"], f32=True, bf16=False) if sc else h.x
        out, cs2 = ops.conv3x3(y2, nf, h.H, h.W, p["w2"], bias=p["b2"], residual=xs, stats=True)
        return h.like(out, cs=cs2)

    def forward(self, x, temb=None):
        if temb is not None:
            raise NotImplementedError("temb is None on the VAE path (temb_ch = 0)")
        n, c, hh, ww = x.shape
        rows = x.float().permute(0, 2, 3, 1).reshape(-1, c).contiguous()
        o = self._fwd(Act(rows, n, 1, hh, ww))
        return o.x.reshape(n, hh, ww, -1).permute(0, 3, 1, 2).contiguous()


class AttnBlock(PackedModule):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = torch.nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def _pack(self):
        return {"wqkv": torch.cat([pack_linear(self.q.weight), pack_linear(self.k.weight), pack_linear(self.v.weight)],
                                  0).contiguous(),
                "bqkv": torch.cat([_f32(self.q.bias), _f32(self.k.bias), _f32(self.v.bias)], 0).contiguous(),
                "wo": pack_linear(self.proj_out.weight), "bo": _f32(self.proj_out.bias)}

    def _fwd(self, h):
        """GroupNorm -> fused q/k/v 1x1 convs (one GEMM) -> flash attention with d = C over the h*w tokens of every frame
        (one launch for all frames, scores stay on the SM) -> proj_out + residual (ae_modules.py:53-78)."""
        p = self.packed()
        nf, S, C = h.B * h.T, h.H * h.W, h.C
        y = _gn(self.norm, h, False)                              # bf16 [nf*S, C]
        qkv = ops.linear(y, p["wqkv"], bias=p["bqkv"])            # bf16 [nf*S, 3C]
        scale = float(int(C) ** (-0.5))                           # ae_modules.py:67
        o = ops.vae_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], nf, S, C, scale)
        out, cs = ops.linear(o, p["wo"], bias=p["bo"], residual=h.x, f32=True, bf16=False, stats_rows=S)
        return h.like(out, cs=cs)

    def forward(self, x):
        n, c, hh, ww = x.shape
        rows = x.float().permute(0, 2, 3, 1).reshape(-1, c).contiguous()
        out = self._fwd(Act(rows, n, 1, hh, ww))
        return out.x.reshape(n, hh, ww, c).permute(0, 3, 1, 2).contiguous()


def make_attn(in_channels, attn_type="vanilla"):
    if attn_type != "vanilla":
        raise NotImplementedError("only vanilla attention is used by the DynamiCrafter VAE")
    return AttnBlock(in_channels)


class Encoder(PackedModule):
    """ae_modules.py:364-463 on the dcb200 kernels (same conv / GroupNorm / attention kernels as the Decoder)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True,
                 use_linear_attn=False, attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        if use_linear_attn:
            raise NotImplementedError("linear attention is not used by DynamiCrafter")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = torch.nn.Conv2d(in_channels, self.ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res = curr_res // 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = torch.nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3,
                                        stride=1, padding=1)
        self.out_channels = self.conv_out.out_channels

    def _pack(self):
        return {"w_in": pack_conv3x3(self.conv_in.weight), "b_in": _f32(self.conv_in.bias),
                "w_out": pack_conv3x3(self.conv_out.weight), "b_out": _f32(self.conv_out.bias)}

    def _own_version(self):
        return sum(q._version for q in (self.conv_in.weight, self.conv_in.bias, self.conv_out.weight, self.conv_out.bias))

    def _fwd_rows(self, x_rows, n, hh, ww):
        """x_rows: bf16 [n*hh*ww, 64] channels-last image (zero-padded channels) -> (fp32 [n*(hh/8)*(ww/8), 2z], h, w)."""
        p = self.packed()
        h0, cs0 = ops.conv3x3(x_rows, n, hh, ww, p["w_in"], bias=p["b_in"], stats=True)
        h = Act(h0, n, 1, hh, ww, cs0)
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block]._fwd(h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block]._fwd(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample._fwd(h)
        h = self.mid.block_1._fwd(h)
        h = self.mid.attn_1._fwd(h)
        h = self.mid.block_2._fwd(h)
        y = _gn(self.norm_out, h, True)
        return ops.conv3x3(y, n, h.H, h.W, p["w_out"], bias=p["b_out"]), h.H, h.W

    @torch.no_grad()
    def forward(self, x):
        """x: [n, in_channels, H, W] fp32 -> [n, 2*z_channels, H/8, W/8] (ae_modules.py:438)."""
        if not x.is_cuda:
            raise RuntimeError("dcb200 Encoder runs on CUDA tensors only (no CPU path)")
        n, c, hh, ww = x.shape
        with ops.stats_scope(x.device):
            o, H, W = self._fwd_rows(ops.nchw_to_rows(x, 64), n, hh, ww)
        return ops.rows_to_nchw(o, n, self.out_channels, 1, H, W, five_d=False).to(x.dtype)


class Decoder(PackedModule):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if use_linear_attn or give_pre_end or tanh_out:
            raise NotImplementedError("linear attention / give_pre_end / tanh_out are not used by DynamiCrafter")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.give_pre_end, self.tanh_out = give_pre_end, tanh_out
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = torch.nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = make_attn(block_in, attn_type=attn_type)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(self.num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(make_attn(block_in, attn_type=attn_type))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res = curr_res * 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = torch.nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self.out_ch = out_ch

    def _pack(self):
        return {"w_in": pack_conv3x3(self.conv_in.weight), "b_in": _f32(self.conv_in.bias),
                "w_out": pack_conv3x3(self.conv_out.weight), "b_out": _f32(self.conv_out.bias)}

    def _own_version(self):
        return sum(q._version for q in (self.conv_in.weight, self.conv_in.bias, self.conv_out.weight, self.conv_out.bias))

    def _fwd_rows(self, z_rows, n, hh, ww):
        """z_rows: bf16 [n*hh*ww, 64] channels-last latent (zero-padded channels) -> fp32 [n*8hh*8ww, out_ch]."""
        p = self.packed()
        self.last_z_shape = (n, self.conv_in.in_channels, hh, ww)
        h0, cs0 = ops.conv3x3(z_rows, n, hh, ww, p["w_in"], bias=p["b_in"], stats=True)
        h = Act(h0, n, 1, hh, ww, cs0)
        h = self.mid.block_1._fwd(h)
        h = self.mid.attn_1._fwd(h)
        h = self.mid.block_2._fwd(h)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block]._fwd(h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block]._fwd(h)
            if i_level != 0:
                h = self.up[i_level].upsample._fwd(h)
        y = _gn(self.norm_out, h, True)
        return ops.conv3x3(y, n, h.H, h.W, p["w_out"], bias=p["b_out"]), h.H, h.W

    @torch.no_grad()
    def forward(self, z):
        """z: [n, z_channels, h, w] fp32 -> [n, out_ch, 8h, 8w] (ae_modules.py:539)."""
        if not z.is_cuda:
            raise RuntimeError("dcb200 Decoder runs on CUDA tensors only (no CPU path)")
        n, c, hh, ww = z.shape
        with ops.stats_scope(z.device):
            o, H, W = self._fwd_rows(ops.nchw_to_rows(z, 64), n, hh, ww)
        return ops.rows_to_nchw(o, n, self.out_ch, 1, H, W, five_d=False).to(z.dtype)
