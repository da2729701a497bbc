"""Drop-in mirror of lvdm/modules/networks/openaimodel3d.py (reference) on the dcb200 CUDA kernels.

Classes keep the reference's constructor arguments, ``state_dict`` keys (including the ``temopral_conv``
typo, openaimodel3d.py:190) and call signatures: UNetModel.forward(x, timesteps, context,
features_adapter, fs, **kwargs) (openaimodel3d.py:548), ResBlock.forward(x, emb, batch_size) (:197),
TemporalConvBlock.forward(x) (:272), Downsample/Upsample.forward(x) (:75,:98),
TimestepEmbedSequential.forward(x, emb, context, batch_size) (:36).

Execution plan of one UNet forward (all in libdcb200.so):
  pack x (NCTHW fp32 -> channels-last bf16, C padded to 64)  ->  3x3 implicit-GEMM conv_in
  per ResBlock: GN-stats, GN-apply+SiLU (bf16 operand), conv3x3 (+bias +emb), GN, conv3x3 (+bias +skip),
                4 x [GN(5-D) + SiLU, temporal 3-tap conv] (+identity in the last epilogue)
  per transformer: GN, proj_in, [LN, fused-QKV GEMM, attention, out-proj(+res)] x2, LN, GEGLU GEMM, FF2(+res), proj_out(+res)
  skip concats are read as two sources by the GN kernels; the residual stream is fp32, GEMM operands bf16.
"""
import torch
import torch.nn as nn

from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm import _packing
from dynamicrafter_b200.lvdm._packing import PackedModule, Act, pack_linear, pack_conv3x3, pack_tconv3, pack_conv3x3_up2
from dynamicrafter_b200.lvdm.modules.attention import SpatialTransformer, TemporalTransformer, make_context, CrossAttentionContext, TEXT_CONTEXT_LEN

IMG_TOKENS_PER_FRAME = 16      # openaimodel3d.py:556 ("77 + t*16" hard-coded in the reference)


def _f32(t):
    return t.detach().float().contiguous()


def _rows_from_nchw(x):
    n, c, h, w = x.shape
    return x.float().permute(0, 2, 3, 1).reshape(n * h * w, c).contiguous()


def _nchw_from_rows(r, n, h, w):
    return r.reshape(n, h, w, -1).permute(0, 3, 1, 2).contiguous()


class TimestepBlock(nn.Module):
    """Marker: forward takes the timestep embedding as 2nd argument (openaimodel3d.py:19-27)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb, context=None, batch_size=None):
        """Reference-layout dispatch (x: [(b f), c, h, w]); UNetModel uses the fused _fwd path instead."""
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb, batch_size=batch_size)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            elif isinstance(layer, TemporalTransformer):
                n, c, h, w = x.shape
                x5 = x.reshape(batch_size, n // batch_size, c, h, w).permute(0, 2, 1, 3, 4)
                x5 = layer(x5, context)
                x = x5.permute(0, 2, 1, 3, 4).reshape(n, c, h, w)
            else:
                x = layer(x)
        return x


class Downsample(PackedModule):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or not use_conv or padding != 1:
            raise NotImplementedError("only the 2-D stride-2 conv Downsample is used by DynamiCrafter")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv, self.dims = use_conv, dims
        self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)

    def _pack(self):
        return {"w": pack_conv3x3(self.op.weight), "b": _f32(self.op.bias)}

    def _fwd(self, h):
        p = self.packed()
        nf = h.B * h.T
        planes = ops.space_to_depth(h.x, nf, h.H, h.W)
        out, cs = ops.conv3x3_s2(planes, nf, h.H, h.W, p["w"], bias=p["b"], stats=True)
        return h.like(out, (h.H - 1) // 2 + 1, (h.W - 1) // 2 + 1, cs=cs)

    def forward(self, x):
        assert x.shape[1] == self.channels
        n, c, hh, ww = x.shape
        o = self._fwd(Act(_rows_from_nchw(x), n, 1, hh, ww))
        return _nchw_from_rows(o.x, n, o.H, o.W)


class Upsample(PackedModule):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if dims != 2 or not use_conv or padding != 1:
            raise NotImplementedError("only the 2-D nearest+conv Upsample is used by DynamiCrafter")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv, self.dims = use_conv, dims
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=padding)

    def _pack(self):
        if _packing.FOLD_UPSAMPLE:
            return {"w4": pack_conv3x3_up2(self.conv.weight), "b": _f32(self.conv.bias)}
        return {"w": pack_conv3x3(self.conv.weight), "b": _f32(self.conv.bias)}

    def _fwd(self, h):
        p = self.packed()
        nf = h.B * h.T
        if "w4" in p:            # nearest x2 folded into the conv: four 2x2-tap phase GEMMs on the low-resolution tensor
            xb = h.x if h.x.dtype == torch.bfloat16 else ops.cast_bf16(h.x)
            out, cs = ops.conv3x3_up2(xb, nf, h.H, h.W, p["w4"], bias=p["b"], stats=True)
            return h.like(out, 2 * h.H, 2 * h.W, cs=cs)
        up = ops.upsample2x(h.x, nf, h.H, h.W)
        out, cs = ops.conv3x3(up, nf, 2 * h.H, 2 * h.W, p["w"], bias=p["b"], stats=True)
        return h.like(out, 2 * h.H, 2 * h.W, cs=cs)

    def forward(self, x):
        assert x.shape[1] == self.channels
        n, c, hh, ww = x.shape
        o = self._fwd(Act(_rows_from_nchw(x), n, 1, hh, ww))
        return _nchw_from_rows(o.x, n, o.H, o.W)


class TemporalConvBlock(PackedModule):
    def __init__(self, in_channels, out_channels=None, dropout=0.0, spatial_aware=False):
        super().__init__()
        if spatial_aware:
            raise NotImplementedError("tempspatial_aware convs are not used by DynamiCrafter")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        k, pad = (3, 1, 1), (1, 0, 0)
        self.conv1 = nn.Sequential(nn.GroupNorm(32, in_channels), nn.SiLU(),
                                   nn.Conv3d(in_channels, out_channels, k, padding=pad))
        self.conv2 = nn.Sequential(nn.GroupNorm(32, out_channels), nn.SiLU(), nn.Dropout(dropout),
                                   nn.Conv3d(out_channels, in_channels, k, padding=pad))
        self.conv3 = nn.Sequential(nn.GroupNorm(32, out_channels), nn.SiLU(), nn.Dropout(dropout),
                                   nn.Conv3d(out_channels, in_channels, k, padding=pad))
        self.conv4 = nn.Sequential(nn.GroupNorm(32, out_channels), nn.SiLU(), nn.Dropout(dropout),
                                   nn.Conv3d(out_channels, in_channels, k, padding=pad))
        nn.init.zeros_(self.conv4[-1].weight)         # identity at init (openaimodel3d.py:269-270)
        nn.init.zeros_(self.conv4[-1].bias)

    def _stages(self):
        return [self.conv1, self.conv2, self.conv3, self.conv4]

    def _pack(self):
        return {"w": [pack_tconv3(s[-1].weight) for s in self._stages()],
                "b": [_f32(s[-1].bias) for s in self._stages()]}

    def _pre_norm(self, h, colsums=None):
        """GroupNorm + SiLU of the first stage (openaimodel3d.py:255-257) -> operand rows."""
        gn = self.conv1[0]
        return ops.group_norm(h.x, h.B, h.T * h.H * h.W, gn.weight, gn.bias, gn.eps, silu=True, colsums=colsums)

    def _core(self, y, B, T, S, residual, delta=False):
        """The four temporal convs on rows (b, t, pixel), starting from the normalised input of stage 1; every conv emits
        the column sums the next GroupNorm needs.  delta=True: the block output WITHOUT the identity, as bf16
        (frame-sharded mode)."""
        p = self.packed()
        z = colsums = None
        for i, st in enumerate(self._stages()):
            if i > 0:
                gn = st[0]
                y = ops.group_norm(z, B, T * S, gn.weight, gn.bias, gn.eps, silu=True, colsums=colsums)
            last = i == 3
            lowp = (_packing.BF16_TCONV and not last) or (last and delta and not ops.fp32_mode())
            if last and delta:             # nobody normalises the delta: no statistics
                z, colsums = ops.tconv3(y, B, T, S, p["w"][i], bias=p["b"][i], f32=not lowp, bf16=lowp), None
            else:
                z, colsums = ops.tconv3(y, B, T, S, p["w"][i], bias=p["b"][i], residual=residual if last else None,
                                        f32=not lowp, bf16=lowp, stats=True)
        return z, colsums

    def _fwd(self, h, colsums=None):
        """colsums: fused statistics of h.x from the GEMM that produced it, or None."""
        z, cs = self._core(self._pre_norm(h, colsums), h.B, h.T, h.H * h.W, h.x)
        return h.like(z, cs=cs)

    def forward(self, x):
        """x: [b, c, t, h, w] (openaimodel3d.py:272)."""
        b, c, t, hh, ww = x.shape
        rows = x.float().permute(0, 2, 3, 4, 1).reshape(-1, c).contiguous()
        o = self._fwd(Act(rows, b, t, hh, ww))
        return o.x.reshape(b, t, hh, ww, c).permute(0, 4, 1, 2, 3).contiguous()


class ResBlock(PackedModule, TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_scale_shift_norm=False, dims=2,
                 use_checkpoint=False, use_conv=False, up=False, down=False, use_temporal_conv=False,
                 tempspatial_aware=False):
        super().__init__()
        if use_scale_shift_norm or up or down or dims != 2 or use_conv:
            raise NotImplementedError("scale-shift norm / resblock up-down are not used by DynamiCrafter")
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint = use_conv, use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.use_temporal_conv = use_temporal_conv
        self.updown = False
        self.in_layers = nn.Sequential(nn.GroupNorm(32, channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1))
        for p in self.out_layers[-1].parameters():    # zero_module (openaimodel3d.py:179)
            nn.init.zeros_(p)
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)
        if self.use_temporal_conv:
            self.temopral_conv = TemporalConvBlock(self.out_channels, self.out_channels, dropout=0.1,
                                                   spatial_aware=tempspatial_aware)

    def _pack(self):
        p = {"w1": pack_conv3x3(self.in_layers[2].weight), "b1": _f32(self.in_layers[2].bias),
             "w2": pack_conv3x3(self.out_layers[3].weight), "b2": _f32(self.out_layers[3].bias),
             "we": _f32(self.emb_layers[1].weight), "be": _f32(self.emb_layers[1].bias)}
        if not isinstance(self.skip_connection, nn.Identity):
            p["ws"] = pack_linear(self.skip_connection.weight)
            p["bs"] = _f32(self.skip_connection.bias)
        return p

    def _fwd(self, h, emb_out, rows_per_emb, skip=None, temporal=True):
        """h: Act; skip: optional fp32 rows concatenated on channels (openaimodel3d.py:596);
        emb_out: fp32 [groups, Cout] = emb_layers(emb), one row per `rows_per_emb` activation rows."""
        p = self.packed()
        nf, S = h.B * h.T, h.H * h.W
        g1 = self.in_layers[0]
        has_skip_conv = "ws" in p
        # skip: fp32 rows, or the Act of the stored skip connection (then its producer's statistics serve the GroupNorm of
        # the concatenation where the group boundaries line up)
        skip_cs = getattr(skip, "cs", None)
        skip = getattr(skip, "x", skip)
        res = ops.group_norm(h.x, nf, S, g1.weight, g1.bias, g1.eps, x2=skip, silu=True, want_raw=has_skip_conv,
                             colsums=h.cs if (skip is None or skip_cs is not None) else None, colsums2=skip_cs)
        y, raw = res if has_skip_conv else (res, None)
        # each conv also emits the column sums its consumer GroupNorm needs (fused statistics, no extra pass)
        h1, cs1 = ops.conv3x3(y, nf, h.H, h.W, p["w1"], bias=p["b1"], group_bias=emb_out, rows_per_group=rows_per_emb,
                              f32=not _packing.BF16_RES, bf16=_packing.BF16_RES, stats=True)
        g2 = self.out_layers[0]
        y2 = ops.group_norm(h1, nf, S, g2.weight, g2.bias, g2.eps, silu=True, colsums=cs1)
        xs = ops.linear(raw, p["ws"], bias=p["bs"], f32=True, bf16=False) if has_skip_conv else h.x
        h2, cs2 = ops.conv3x3(y2, nf, h.H, h.W, p["w2"], bias=p["b2"], residual=xs, stats=True)
        out = h.like(h2, cs=cs2)
        if self.use_temporal_conv and temporal:
            out = self.temopral_conv._fwd(out, colsums=cs2)
        return out

    def forward(self, x, emb, batch_size=None):
        """x: [(b t), c, h, w]; emb: [(b t), emb_channels] (openaimodel3d.py:197-236)."""
        n, c, hh, ww = x.shape
        p = self.packed()
        emb_out = ops.small_linear(emb.float().contiguous(), p["we"], p["be"], silu_in=True)
        b = batch_size if batch_size else n
        act = Act(_rows_from_nchw(x), b, n // b, hh, ww)
        o = self._fwd(act, emb_out, hh * ww, temporal=bool(batch_size))
        return _nchw_from_rows(o.x, n, hh, ww)


class UNetModel(PackedModule):
    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0.0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, context_dim=None,
                 use_scale_shift_norm=False, resblock_updown=False, num_heads=-1, num_head_channels=-1,
                 transformer_depth=1, use_linear=False, use_checkpoint=False, temporal_conv=False,
                 tempspatial_aware=False, temporal_attention=True, use_relative_position=True,
                 use_causal_attention=False, temporal_length=None, use_fp16=False, addition_attention=False,
                 temporal_selfatt_only=True, image_cross_attention=False,
                 image_cross_attention_scale_learnable=False, default_fs=4, fs_condition=False):
        super().__init__()
        if num_heads == -1:
            assert num_head_channels != -1, 'Either num_heads or num_head_channels has to be set'
        if num_head_channels == -1:
            assert num_heads != -1, 'Either num_heads or num_head_channels has to be set'
        if resblock_updown or not conv_resample or dims != 2:
            raise NotImplementedError("resblock_updown / pooling resample are not used by DynamiCrafter")
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, attention_resolutions
        self.dropout, self.channel_mult, self.conv_resample = dropout, channel_mult, conv_resample
        self.temporal_attention = temporal_attention
        time_embed_dim = model_channels * 4
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.addition_attention = addition_attention
        self.temporal_length = temporal_length
        self.image_cross_attention = image_cross_attention
        self.image_cross_attention_scale_learnable = image_cross_attention_scale_learnable
        self.default_fs, self.fs_condition = default_fs, fs_condition

        def mlp():
            return nn.Sequential(nn.Linear(model_channels, time_embed_dim), nn.SiLU(),
                                 nn.Linear(time_embed_dim, time_embed_dim))
        self.time_embed = mlp()
        if fs_condition:
            self.fps_embedding = mlp()
            nn.init.zeros_(self.fps_embedding[-1].weight)
            nn.init.zeros_(self.fps_embedding[-1].bias)

        def res(cin, cout):
            return ResBlock(cin, time_embed_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, tempspatial_aware=tempspatial_aware,
                            use_temporal_conv=temporal_conv)

        def attn(ch):
            if num_head_channels == -1:
                heads, dim_head = num_heads, ch // num_heads
            else:
                heads, dim_head = ch // num_head_channels, num_head_channels
            layers = [SpatialTransformer(ch, heads, dim_head, depth=transformer_depth, context_dim=context_dim,
                                         use_linear=use_linear, use_checkpoint=use_checkpoint, disable_self_attn=False,
                                         video_length=temporal_length, image_cross_attention=image_cross_attention,
                                         image_cross_attention_scale_learnable=image_cross_attention_scale_learnable)]
            if temporal_attention:
                layers.append(TemporalTransformer(ch, heads, dim_head, depth=transformer_depth, context_dim=context_dim,
                                                  use_linear=use_linear, use_checkpoint=use_checkpoint,
                                                  only_self_att=True, causal_attention=use_causal_attention,
                                                  relative_position=use_relative_position,
                                                  temporal_length=temporal_length))
            return layers

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        if addition_attention:
            self.init_attn = TimestepEmbedSequential(
                TemporalTransformer(model_channels, n_heads=8, d_head=num_head_channels, depth=transformer_depth,
                                    context_dim=context_dim, use_checkpoint=use_checkpoint,
                                    only_self_att=temporal_selfatt_only, causal_attention=False,
                                    relative_position=use_relative_position, temporal_length=temporal_length))
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers += attn(ch)
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), *attn(ch), res(ch, ch))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [res(ch + ich, mult * model_channels)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers += attn(ch)
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))
        for p in self.out[-1].parameters():           # zero_module (openaimodel3d.py:545)
            nn.init.zeros_(p)

    # ------------------------------------------------------------------------------------------
    def _resblocks(self):
        return [m for m in self.modules() if isinstance(m, ResBlock)]

    def _pack(self):
        rbs = self._resblocks()
        offs, o = [], 0
        for rb in rbs:
            offs.append(o)
            o += rb.out_channels
        p = {"w_in": pack_conv3x3(self.input_blocks[0][0].weight), "b_in": _f32(self.input_blocks[0][0].bias),
             "w_out": pack_conv3x3(self.out[2].weight), "b_out": _f32(self.out[2].bias),
             "te": [_f32(self.time_embed[0].weight), _f32(self.time_embed[0].bias),
                    _f32(self.time_embed[2].weight), _f32(self.time_embed[2].bias)],
             # all ResBlock emb_layers as ONE [sum(Cout), 4*mc] matrix (one launch per forward, not 22)
             "we_all": torch.cat([_f32(rb.emb_layers[1].weight) for rb in rbs], 0).contiguous(),
             "be_all": torch.cat([_f32(rb.emb_layers[1].bias) for rb in rbs], 0).contiguous(),
             "emb_off": {id(rb): (off, rb.out_channels) for rb, off in zip(rbs, offs)}}
        if self.fs_condition:
            p["fe"] = [_f32(self.fps_embedding[0].weight), _f32(self.fps_embedding[0].bias),
                       _f32(self.fps_embedding[2].weight), _f32(self.fps_embedding[2].bias)]
        return p

    def _all_params(self):
        """Flat parameter list, built once: walking the module tree (1500 tensors, 700 modules) costs ~1 ms of host time,
        which an end-to-end step pays with the GPU idle.  Dropped whenever the module is converted / moved."""
        lst = self.__dict__.get("_param_list")
        if lst is None:
            lst = self.__dict__["_param_list"] = (list(self.parameters()),
                                                   [q for a in self._cross_layers() for q in a.parameters()])
        return lst

    def _apply(self, fn, *a, **kw):
        self.__dict__.pop("_param_list", None)
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self.__dict__.pop("_param_list", None)
        return super().load_state_dict(*a, **kw)

    def _own_version(self):
        return sum([q._version for q in self._all_params()[0]])

    def _context(self, context, b, t):
        """openaimodel3d.py:555-562 without replicating anything: text rows are shared by the t frames of a
        video (kv batch = frame // t), image tokens are already per frame in memory."""
        L = context.shape[1]
        if L == TEXT_CONTEXT_LEN + t * IMG_TOKENS_PER_FRAME:
            ctx = context.float()
            D = ctx.shape[2]
            text = ops.cast_bf16(ctx[:, :TEXT_CONTEXT_LEN].reshape(b * TEXT_CONTEXT_LEN, D).contiguous())
            image = ops.cast_bf16(ctx[:, TEXT_CONTEXT_LEN:].reshape(b * t * IMG_TOKENS_PER_FRAME, D).contiguous())
            return CrossAttentionContext(text, TEXT_CONTEXT_LEN, t, image, IMG_TOKENS_PER_FRAME, 1)
        return make_context(context, t, self.image_cross_attention)

    @staticmethod
    def _temporal(mod, h, shard):
        """Frame-sharded mode: a temporal block needs every frame of a pixel.  Its GroupNorm runs on this rank's frames
        (statistics all-reduced), the normalised bf16 operand is transposed frames -> pixels (all-to-all), the block core
        runs on this rank's pixel columns with all frames and returns its DELTA (no residual) as bf16, which is
        transposed back and added to the fp32 residual stream by the unpack kernel.  Only bf16 crosses
        NVLink; DCB_SHARD_V1=1 restores the round-1 flow (fp32 residual stream transposed both ways)."""
        from dynamicrafter_b200 import parallel
        hw = h.H * h.W
        if _packing.SHARD_V1:
            hp = Act(parallel.frames_to_pixels(h.x, h.B, h.T, hw, shard), h.B, h.T * shard.world, 1, hw // shard.world)
            with ops.sharded_group_norm(shard):
                op = mod._fwd(hp)
            return h.like(parallel.pixels_to_frames(op.x, h.B, hp.T, hp.W, shard))
        with ops.sharded_group_norm(shard):
            y = mod._pre_norm(h)                                     # frame side: rows (b, t_loc, p)
        yp = parallel.frames_to_pixels_v2(y, h.B, h.T, hw, shard)    # pixel side: rows (b, t, p_loc)
        with ops.sharded_group_norm(shard):
            delta, _ = mod._core(yp, h.B, h.T * shard.world, hw // shard.world, None, delta=True)
        return h.like(parallel.pixels_to_frames_add(delta, h.x, h.B, h.T * shard.world, hw // shard.world, shard))

    def _run(self, layers, h, emb_all, pk, ctx, skip=None, shard=None):
        for layer in layers:
            if isinstance(layer, ResBlock):
                off, n = pk["emb_off"][id(layer)]
                h = layer._fwd(h, emb_all[:, off:off + n], h.T * h.H * h.W, skip=skip, temporal=shard is None)
                if shard is not None and layer.use_temporal_conv:
                    h = self._temporal(layer.temopral_conv, h, shard)
                skip = None
            elif isinstance(layer, SpatialTransformer):
                h = layer._fwd(h, ctx)
            elif isinstance(layer, TemporalTransformer):
                h = layer._fwd(h) if shard is None else self._temporal(layer, h, shard)
            elif isinstance(layer, (Downsample, Upsample)):
                h = layer._fwd(h)
            else:
                raise TypeError(type(layer))
        return h

    @torch.no_grad()
    def forward(self, x, timesteps, context=None, features_adapter=None, fs=None, **kwargs):
        """x: [b, in_channels, t, h, w]; timesteps: [b] long; context: [b, 77 + 16 t, context_dim];
        fs: [b] long or None (-> default_fs).  Extra kwargs are accepted and ignored like the reference."""
        if features_adapter is not None:
            raise NotImplementedError("features_adapter is not used on the DynamiCrafter inference path")
        if context is None:
            raise ValueError("UNetModel.forward needs the cross-attention context")
        if not x.is_cuda:
            raise RuntimeError("dcb200 UNetModel runs on CUDA tensors only (no CPU path)")
        dev = x.device
        b = x.shape[0]
        timesteps = timesteps.to(device=dev, dtype=torch.long)
        if self.fs_condition:
            fs = torch.full((b,), self.default_fs, dtype=torch.long, device=dev) if fs is None \
                else fs.to(device=dev, dtype=torch.long)
        context = context.to(device=dev, dtype=torch.float32)
        if not getattr(self, "use_cuda_graph", False) or \
                (getattr(self, "frame_shard", None) is not None and not _packing.GRAPH_COLLECTIVES):
            return self._forward_impl(x, timesteps, context, fs)
        return self._forward_graphed(x, timesteps, context, fs)

    # ---- cross-attention context: cast once, K/V of every cross-attention layer projected once per context tensor
    def _cross_layers(self):
        return [m.transformer_blocks[0].attn2 for m in self.modules() if isinstance(m, SpatialTransformer)]

    def _build_ctx(self, context, b, t_full, shard):
        """context [b, L, D] -> CrossAttentionContext with the text / image keys and values of all cross-attention layers
        (attention.py:86-94: k, v depend on the context only, not on x or the timestep -> step invariant)."""
        t = t_full
        if shard is not None:
            t0, t1 = shard.frames(t_full)
            t = t1 - t0
            if context.shape[1] == TEXT_CONTEXT_LEN + t_full * IMG_TOKENS_PER_FRAME:
                i0 = TEXT_CONTEXT_LEN + t0 * IMG_TOKENS_PER_FRAME
                i1 = TEXT_CONTEXT_LEN + t1 * IMG_TOKENS_PER_FRAME
                context = torch.cat([context[:, :TEXT_CONTEXT_LEN], context[:, i0:i1]], 1)
        ctx = self._context(context, b, t)
        if _packing.CACHE_CROSS_KV:
            ctx.kv = {id(a): a._project_context(ctx) for a in self._cross_layers()}
        return ctx

    def _ctx_key(self, context):
        return (context._version, tuple(context.shape), self.__dict__["_packed_cache"][0],
                sum([q._version for q in self._all_params()[1]]))

    def _cached_ctx(self, context, b, t_full, shard):
        """Eager path: reuse the projected context while the caller passes the SAME tensor object (unchanged version) --
        the DDIM loop does for all its steps.  The cache keeps the tensor alive, so its address cannot be recycled."""
        self.packed()
        if not _packing.CACHE_CROSS_KV:
            return self._build_ctx(context, b, t_full, shard)
        ent = self.__dict__.get("_ctx_cache")
        key = self._ctx_key(context) + (None if shard is None else (shard.rank, shard.world),)
        if ent is None or ent[0] is not context or ent[1] != key:
            ent = (context, key, self._build_ctx(context, b, t_full, shard))
            self.__dict__["_ctx_cache"] = ent
        return ent[2]

    def _forward_graphed(self, x, timesteps, context, fs):
        """Replay CUDA graphs of the whole forward (static shapes; ~800 kernel launches collapse into one graph launch).
        Captured lazily per (shapes, weight version); inputs are copied into static buffers.  Two graphs: the context
        graph (casts + the K/V projections of all cross-attention layers) is replayed only when a new context tensor
        arrives, the main graph every call."""
        self.packed()
        shard = getattr(self, "frame_shard", None)
        key = (tuple(x.shape), tuple(context.shape), x.dtype, self.__dict__["_packed_cache"][0],
               None if shard is None else (id(shard), shard.rank, shard.world))
        cache = self.__dict__.setdefault("_graphs", {})
        ent = cache.get(key)
        if ent is None:
            sx, st, sc = x.clone(), timesteps.clone(), context.clone()
            sf = None if fs is None else fs.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                         # warm-up: lazy packing, func attributes, scratch
                    self._forward_body(sx, st, self._build_ctx(sc, x.shape[0], x.shape[2], shard), sf)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            cgraph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cgraph):
                sctx = self._build_ctx(sc, x.shape[0], x.shape[2], shard)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                sy = self._forward_body(sx, st, sctx, sf)
            if len(cache) >= 4:
                cache.clear()
            ent = cache[key] = [graph, sx, st, sc, sf, sy, cgraph, sctx, None, None]
        graph, sx, st, sc, sf, sy, cgraph = ent[:7]
        sx.copy_(x)
        st.copy_(timesteps)
        ckey = self._ctx_key(context)
        if not _packing.CACHE_CROSS_KV or ent[8] is not context or ent[9] != ckey:
            sc.copy_(context)
            cgraph.replay()
            ent[8], ent[9] = context, ckey                 # keeps the tensor alive: its address cannot be recycled
        if sf is not None:
            sf.copy_(fs)
        graph.replay()
        return sy.clone()

    def _forward_impl(self, x, timesteps, context, fs):
        shard = getattr(self, "frame_shard", None)
        ctx = self._cached_ctx(context, x.shape[0], x.shape[2], shard)
        return self._forward_body(x, timesteps, ctx, fs)

    def _forward_body(self, x, timesteps, ctx, fs):
        with ops.stats_scope(x.device):          # fused GroupNorm statistics accumulators: one memset per forward
            return self._forward_steps(x, timesteps, ctx, fs)

    def _forward_steps(self, x, timesteps, ctx, fs):
        pk = self.packed()
        shard = getattr(self, "frame_shard", None)      # parallel.FrameShard: this rank denoises frames t0..t1 only
        if shard is not None:
            t0, t1 = shard.frames(x.shape[2])
            x = x[:, :, t0:t1].contiguous()
        b, cin, t, hh, ww = x.shape
        mc = self.model_channels
        dev = x.device
        # ---- embeddings (fp32, tiny): emb = time_embed(sin(t)) + fps_embedding(sin(fs))   (:550-577)
        te = pk["te"]
        e1 = ops.small_linear(ops.timestep_embedding(timesteps.to(dev), mc), te[0], te[1], silu_out=True)
        emb = ops.small_linear(e1, te[2], te[3])
        if self.fs_condition:
            if fs is None:
                fs = torch.full((b,), self.default_fs, dtype=torch.long, device=dev)
            fe = pk["fe"]
            f1 = ops.small_linear(ops.timestep_embedding(fs.to(dev), mc), fe[0], fe[1], silu_out=True)
            ops.small_linear(f1, fe[2], fe[3], out=emb, accumulate=True)
        emb_all = ops.small_linear(emb, pk["we_all"], pk["be_all"], silu_in=True)      # [b, sum(Cout)]
        # ---- input conv
        xr = ops.nchw_to_rows(x, 64 if cin <= 64 else (cin + 63) // 64 * 64)
        h0, cs0 = ops.conv3x3(xr, b * t, hh, ww, pk["w_in"], bias=pk["b_in"], stats=True)
        h = Act(h0, b, t, hh, ww, cs0)
        hs = []
        for i, module in enumerate(self.input_blocks):
            if i > 0:
                h = self._run(module, h, emb_all, pk, ctx, shard=shard)
            if i == 0 and self.addition_attention:
                h = self.init_attn[0]._fwd(h) if shard is None else self._temporal(self.init_attn[0], h, shard)
            hs.append(h)
        h = self._run(self.middle_block, h, emb_all, pk, ctx, shard=shard)
        for module in self.output_blocks:
            h = self._run(module, h, emb_all, pk, ctx, skip=hs.pop(), shard=shard)
        g = self.out[0]
        y = ops.group_norm(h.x, b * t, hh * ww, g.weight, g.bias, g.eps, silu=True, colsums=h.cs)
        o = ops.conv3x3(y, b * t, hh, ww, pk["w_out"], bias=pk["b_out"])
        out = ops.rows_to_nchw(o, b, self.out_channels, t, hh, ww)
        if shard is not None:
            from dynamicrafter_b200.parallel import all_gather_frames
            out = all_gather_frames(out, shard)          # every rank applies the (tiny) DDIM update to all frames
        return out.to(x.dtype)
