"""Drop-in mirrors of lvdm/modules/attention.py (reference) running on the dcb200 CUDA kernels.

Same constructor arguments, ``state_dict`` keys and forward signatures as the reference classes
(CrossAttention attention.py:42-144, BasicTransformerBlock :212-246, SpatialTransformer :249-310,
TemporalTransformer :313-412, GEGLU :415-422, FeedForward :425-442); the torch ``nn.Linear`` /
``nn.LayerNorm`` / ``nn.GroupNorm`` children only hold parameters -- every FLOP runs in libdcb200.so.

Internally activations stay channels-last ([rows, C], rows = (b, t, y, x)) so the reference's
'(b t) c h w <-> b (h w) c' and 'b c t h w <-> (b h w) t c' rearranges (attention.py:300,307,369-372,
391,406) do not exist: per-token ops are order independent and the temporal attention kernel reads
its 16 frames with a stride.
"""
import torch
import torch.nn as nn

from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm._packing import PackedModule, Act, pack_linear, pack_geglu

TEXT_CONTEXT_LEN = 77


def _default(v, d):
    return d if v is None else v


def _fold_ln_weights(norm, w, b):
    """LayerNorm affine folded into the Linear behind it: W'[n,k] = W[n,k] gamma[k], b'[n] = b[n] + sum_k W[n,k] beta[k]."""
    w = w.detach().float()
    wf = w * norm.weight.detach().float()[None, :]
    bf = w @ norm.bias.detach().float()
    if b is not None:
        bf = bf + b.detach().float()
    return wf, bf


def fold_layer_norm(norm, w, b):
    """-> ops.LnFold: bf16 gamma-scaled weights, folded bias, and the row sums of the weights AS ROUNDED to bf16 (what the
    tensor core multiplies with), so that subtracting mean * rowsum removes the mean exactly."""
    wf, bf = _fold_ln_weights(norm, w, b)
    wq = wf.to(torch.bfloat16).contiguous()
    return ops.LnFold(wq, bf.contiguous(), wq.float().sum(1).contiguous(), float(norm.eps))


class CrossAttentionContext:
    """Per-forward conditioning in kernel layout: bf16 token matrices + how frames map onto them."""

    def __init__(self, text, n_text, text_div, image=None, n_image=0, image_div=1):
        self.text, self.n_text, self.text_div = text, n_text, text_div
        self.image, self.n_image, self.image_div = image, n_image, image_div
        self.kv = None          # {id(CrossAttention): (kv_text, kv_image)}: projected once per context (UNetModel._build_ctx)


class CrossAttention(PackedModule):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.,
                 relative_position=False, temporal_length=None, video_length=None, image_cross_attention=False,
                 image_cross_attention_scale=1.0, image_cross_attention_scale_learnable=False, text_context_len=77):
        super().__init__()
        if relative_position:
            raise NotImplementedError("relative_position attention is not used by the DynamiCrafter configs")
        if dim_head != 64:
            raise NotImplementedError("dcb200 attention kernels are specialised for dim_head == 64")
        inner_dim = dim_head * heads
        self.is_self = context_dim is None
        context_dim = _default(context_dim, query_dim)
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head, self.inner_dim = heads, dim_head, inner_dim
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        self.relative_position = False
        self.temporal_length = temporal_length
        self.video_length = video_length
        self.image_cross_attention = image_cross_attention
        self.image_cross_attention_scale = image_cross_attention_scale
        self.text_context_len = text_context_len
        self.image_cross_attention_scale_learnable = image_cross_attention_scale_learnable
        if image_cross_attention:
            self.to_k_ip = nn.Linear(context_dim, inner_dim, bias=False)
            self.to_v_ip = nn.Linear(context_dim, inner_dim, bias=False)
            if image_cross_attention_scale_learnable:
                self.register_parameter('alpha', nn.Parameter(torch.tensor(0.)))

    def _pack(self):
        p = {"wo": pack_linear(self.to_out[0].weight), "bo": self.to_out[0].bias.detach().float().contiguous()}
        if self.is_self:
            p["wqkv"] = torch.cat([pack_linear(self.to_q.weight), pack_linear(self.to_k.weight),
                                   pack_linear(self.to_v.weight)], 0).contiguous()
        else:
            p["wq"] = pack_linear(self.to_q.weight)
            p["wkv"] = torch.cat([pack_linear(self.to_k.weight), pack_linear(self.to_v.weight)], 0).contiguous()
            if self.image_cross_attention:
                p["wkv_ip"] = torch.cat([pack_linear(self.to_k_ip.weight), pack_linear(self.to_v_ip.weight)], 0).contiguous()
                s = float(self.image_cross_attention_scale)
                if self.image_cross_attention_scale_learnable:
                    s = s * (float(torch.tanh(self.alpha.detach().float())) + 1.0)     # attention.py:139-140
                p["ip_scale"] = s
        return p

    # ---- kernel-layout entry points -------------------------------------------------------------
    def _attend_self_spatial(self, xn, nbatch, ntok, residual, out_bf16=False, ln=None, ln_stats=False):
        """xn: bf16 [nbatch*ntok, C] (already layer-normed), or an ops.LnSrc with ln = the folded LayerNorm.
        Returns to_out(attn) + residual (fp32); with ln_stats also the LnSrc of that result for the next LayerNorm."""
        p = self.packed()
        inner = self.inner_dim
        qkv = ops.linear(xn, p["wqkv"]) if ln is None else ops.linear(xn, None, ln=ln)
        a = ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], ntok, 1, 1.0)],
                                nbatch, self.heads, ntok, self.scale)
        return ops.linear(a, p["wo"], bias=p["bo"], residual=residual, f32=not out_bf16, bf16=out_bf16, ln_stats=ln_stats)

    def _attend_self_temporal(self, xn, B, T, HW, residual, ln=None, ln_stats=False):
        p = self.packed()
        inner = self.inner_dim
        qkv = ops.linear(xn, p["wqkv"]) if ln is None else ops.linear(xn, None, ln=ln)
        a = ops.temporal_attention(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], B, T, HW,
                                   self.heads, self.scale)
        return ops.linear(a, p["wo"], bias=p["bo"], residual=residual, f32=True, bf16=False, ln_stats=ln_stats)

    def _attend_cross(self, xn, nbatch, ntok, ctx, residual, ln=None, ln_stats=False):
        """ctx: CrossAttentionContext.  Text and image keys get independent softmaxes (attention.py:129-142)."""
        p = self.packed()
        inner = self.inner_dim
        q = ops.linear(xn, p["wq"]) if ln is None else ops.linear(xn, None, ln=ln)
        kv, kv_ip = ctx.kv[id(self)] if (ctx.kv is not None and id(self) in ctx.kv) else self._project_context(ctx)
        segs = [(kv[:, :inner], kv[:, inner:], ctx.n_text, ctx.text_div, 1.0)]
        if kv_ip is not None:
            segs.append((kv_ip[:, :inner], kv_ip[:, inner:], ctx.n_image, ctx.image_div, p["ip_scale"]))
        a = ops.flash_attention(q, segs, nbatch, self.heads, ntok, self.scale)
        return ops.linear(a, p["wo"], bias=p["bo"], residual=residual, f32=True, bf16=False, ln_stats=ln_stats)

    def _project_context(self, ctx):
        """(to_k | to_v)(text) and (to_k_ip | to_v_ip)(image) of this layer (attention.py:86-94): functions of the context
        only, so UNetModel projects them once per context tensor instead of once per denoising step."""
        p = self.packed()
        kv = ops.linear(ctx.text, p["wkv"])
        kv_ip = None
        if self.image_cross_attention and ctx.image is not None and ctx.n_image > 0:
            kv_ip = ops.linear(ctx.image, p["wkv_ip"])
        return kv, kv_ip

    def _ln_fold(self, norm):
        """LnFold of `norm` followed by this attention's query-side projection (fused q/k/v for self-attention)."""
        if self.is_self:
            w = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0)
        else:
            w = self.to_q.weight
        return fold_layer_norm(norm, w, None)

    # ---- reference signature ----------------------------------------------------------------------
    def forward(self, x, context=None, mask=None):
        """x: [b, n, query_dim] fp32; context: [b, m, context_dim] or None (self-attention)."""
        if mask is not None:
            raise NotImplementedError("attention masks are not used on the DynamiCrafter inference path")
        b, n, c = x.shape
        xn = ops.cast_bf16(x.reshape(b * n, c).float().contiguous())
        zero = None
        if context is None:
            if self.temporal_length is not None and n <= 16:
                # '(b hw) t c' caller layout: rows (bhw, t) -> the temporal kernel wants (b, t, hw) with hw=1
                y = self._attend_self_temporal(xn, b, n, 1, zero)
            else:
                y = self._attend_self_spatial(xn, b, n, zero)
        else:
            y = self._attend_cross(xn, b, n, make_context(context, 1, self.image_cross_attention,
                                                          self.text_context_len), zero)
        return y.reshape(b, n, -1)


def make_context(context, frames_per_batch, image_cross_attention=True, text_len=TEXT_CONTEXT_LEN):
    """[nb, L, D] per-batch context -> CrossAttentionContext shared by `frames_per_batch` attention batches
    (attention.py:89-98: text = first 77 tokens, image = the rest)."""
    nb, L, D = context.shape
    ctx = context.float()
    nt = min(L, text_len)
    text = ops.cast_bf16(ctx[:, :nt].reshape(nb * nt, D).contiguous())
    image, ni = None, 0
    if image_cross_attention and L > text_len:
        ni = L - text_len
        image = ops.cast_bf16(ctx[:, text_len:].reshape(nb * ni, D).contiguous())
    return CrossAttentionContext(text, nt, frames_per_batch, image, ni, frames_per_batch)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(PackedModule):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        if not glu:
            raise NotImplementedError("only the GEGLU feed-forward is used by DynamiCrafter")
        inner_dim = int(dim * mult)
        dim_out = _default(dim_out, dim)
        self.net = nn.Sequential(GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def _pack(self):
        w1, b1, bn = pack_geglu(self.net[0].proj.weight, self.net[0].proj.bias)
        return {"w1": w1, "b1": b1, "bn": bn, "w2": pack_linear(self.net[2].weight),
                "b2": self.net[2].bias.detach().float().contiguous()}

    def _fwd(self, xn, residual, out_bf16, ln=None):
        p = self.packed()
        if ln is None:
            g = ops.linear(xn, p["w1"], bias=p["b1"], geglu=True, block_n=p["bn"])
        else:
            g = ops.linear(xn, None, geglu=True, block_n=p["bn"], ln=ln)
        return ops.linear(g, p["w2"], bias=p["b2"], residual=residual, f32=not out_bf16, bf16=out_bf16)

    def _ln_fold(self, norm):
        """LnFold of `norm` followed by the GEGLU projection (rows interleaved per output tile like w1)."""
        proj = self.net[0].proj
        wf, bf = _fold_ln_weights(norm, proj.weight, proj.bias)
        w1, b1, _ = pack_geglu(wf, bf)
        return ops.LnFold(w1, b1, w1.float().sum(1).contiguous(), float(norm.eps))

    def forward(self, x):
        shp = x.shape
        xn = ops.cast_bf16(x.reshape(-1, shp[-1]).float().contiguous())
        return self._fwd(xn, None, False).reshape(*shp[:-1], -1)


class BasicTransformerBlock(PackedModule):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True,
                 disable_self_attn=False, attention_cls=None, video_length=None, image_cross_attention=False,
                 image_cross_attention_scale=1.0, image_cross_attention_scale_learnable=False, text_context_len=77):
        super().__init__()
        attn_cls = CrossAttention if attention_cls is None else attention_cls
        if disable_self_attn:
            raise NotImplementedError("disable_self_attn is not used by DynamiCrafter")
        self.disable_self_attn = disable_self_attn
        self.attn1 = attn_cls(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout, context_dim=None)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = attn_cls(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout,
                              video_length=video_length, image_cross_attention=image_cross_attention,
                              image_cross_attention_scale=image_cross_attention_scale,
                              image_cross_attention_scale_learnable=image_cross_attention_scale_learnable,
                              text_context_len=text_context_len)
        self.image_cross_attention = image_cross_attention
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint

    def _pack(self):
        """Folded LayerNorms (bf16 mode): norm1 -> attn1 q/k/v, norm2 -> attn2 q (or q/k/v), norm3 -> GEGLU."""
        if not ops.ln_fold_enabled():
            return {}
        return {"ln1": self.attn1._ln_fold(self.norm1), "ln2": self.attn2._ln_fold(self.norm2),
                "ln3": self.ff._ln_fold(self.norm3)}

    def _own_version(self):
        lst = self.__dict__.get("_allp")                           # the folds read the attention / FF weights too
        if lst is None:
            lst = self.__dict__["_allp"] = list(self.parameters())
        return sum([q._version for q in lst])

    def _ln(self, norm, x):
        return ops.layer_norm(x, norm.weight, norm.bias, norm.eps)

    def _fwd_spatial(self, x, nbatch, ntok, ctx, ls=None):
        """x: fp32 [nbatch*ntok, dim] residual stream -> bf16 (the only consumer is proj_out).  ls: ops.LnSrc of x from
        the GEMM that produced it -> the three LayerNorms are folded into their neighbour GEMMs (no layer_norm launch)."""
        if ls is not None:
            p = self.packed()
            x, ls = self.attn1._attend_self_spatial(ls, nbatch, ntok, x, ln=p["ln1"], ln_stats=True)
            x, ls = self.attn2._attend_cross(ls, nbatch, ntok, ctx, x, ln=p["ln2"], ln_stats=True)
            return self.ff._fwd(ls, x, True, ln=p["ln3"])
        x = self.attn1._attend_self_spatial(self._ln(self.norm1, x), nbatch, ntok, x)
        x = self.attn2._attend_cross(self._ln(self.norm2, x), nbatch, ntok, ctx, x)
        return self.ff._fwd(self._ln(self.norm3, x), x, True)

    def _fwd_temporal(self, x, B, T, HW, ls=None):
        if ls is not None:
            p = self.packed()
            x, ls = self.attn1._attend_self_temporal(ls, B, T, HW, x, ln=p["ln1"], ln_stats=True)
            x, ls = self.attn2._attend_self_temporal(ls, B, T, HW, x, ln=p["ln2"], ln_stats=True)
            return self.ff._fwd(ls, x, True, ln=p["ln3"])
        x = self.attn1._attend_self_temporal(self._ln(self.norm1, x), B, T, HW, x)
        x = self.attn2._attend_self_temporal(self._ln(self.norm2, x), B, T, HW, x)
        return self.ff._fwd(self._ln(self.norm3, x), x, True)

    def forward(self, x, context=None, mask=None, **kwargs):
        """Reference signature: x [b, n, dim]; context None => both attentions are self-attention."""
        if mask is not None:
            raise NotImplementedError("attention masks are not used on the DynamiCrafter inference path")
        b, n, c = x.shape
        xr = x.reshape(b * n, c).float().contiguous()
        if context is None:
            if self.attn1.temporal_length is not None:
                y = self._fwd_temporal(xr, b, n, 1)
            else:
                x1 = self.attn1._attend_self_spatial(self._ln(self.norm1, xr), b, n, xr)
                x2 = self.attn2._attend_self_spatial(self._ln(self.norm2, x1), b, n, x1)
                y = self.ff._fwd(self._ln(self.norm3, x2), x2, True)
        else:
            ctx = make_context(context, 1, self.image_cross_attention)
            y = self._fwd_spatial(xr, b, n, ctx)
        return y.float().reshape(b, n, c)


class SpatialTransformer(PackedModule):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None,
                 use_checkpoint=True, disable_self_attn=False, use_linear=False, video_length=None,
                 image_cross_attention=False, image_cross_attention_scale_learnable=False):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        if not use_linear:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        else:
            self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim,
                                  disable_self_attn=disable_self_attn, checkpoint=use_checkpoint,
                                  video_length=video_length, image_cross_attention=image_cross_attention,
                                  image_cross_attention_scale_learnable=image_cross_attention_scale_learnable)
            for _ in range(depth)])
        if not use_linear:
            self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)
        else:
            self.proj_out = nn.Linear(inner_dim, in_channels)
        for p in self.proj_out.parameters():          # zero_module (attention.py:288-290)
            nn.init.zeros_(p)
        self.use_linear = use_linear
        if depth != 1:
            raise NotImplementedError("transformer_depth != 1 is not used by DynamiCrafter")

    def _pack(self):
        return {"w_in": pack_linear(self.proj_in.weight), "b_in": self.proj_in.bias.detach().float().contiguous(),
                "w_out": pack_linear(self.proj_out.weight), "b_out": self.proj_out.bias.detach().float().contiguous()}

    def _fwd(self, h, ctx):
        """h: Act (fp32 rows).  ctx: CrossAttentionContext.  Returns Act."""
        p = self.packed()
        nf, S = h.B * h.T, h.H * h.W
        y = ops.group_norm(h.x, nf, S, self.norm.weight, self.norm.bias, self.norm.eps, silu=False, colsums=h.cs)
        if ops.ln_fold_enabled():
            x, ls = ops.linear(y, p["w_in"], bias=p["b_in"], f32=True, bf16=False, ln_stats=True)
        else:
            x, ls = ops.linear(y, p["w_in"], bias=p["b_in"], f32=True, bf16=False), None
        xb = self.transformer_blocks[0]._fwd_spatial(x, nf, S, ctx, ls)
        out, cs = ops.linear(xb, p["w_out"], bias=p["b_out"], residual=h.x, f32=True, bf16=False, stats_rows=S)
        return h.like(out, cs=cs)

    def forward(self, x, context=None, **kwargs):
        """x: [(b t), c, h, w] fp32; context: [(b t), 77+16, context_dim] (attention.py:294)."""
        n, c, hh, ww = x.shape
        rows = x.float().permute(0, 2, 3, 1).reshape(n * hh * ww, c).contiguous()
        ctx = make_context(context, 1, self.transformer_blocks[0].image_cross_attention)
        out = self._fwd(Act(rows, n, 1, hh, ww), ctx)
        return out.x.reshape(n, hh, ww, c).permute(0, 3, 1, 2).contiguous()


class TemporalTransformer(PackedModule):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None,
                 use_checkpoint=True, use_linear=False, only_self_att=True, causal_attention=False,
                 causal_block_size=1, relative_position=False, temporal_length=None):
        super().__init__()
        if not only_self_att or causal_attention or relative_position:
            raise NotImplementedError("DynamiCrafter uses temporal self-attention without mask / rel-pos only")
        self.only_self_att = only_self_att
        self.relative_position = relative_position
        self.causal_attention = causal_attention
        self.causal_block_size = causal_block_size
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)
        if not use_linear:
            self.proj_in = nn.Conv1d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        else:
            self.proj_in = nn.Linear(in_channels, inner_dim)
        from functools import partial
        attention_cls = partial(CrossAttention, temporal_length=temporal_length)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=None,
                                  attention_cls=attention_cls, checkpoint=use_checkpoint) for _ in range(depth)])
        if not use_linear:
            self.proj_out = nn.Conv1d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)
        else:
            self.proj_out = nn.Linear(inner_dim, in_channels)
        for p in self.proj_out.parameters():          # zero_module (attention.py:360-362)
            nn.init.zeros_(p)
        self.use_linear = use_linear
        if depth != 1:
            raise NotImplementedError("transformer_depth != 1 is not used by DynamiCrafter")

    def _pack(self):
        return {"w_in": pack_linear(self.proj_in.weight), "b_in": self.proj_in.bias.detach().float().contiguous(),
                "w_out": pack_linear(self.proj_out.weight), "b_out": self.proj_out.bias.detach().float().contiguous()}

    def _pre_norm(self, h):
        """GroupNorm over (t, h, w) per video (attention.py:367) -> operand rows."""
        return ops.group_norm(h.x, h.B, h.T * h.H * h.W, self.norm.weight, self.norm.bias, self.norm.eps, silu=False,
                              colsums=h.cs)

    def _core(self, y, B, T, S, residual, delta=False):
        """proj_in -> transformer block -> proj_out on normalised rows (b, t, pixel).  delta=True: the block output WITHOUT
        the residual, as bf16 (frame-sharded mode: the add happens where the fp32 residual stream lives)."""
        p = self.packed()
        if T > 16:
            raise NotImplementedError("temporal attention kernel handles at most 16 frames")
        if ops.ln_fold_enabled():
            x, ls = ops.linear(y, p["w_in"], bias=p["b_in"], f32=True, bf16=False, ln_stats=True)
        else:
            x, ls = ops.linear(y, p["w_in"], bias=p["b_in"], f32=True, bf16=False), None
        xb = self.transformer_blocks[0]._fwd_temporal(x, B, T, S, ls)
        if delta:
            return ops.linear(xb, p["w_out"], bias=p["b_out"], f32=ops.fp32_mode(), bf16=not ops.fp32_mode()), None
        return ops.linear(xb, p["w_out"], bias=p["b_out"], residual=residual, f32=True, bf16=False, stats_rows=S)

    def _fwd(self, h):
        out, cs = self._core(self._pre_norm(h), h.B, h.T, h.H * h.W, h.x)
        return h.like(out, cs=cs)

    def forward(self, x, context=None):
        """x: [b, c, t, h, w] fp32 (attention.py:365)."""
        b, c, t, hh, ww = x.shape
        rows = x.float().permute(0, 2, 3, 4, 1).reshape(b * t * hh * ww, c).contiguous()
        out = self._fwd(Act(rows, b, t, hh, ww))
        return out.x.reshape(b, t, hh, ww, c).permute(0, 4, 1, 2, 3).contiguous()
