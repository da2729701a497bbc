"""Drop-in mirror of lvdm/modules/encoders/resampler.py::Resampler (the `image_proj_model` of image-to-video: CLIP image
tokens -> 16 context tokens per frame).  Same constructor arguments and state_dict keys; the
projections run on dcb_tap_gemm, the perceiver attention on the flash-attention kernels (head dim 64), LayerNorm / GELU
on the norm / elementwise kernels.  bf16 operands, fp32 residual stream, like the UNet."""
import torch
import torch.nn as nn

from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm._packing import PackedModule, pack_linear


def _f32(t):
    return t.detach().float().contiguous()


def FeedForward(dim, mult=4):
    inner_dim = int(dim * mult)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner_dim, bias=False), nn.GELU(),
                         nn.Linear(inner_dim, dim, bias=False))


class PerceiverAttention(nn.Module):
    def __init__(self, *, dim, dim_head=64, heads=8):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("dcb200 attention kernels are specialised for dim_head == 64")
        self.scale, self.dim_head, self.heads = dim_head ** -0.5, dim_head, heads
        inner_dim = dim_head * heads
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, inner_dim * 2, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)


class Resampler(PackedModule):
    def __init__(self, dim=1024, depth=8, dim_head=64, heads=16, num_queries=8, embedding_dim=768, output_dim=1024,
                 ff_mult=4, video_length=None):
        super().__init__()
        self.num_queries, self.video_length = num_queries, video_length
        if video_length is not None:
            num_queries = num_queries * video_length
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.layers = nn.ModuleList([nn.ModuleList([PerceiverAttention(dim=dim, dim_head=dim_head, heads=heads),
                                                    FeedForward(dim=dim, mult=ff_mult)]) for _ in range(depth)])

    def _own_version(self):
        return sum(q._version for q in self.parameters())

    def _pack(self):
        p = {"w_in": pack_linear(self.proj_in.weight), "b_in": _f32(self.proj_in.bias),
             "w_out": pack_linear(self.proj_out.weight), "b_out": _f32(self.proj_out.bias), "layers": []}
        for attn, ff in self.layers:
            p["layers"].append({"wq": pack_linear(attn.to_q.weight), "wkv": pack_linear(attn.to_kv.weight),
                                "wo": pack_linear(attn.to_out.weight), "w1": pack_linear(ff[1].weight),
                                "w2": pack_linear(ff[3].weight)})
        return p

    @torch.no_grad()
    def forward(self, x):
        """x: [b, n_tokens, embedding_dim] -> [b, num_queries (* video_length), output_dim] (resampler.py:132-145)."""
        if not x.is_cuda:
            raise RuntimeError("dcb200 Resampler runs on CUDA tensors only (no CPU path)")
        from dynamicrafter_b200.precision import use_precision
        with use_precision("bf16"):
            p = self.packed()
            b, n1, E = x.shape
            n2, D = self.latents.shape[1], self.latents.shape[2]
            xp = ops.linear(ops.cast_bf16(x.float().reshape(b * n1, E).contiguous()), p["w_in"], bias=p["b_in"],
                            f32=True, bf16=False)
            lat = self.latents.detach().float().repeat(b, 1, 1).reshape(b * n2, D).contiguous()
            for (attn, ff), pl in zip(self.layers, p["layers"]):
                inner = attn.heads * attn.dim_head
                xn = ops.layer_norm(xp.contiguous(), attn.norm1.weight, attn.norm1.bias, attn.norm1.eps)
                ln = ops.layer_norm(lat.contiguous(), attn.norm2.weight, attn.norm2.bias, attn.norm2.eps)
                q = ops.linear(ln, pl["wq"])
                kv_in = torch.cat([xn.reshape(b, n1, D), ln.reshape(b, n2, D)], 1).reshape(b * (n1 + n2), D)
                kv = ops.linear(kv_in, pl["wkv"])
                a = ops.flash_attention(q, [(kv[:, :inner], kv[:, inner:], n1 + n2, 1, 1.0)], b, attn.heads, n2, attn.scale)
                lat = ops.linear(a, pl["wo"], residual=lat, f32=True, bf16=False)
                hn = ops.layer_norm(lat.contiguous(), ff[0].weight, ff[0].bias, ff[0].eps)
                h1 = ops.linear(hn, pl["w1"], f32=True, bf16=False)
                lat = ops.linear(ops.gelu_bf16(h1.contiguous()), pl["w2"], residual=lat, f32=True, bf16=False)
            out = ops.linear(ops.cast_bf16(lat.contiguous()), p["w_out"], bias=p["b_out"], f32=True, bf16=False)
            y = ops.layer_norm(out.contiguous(), self.norm_out.weight, self.norm_out.bias, self.norm_out.eps, out_f32=True)
        return y.reshape(b, n2, -1).to(x.dtype)
