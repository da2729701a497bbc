"""Drop-in mirror of lvdm/models/autoencoder.py::AutoencoderKL (autoencoder.py:97-107).

Same constructor arguments and ``state_dict`` keys (encoder.*, decoder.*, quant_conv.*, post_quant_conv.*) so
the reference checkpoint loads unchanged; ``decode`` runs post_quant_conv + Decoder and ``encode`` runs Encoder +
quant_conv on the dcb200 kernels (encode: the conditioning prologue of image-to-video) and returns
the same DiagonalGaussianDistribution interface as lvdm/distributions.py:24-69.
"""
import torch
import torch.nn as nn

from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm._packing import PackedModule
from dynamicrafter_b200.lvdm.modules.networks.ae_modules import Encoder, Decoder


class DiagonalGaussianDistribution:
    """lvdm/distributions.py:24-69 (host-side elementwise glue on the tiny moments tensor)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean).to(device=self.parameters.device)

    def sample(self, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape)                 # CPU generator, like the reference (:37)
        return self.mean + self.std * noise.to(device=self.parameters.device)

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0
                               - self.logvar + other.logvar, dim=[1, 2, 3])

    def nll(self, sample, dims=(1, 2, 3)):
        if self.deterministic:
            return torch.Tensor([0.])
        logtwopi = 1.8378770664093453
        return 0.5 * torch.sum(logtwopi + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=list(dims))

    def mode(self):
        return self.mean


class AutoencoderKL(PackedModule):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, test=False, logdir=None, input_dim=4, test_args=None):
        super().__init__()
        self.image_key = image_key
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        assert ddconfig["double_z"]
        self.quant_conv = torch.nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = torch.nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim, self.input_dim = embed_dim, input_dim
        if monitor is not None:
            self.monitor = monitor
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu")
            self.load_state_dict(sd.get("state_dict", sd), strict=False)

    def _own_version(self):
        return sum(q._version for q in (self.post_quant_conv.weight, self.post_quant_conv.bias, self.quant_conv.weight,
                                        self.quant_conv.bias))

    def _pack(self):
        w = self.post_quant_conv.weight.detach()
        zc, ed = w.shape[0], w.shape[1]
        wp = torch.zeros((zc, 64), dtype=torch.bfloat16, device=w.device)      # K padded to one 64-wide chunk
        wp[:, :ed] = w.reshape(zc, ed).to(torch.bfloat16)
        wq = self.quant_conv.weight.detach()
        return {"w": wp, "b": self.post_quant_conv.bias.detach().float().contiguous(), "zc": zc,
                "wq": wq.reshape(wq.shape[0], wq.shape[1]).to(torch.bfloat16).contiguous(),
                "bq": self.quant_conv.bias.detach().float().contiguous()}

    @torch.no_grad()
    def encode(self, x, **kwargs):
        """x: [n, 3, H, W] fp32 in [-1, 1] -> DiagonalGaussianDistribution over [n, embed_dim, H/8, W/8]
        (autoencoder.py:97-102)."""
        if not x.is_cuda:
            raise RuntimeError("dcb200 AutoencoderKL.encode runs on CUDA tensors only (no CPU path)")
        from dynamicrafter_b200.precision import use_precision
        with use_precision("bf16"):
            p = self.packed()
            n, c, hh, ww = x.shape
            with ops.stats_scope(x.device):
                h, H, W = self.encoder._fwd_rows(ops.nchw_to_rows(x, 64), n, hh, ww)      # fp32 [n*H*W, 2z]
            hb = ops.cast_bf16(h.contiguous())
            m = ops.linear(hb, p["wq"], bias=p["bq"], f32=True, bf16=False)             # 1x1 quant_conv
            moments = ops.rows_to_nchw(m, n, m.shape[1], 1, H, W, five_d=False).to(x.dtype)
        return DiagonalGaussianDistribution(moments)

    @torch.no_grad()
    def decode(self, z, **kwargs):
        """z: [n, embed_dim, h, w] fp32 (already divided by scale_factor) -> [n, 3, 8h, 8w]."""
        if not z.is_cuda:
            raise RuntimeError("dcb200 AutoencoderKL.decode runs on CUDA tensors only (no CPU path)")
        from dynamicrafter_b200.precision import use_precision
        with use_precision("bf16"):                             # the decoder has no fp32 parity mode (PSNR bar)
            p = self.packed()
            n, c, hh, ww = z.shape
            zr = ops.nchw_to_rows(z, 64)
            pq = torch.zeros((n * hh * ww, 64), dtype=torch.bfloat16, device=z.device)
            ops.linear(zr, p["w"], bias=p["b"], out=pq)      # 1x1 post_quant_conv, written into the padded operand
            with ops.stats_scope(z.device):
                o, H, W = self.decoder._fwd_rows(pq, n, hh, ww)
            return ops.rows_to_nchw(o, n, self.decoder.out_ch, 1, H, W, five_d=False).to(z.dtype)

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("training forward is outside the dcb200 hot path")
