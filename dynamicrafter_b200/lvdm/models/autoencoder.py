"""Drop-in mirror of lvdm/models/autoencoder.py::AutoencoderKL for the decode path (autoencoder.py:104-107).

Same constructor arguments and ``state_dict`` keys (encoder.*, decoder.*, quant_conv.*, post_quant_conv.*) so
the reference checkpoint loads unchanged; ``decode`` runs post_quant_conv + Decoder on the dcb200 kernels.
``encode`` is outside the hot path (once per video) and is not provided.
"""
import torch
import torch.nn as nn

from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm._packing import PackedModule
from dynamicrafter_b200.lvdm.modules.networks.ae_modules import Encoder, Decoder


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
        return self.post_quant_conv.weight._version + self.post_quant_conv.bias._version

    def _pack(self):
        w = self.post_quant_conv.weight.detach()
        zc, ed = w.shape[0], w.shape[1]
        wp = torch.zeros((zc, 64), dtype=torch.bfloat16, device=w.device)      # K padded to one 64-wide chunk
        wp[:, :ed] = w.reshape(zc, ed).to(torch.bfloat16)
        return {"w": wp, "b": self.post_quant_conv.bias.detach().float().contiguous(), "zc": zc}

    def encode(self, x, **kwargs):
        raise NotImplementedError("VAE encode is outside the dcb200 hot path (SURVEY.md 8(f) N2)")

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
            o, H, W = self.decoder._fwd_rows(pq, n, hh, ww)
            return ops.rows_to_nchw(o, n, self.decoder.out_ch, 1, H, W, five_d=False).to(z.dtype)

    def forward(self, input, sample_posterior=True):
        raise NotImplementedError("training forward is outside the dcb200 hot path")
