"""Weight packing shared by the mirror modules: fp32 nn.Parameters -> bf16 K-major GEMM operands.

Packed tensors are plain attributes (never in ``state_dict``), rebuilt lazily whenever a parameter's
``_version`` changes (``load_state_dict`` / in-place re-randomisation), so the modules keep the
reference's checkpoint format (inference.py:34-59 loads with strict=True).
"""
import torch
import torch.nn as nn

BF16 = torch.bfloat16

import os as _os

# Block-internal activations that only feed a GroupNorm may be written by their producer GEMM as bf16 (the
# GroupNorm output is a bf16 operand anyway; it halves the HBM traffic of the three kernels touching them).
# DCB_BF16_INTERNAL: comma list of {tconv (TemporalConvBlock stages 1-3), res (ResBlock h after conv1),
# vae (VAE ResnetBlock h after conv1)}; "0" = none.  Default: see DESIGN.md (precision budget).
_LOWP = set(_os.environ.get("DCB_BF16_INTERNAL", "tconv,res").split(","))
BF16_TCONV = "tconv" in _LOWP or "all" in _LOWP
BF16_RES = "res" in _LOWP or "all" in _LOWP
BF16_VAE = "vae" in _LOWP or "all" in _LOWP
# Producer-fused GroupNorm statistics: every GEMM / conv whose output feeds a GroupNorm folds per-column (sum, sumsq)
# partials in its epilogue (read back from the staged output tile in shared memory); the consumer is a tiny finalize
# kernel + ONE streaming apply pass (no statistics pass over the tensor, no cooperative launch / grid barrier).
# DCB_FUSED_GN_STATS=0 restores the self-contained cooperative GroupNorm everywhere (A/B testing); it is still used
# for skip-concat inputs (two sources), the 1x1-skip ResBlocks (raw copy wanted) and tiles the epilogue cannot segment.
FUSED_GN_STATS = _os.environ.get("DCB_FUSED_GN_STATS", "1") == "1"
# frame-sharded mode: also capture the NCCL calls of a forward in the CUDA graph.  Works (94.4 vs 97.7 ms/step at the
# 1024 config on 2 GPUs) but process-group teardown then hangs with torch 2.11 / NCCL 2.28 -> off by default.
GRAPH_COLLECTIVES = _os.environ.get("DCB_GRAPH_COLLECTIVES", "0") == "1"
# frame-sharded mode: DCB_SHARD_V1=1 transposes the fp32 residual stream both ways around every temporal block (round 1);
# default: bf16 operand one way, bf16 delta back, residual add fused into the unpack (UNetModel._temporal).
SHARD_V1 = _os.environ.get("DCB_SHARD_V1", "0") == "1"


# Cross-attention keys / values depend on the conditioning only (attention.py:86-94): UNetModel projects them once per
# context TENSOR (same object, same version) and reuses them over the DDIM steps.  DCB_CACHE_CROSS_KV=0: every forward.
CACHE_CROSS_KV = _os.environ.get("DCB_CACHE_CROSS_KV", "1") == "1"


# nearest-x2 upsample folded into the conv behind it (four 2x2-tap phase GEMMs, 2.25x fewer FLOPs, no 4x tensor).
# DCB_FOLD_UPSAMPLE=0: materialise the upsampled tensor and run the 3x3 conv on it.
FOLD_UPSAMPLE = _os.environ.get("DCB_FOLD_UPSAMPLE", "1") == "1"


from dynamicrafter_b200.precision import fp32_mode, get_precision


def _pad64(c):
    return (c + 63) // 64 * 64


def _split_w(w):
    """fp32 parity mode: fp32 [N, taps, C] -> bf16 [N, taps * 3C], per tap (hi | hi | lo); pairs with the
    (hi | lo | hi) split of the activations so that one GEMM sums a_hi w_hi + a_lo w_hi + a_hi w_lo."""
    w = w.float()
    hi = w.to(BF16)
    lo = (w - hi.float()).to(BF16)
    return torch.cat([hi, hi, lo], dim=-1).reshape(w.shape[0], -1).contiguous()


def pack_linear(w):
    """[N, K] (or Conv1d/Conv2d k=1 [N, K, 1(,1)]) -> bf16 [N, K]."""
    if fp32_mode():
        return _split_w(w.detach().reshape(w.shape[0], 1, w.shape[1]))
    return w.detach().reshape(w.shape[0], w.shape[1]).to(BF16).contiguous()


def pack_conv3x3(w):
    """[N, C, 3, 3] -> bf16 [N, 9*Cpad] ordered (ky, kx, c); input channels zero-padded to 64."""
    N, C = w.shape[0], w.shape[1]
    Cp = _pad64(C)
    if fp32_mode():
        full = torch.zeros((N, 3, 3, Cp), dtype=torch.float32, device=w.device)
        full[..., :C] = w.detach().permute(0, 2, 3, 1).float()
        return _split_w(full.reshape(N, 9, Cp))
    out = torch.zeros((N, 3, 3, Cp), dtype=BF16, device=w.device)
    out[..., :C] = w.detach().permute(0, 2, 3, 1).to(BF16)
    return out.reshape(N, 9 * Cp).contiguous()


def pack_conv3x3_up2(w):
    """[N, C, 3, 3] -> 4 phase matrices (py, px) for ops.conv3x3_up2, each bf16 [N, 4*Cpad] ordered (a_y, b_x, c): the 3x3
    kernel applied to a nearest-x2 upsampled image collapses, per output phase, to a 2x2 kernel on the original image
    whose taps are sums of the 3x3 taps that land on the same source pixel."""
    N, C = w.shape[0], w.shape[1]
    Cp = _pad64(C)
    wf = w.detach().float()
    sets = (((0,), (1, 2)), ((0, 1), (2,)))          # phase -> (kernel taps folded into offset a = 0, a = 1)
    out = []
    for py in range(2):
        for px in range(2):
            full = torch.zeros((N, 2, 2, Cp), dtype=torch.float32, device=w.device)
            for a in range(2):
                for b in range(2):
                    acc = 0
                    for ky in sets[py][a]:
                        for kx in sets[px][b]:
                            acc = acc + wf[:, :, ky, kx]
                    full[:, a, b, :C] = acc
            if fp32_mode():
                out.append(_split_w(full.reshape(N, 4, Cp)))
            else:
                out.append(full.reshape(N, 4 * Cp).to(BF16).contiguous())
    return out


def pack_tconv3(w):
    """[N, C, 3, 1, 1] -> bf16 [N, 3*C] ordered (kt, c)."""
    N, C = w.shape[0], w.shape[1]
    if fp32_mode():
        return _split_w(w.detach().reshape(N, C, 3).permute(0, 2, 1))
    return w.detach().reshape(N, C, 3).permute(0, 2, 1).to(BF16).reshape(N, 3 * C).contiguous()


def geglu_block_n(inner):
    """The GEGLU epilogue pairs value/gate columns inside one N tile: (BN/2 values | BN/2 gates)."""
    if inner % 128 == 0:
        return 256
    if inner % 64 == 0:
        return 128
    raise ValueError(f"GEGLU inner dim {inner} must be a multiple of 64")


def pack_geglu(w, b):
    """[2*inner, K] -> rows interleaved per output tile: (BN/2 value rows, BN/2 gate rows) * tiles."""
    inner = w.shape[0] // 2
    if fp32_mode():             # plain GEMM + separate exact gate: rows stay in (value..., gate...) order
        return _split_w(w.detach().reshape(w.shape[0], 1, w.shape[1])), b.detach().float().contiguous(), 0
    bn = geglu_block_n(inner)
    half = bn // 2
    idx = torch.arange(inner, device=w.device).reshape(inner // half, half)
    idx = torch.cat([idx, idx + inner], dim=1).reshape(-1)
    return w.detach()[idx].to(BF16).contiguous(), b.detach()[idx].float().contiguous(), bn


class PackedModule(nn.Module):
    """nn.Module whose kernels consume a lazily built dict of packed device tensors."""

    def _pack(self):           # -> dict
        raise NotImplementedError

    def _own_params(self):
        """Own parameters + those of the plain torch children (whose weights this module packs), as a flat list built
        once: walking the module tree on every packed() call cost an eager forward ~20 % of its host time."""
        lst = self.__dict__.get("_plist")
        if lst is None:
            lst = list(self.parameters(recurse=False))
            for m in self.children():
                if not isinstance(m, PackedModule):
                    lst += list(m.parameters())
            first = lst[0] if lst else next(self.parameters(), None)
            lst = self.__dict__["_plist"] = (lst, first)
        return lst

    def _drop_param_caches(self):
        for k in ("_plist", "_allp", "_param_list"):
            self.__dict__.pop(k, None)

    def _apply(self, fn, *a, **kw):
        self._drop_param_caches()
        return super()._apply(fn, *a, **kw)

    def _load_from_state_dict(self, *a, **kw):
        # called for every module of the tree by any load_state_dict (also a parent's, also assign=True, which replaces
        # the Parameter objects): the cached lists are rebuilt on the next packed()
        self._drop_param_caches()
        return super()._load_from_state_dict(*a, **kw)

    def _own_version(self):
        v = 0
        for p in self._own_params()[0]:
            v += p._version
        return v

    def packed(self):
        p0 = self._own_params()[1]
        if p0 is None or not p0.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: parameters must live on a CUDA device "
                               "(dcb200 has no CPU path)")
        key = (p0.device, self._own_version(), get_precision())
        cache = self.__dict__.get("_packed_cache")
        if cache is None or cache[0] != key:
            with torch.no_grad():
                cache = (key, self._pack())
            self.__dict__["_packed_cache"] = cache
        return cache[1]


class Act:
    """Channels-last activation: x is fp32 [B*T*H*W, C], rows ordered (b, t, y, x).  cs: the per-column statistics
    partials its producer GEMM emitted (ops.ColSums) or None -- lets the consumer GroupNorm skip its statistics pass."""
    __slots__ = ("x", "B", "T", "H", "W", "cs")

    def __init__(self, x, B, T, H, W, cs=None):
        self.x, self.B, self.T, self.H, self.W, self.cs = x, B, T, H, W, cs

    @property
    def C(self):
        return self.x.shape[1]

    def like(self, x, H=None, W=None, cs=None):
        return Act(x, self.B, self.T, self.H if H is None else H, self.W if W is None else W, cs)
