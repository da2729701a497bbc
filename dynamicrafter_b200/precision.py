"""Arithmetic mode of the denoising path.

"bf16" (default): bf16 tensor-core operands, fp32 accumulation / norms / softmax / residual stream (parity bar:
rel-L2 <= 2e-2 per forward).  "fp32": verification mode for the <= 1e-3 bar -- every GEMM operand is split into
bf16 hi/lo terms (three tcgen05 products per real product, accumulated in fp32), everything else stays fp32
(csrc/fp32_mode.cu).  Select with DCB_PRECISION=fp32 or set_precision("fp32") BEFORE the first forward of a module
(packed weights are cached per mode)."""
import os

_MODE = [os.environ.get("DCB_PRECISION", "bf16")]
assert _MODE[0] in ("bf16", "fp32"), "DCB_PRECISION must be bf16 or fp32"


def set_precision(mode):
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _MODE[0] = mode


def get_precision():
    return _MODE[0]


def fp32_mode():
    return _MODE[0] == "fp32"


class use_precision:
    """Context manager: run a region in the given mode (the VAE decode always runs in bf16: its parity bar is a
    PSNR bar and its attention / 1x1 paths are not built for the fp32 mode)."""

    def __init__(self, mode):
        if mode not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.mode = mode

    def __enter__(self):
        self.prev, _MODE[0] = _MODE[0], self.mode

    def __exit__(self, *exc):
        _MODE[0] = self.prev
