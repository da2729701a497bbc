"""Deterministic synthetic weights and inputs (no checkpoints or datasets are reachable offline).

Every tensor is drawn from its own ``torch.Generator`` seeded by crc32(key) ^ seed, so a
tensor's values depend only on its name, shape and the seed -- not on module construction
order -- and are reproducible on any box with the same torch build.  Unlike the reference's
default init (``zero_module`` zeroes 423 of the 1516 UNet tensors, openaimodel3d.py:179,545,
attention.py:288-290) nothing here is zero, so a parity test cannot pass for a no-op.
"""
import zlib

import torch


def synth_tensor(key, shape, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    u = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        t = u / fan_in ** 0.5
    elif len(shape) == 1:
        t = 1.0 + 0.1 * u if key.endswith("weight") else 0.05 * u
    else:
        t = 0.3 * u
    return t.to(dtype)


def synth_state_dict(shapes, seed=0):
    """shapes: mapping key -> shape (e.g. ``{k: v.shape for k, v in m.state_dict().items()}``)."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def synth_input(name, shape, seed=0, dtype=torch.float32):
    """Standard-normal synthetic input tensor, keyed like the weights."""
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(("input/" + name).encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32).to(dtype)
