#!/usr/bin/env python
"""How far is the reference from ITSELF under autocast?  (test infrastructure, build container only)

The reference's CLI samples under `torch.cuda.amp.autocast()` (scripts/evaluation/inference.py:359), i.e. fp16 GEMM /
conv operands and outputs with fp32 norms and softmax.  This probe runs the UNMODIFIED reference UNet (tiny
configurations of tests/golden/unet_*.pt, same synthetic weights and inputs) on the CPU under torch autocast in fp16 and
in bf16 and prints the rel-L2 distance of each to the fp32 run -- the yardstick the bf16 mode of dcb200 is read against
(SURVEY.md 8(f) N4 asks for an fp16-autocast comparison).

    python oracle/autocast_probe.py > profiles/r02_autocast_probe.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    mg.install_shims()
    cases = [("unet_tiny512", mg.TINY_UNET, 1, 8, 16, 3, 601, True),
             ("unet_tiny512_b2", mg.TINY_UNET, 2, 8, 8, 4, 981, False),
             ("unet_tiny256", mg.TINY_UNET_256, 1, 8, 8, 5, 301, True)]
    print("unmodified reference UNetModel.forward on CPU, rel-L2 of the autocast run vs the fp32 run (same weights, inputs)")
    for name, cfg, b, h, w, seed, t_val, with_fs in cases:
        m = mg.build_ref_unet(cfg, seed)
        x, ts, ctx, fs = mg.unet_inputs(cfg, b, h, w, seed, t_val, with_fs)
        gold = torch.load(os.path.join(mg.OUT, name + ".pt"))["y"]
        with torch.no_grad():
            y32 = m(x, ts, context=ctx, fs=fs)
            row = [f"{name:16s} fp32 vs committed golden {rel_l2(y32, gold):.1e}"]
            for dt in (torch.float16, torch.bfloat16):
                try:
                    with torch.autocast("cpu", dtype=dt):
                        y = m(x, ts, context=ctx, fs=fs)
                    row.append(f"{str(dt).split('.')[-1]} autocast {rel_l2(y.float(), y32):.3e}")
                except Exception as exc:             # an op without a CPU kernel in that dtype
                    row.append(f"{str(dt).split('.')[-1]} autocast unavailable ({type(exc).__name__})")
        print("   ".join(row), flush=True)


if __name__ == "__main__":
    main()
