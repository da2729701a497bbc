#!/usr/bin/env python
"""Recipe: stage the UNMODIFIED reference's hot-path Python packages under oracle/_ref/ (git-ignored, NOT
gpurun-ignored, so it travels to the GPU box like a built .so).

    python oracle/build_ref.py          # needs /root/reference (build container only)

Nothing is compiled -- the reference is pure Python -- so "building" oracle/_ref is a verbatim copy of
``lvdm/``, ``utils/`` and ``configs/inference_*.yaml`` from where they lie under /root/reference.  The copy is
TEST / BASELINE infrastructure only: ``bench.py --impl reference`` and ``cpu_baseline`` time it on the host cores,
``tests/`` use it to drive the dcb200 mirror classes through the reference's own host code
(``instantiate_from_config`` -> ``LatentVisualDiffusion`` -> ``DDIMSampler``).  Nothing under
``dynamicrafter_b200/`` imports it, and no reference source file ever enters the git history.
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("DCB_REFERENCE", "/root/reference")


def build(verbose=True):
    """Returns the staged path, or None when the reference tree is not available (GPU box: use the prebuilt copy)."""
    if not os.path.isdir(os.path.join(SRC, "lvdm")):
        return DST if os.path.isdir(os.path.join(DST, "lvdm")) else None
    os.makedirs(DST, exist_ok=True)
    for pkg in ("lvdm", "utils"):
        d = os.path.join(DST, pkg)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(SRC, pkg), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DST, "configs"), exist_ok=True)
    for y in glob.glob(os.path.join(SRC, "configs", "inference_*.yaml")):
        shutil.copy2(y, os.path.join(DST, "configs", os.path.basename(y)))
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"oracle/_ref: {n} files staged from {SRC}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
