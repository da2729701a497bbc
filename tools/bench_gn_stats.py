"""Producer-fused GroupNorm statistics: cost on the producer GEMMs and time of the consumer, measured with CUDA events
around a CUDA graph of N back-to-back launches (no Python launch overhead in the numbers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm import _packing

dev = torch.device("cuda")
torch.manual_seed(0)
NREP = 20


def graph_time(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(NREP):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / NREP * 1e3


def rnd(*s, dt=torch.bfloat16):
    return torch.randn(*s, device=dev).to(dt)


print("== producers (us per launch): stats off / on")
B, T = 2, 16
for name, HW, C in (("level0", 2560, 320), ("level1", 640, 640), ("level2", 160, 1280), ("level3", 40, 1280)):
    M = B * T * HW
    H = {2560: 40, 640: 20, 160: 10, 40: 5}[HW]
    W = HW // H
    y = rnd(M, C)
    w3 = rnd(C, 3 * C) * (3 * C) ** -0.5
    w9 = rnd(C, 9 * C) * (9 * C) ** -0.5
    w1 = rnd(C, C) * C ** -0.5
    bias = torch.randn(C, device=dev)
    res = torch.randn(M, C, device=dev)
    cases = {
        "tconv3 bf16-out": lambda st: ops.tconv3(y, B, T, HW, w3, bias=bias, f32=False, bf16=True, stats=st),
        "tconv3 f32+res": lambda st: ops.tconv3(y, B, T, HW, w3, bias=bias, residual=res, stats=st),
        "conv3x3 bf16-out": lambda st: ops.conv3x3(y, B * T, H, W, w9, bias=bias, f32=False, bf16=True, stats=st),
        "conv3x3 f32+res": lambda st: ops.conv3x3(y, B * T, H, W, w9, bias=bias, residual=res, stats=st),
        "linear f32+res": lambda st: ops.linear(y, w1, bias=bias, residual=res, f32=True, bf16=False, stats_rows=HW if st else 0),
    }
    for cname, fn in cases.items():
        t0 = graph_time(lambda: fn(False))
        with ops.stats_scope(dev):
            t1 = graph_time(lambda: fn(True))
        print(f"{name} {cname:18s} M={M:6d} C={C:4d}: {t0:7.1f} -> {t1:7.1f} us  ({t1 - t0:+.1f})")

print("== consumers (us per GroupNorm): cooperative fused kernel / stream apply from producer statistics")
for name, HW, C in (("level0", 2560, 320), ("level1", 640, 640), ("level2", 160, 1280), ("level3", 40, 1280)):
    M = B * T * HW
    for dt in (torch.float32, torch.bfloat16):
        x = rnd(M, C, dt=dt)
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        for G, S, tag in ((B, T * HW, "5-D"), (B * T, HW, "2-D")):
            t_coop = graph_time(lambda: ops.group_norm(x, G, S, gamma, beta, 1e-5, silu=True))
            acc = torch.ones((B * T, 32, 2), dtype=torch.int64, device=dev) * (1 << 30)
            cs = ops.ColSums(acc, HW, C, 0)
            t_str = graph_time(lambda: ops.group_norm(x, G, S, gamma, beta, 1e-5, silu=True, colsums=cs))
            nbytes = x.numel() * x.element_size() + x.numel() * 2
            print(f"{name} {tag} {str(dt)[6:]:8s} [{M}x{C}]: coop {t_coop:6.1f} us  stream {t_str:6.1f} us "
                  f"({nbytes / t_str / 1e3:5.0f} GB/s)")
