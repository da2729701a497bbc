"""GroupNorm micro-benchmark under a CUDA graph: per shape, N x [producer-like write of x, group_norm(x)] minus the
writes alone (so x is in L2 exactly as after the GEMM that produced it).  python tools/bench_gn.py [bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops

dev = torch.device("cuda")
SHAPES = [(2, 40960, 320), (2, 10240, 640), (2, 2560, 1280), (2, 640, 1280), (32, 2560, 320), (32, 640, 640),
          (32, 160, 1280), (32, 40, 1280)]
N = 10
dt = torch.bfloat16 if "bf16" in sys.argv else torch.float32


def graph_ms(fn, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for G, S, C in (SHAPES if __name__ == "__main__" else []):
    src = torch.randn(G * S, C, device=dev).to(dt)
    xs = [torch.empty_like(src) for _ in range(N)]
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)

    def writes():
        for x in xs:
            x.copy_(src)

    def both():
        for x in xs:
            x.copy_(src)
            ops.group_norm(x, G, S, gam, bet, 1e-5)

    tw, tb = graph_ms(writes), graph_ms(both)
    mb = G * S * C * src.element_size() / 1e6
    print(f"G={G:3d} S={S:6d} C={C:5d} {str(dt)[6:]:>8s} {mb:7.1f} MB  group_norm {1e3 * (tb - tw) / N:7.2f} us  "
          f"(write alone {1e3 * tw / N:6.2f} us)")
