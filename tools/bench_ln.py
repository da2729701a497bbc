"""LayerNorm micro-benchmark under a CUDA graph (N x [write x, layer_norm(x)] minus the writes alone)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops
from bench_gn import graph_ms
dev = torch.device("cuda")
N = 10
for rows, C in [(81920, 320), (20480, 640), (5120, 1280), (1280, 1280), (81920, 512)]:
    src = torch.randn(rows, C, device=dev)
    xs = [torch.empty_like(src) for _ in range(N)]
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)

    def writes():
        for x in xs:
            x.copy_(src)

    def both():
        for x in xs:
            x.copy_(src)
            ops.layer_norm(x, g, b, 1e-5)

    tw, tb = graph_ms(writes), graph_ms(both)
    print(f"rows={rows:6d} C={C:5d} {rows * C * 4 / 1e6:7.1f} MB  layer_norm {1e3 * (tb - tw) / N:7.2f} us  (write alone {1e3 * tw / N:6.2f} us)")
