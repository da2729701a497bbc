"""LayerNorm folding: cost on the producer GEMM (bf16 copy + per-row statistics) and on the consumer GEMMs (per-row
normalisation in the epilogue) against the plain GEMMs + the layer_norm kernel, CUDA-graph timed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from dynamicrafter_b200 import ops
from dynamicrafter_b200.lvdm.modules.attention import fold_layer_norm, _fold_ln_weights
from dynamicrafter_b200.lvdm._packing import pack_geglu
from bench_shapes import graph_time, rnd

dev = torch.device("cuda")
for M, C in ((81920, 320), (20480, 640), (5120, 1280)):
    a, w0 = rnd(M, C), rnd(C, C) * C ** -0.5
    b0, res = torch.randn(C, device=dev), rnd(M, C, dt=torch.float32)
    norm = nn.LayerNorm(C).to(dev)
    t_plain = graph_time(lambda: ops.linear(a, w0, bias=b0, residual=res, f32=True, bf16=False))
    with ops.stats_scope(dev):
        t_prod = graph_time(lambda: ops.linear(a, w0, bias=b0, residual=res, f32=True, bf16=False, ln_stats=True))
    y, src = ops.linear(a, w0, bias=b0, residual=res, f32=True, bf16=False, ln_stats=True)
    t_ln = graph_time(lambda: ops.layer_norm(y, norm.weight, norm.bias, norm.eps))
    xn = ops.layer_norm(y, norm.weight, norm.bias, norm.eps)
    print(f"M={M} C={C}: producer {t_plain:.1f} -> {t_prod:.1f} us (+{t_prod - t_plain:.1f}); layer_norm kernel {t_ln:.1f} us")
    for N in (3 * C, C):
        w1 = torch.randn(N, C, device=dev) * C ** -0.5
        fold = fold_layer_norm(norm, w1, None)
        wq = w1.to(torch.bfloat16).contiguous()
        t0 = graph_time(lambda: ops.linear(xn, wq))
        t1 = graph_time(lambda: ops.linear(src, None, ln=fold))
        print(f"    consumer N={N}: {t0:.1f} -> {t1:.1f} us (+{t1 - t0:.1f})")
    w1 = torch.randn(8 * C, C, device=dev) * C ** -0.5
    b1 = torch.randn(8 * C, device=dev) * 0.1
    wf, bf = _fold_ln_weights(norm, w1, b1)
    wq, bq, bn = pack_geglu(wf, bf)
    fold = ops.LnFold(wq, bq, wq.float().sum(1).contiguous(), norm.eps)
    wp, bp, _ = pack_geglu(w1, b1)
    t0 = graph_time(lambda: ops.linear(xn, wp, bias=bp, geglu=True, block_n=bn))
    t1 = graph_time(lambda: ops.linear(src, None, geglu=True, block_n=bn, ln=fold))
    print(f"    consumer GEGLU N={8 * C}: {t0:.1f} -> {t1:.1f} us (+{t1 - t0:.1f})")
