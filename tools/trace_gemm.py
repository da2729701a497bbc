"""In-kernel timeline of one tap_gemm launch (CTA 0): build with DCB_NVCC_EXTRA=-DDCB_TRACE, then
    python tools/trace_gemm.py M N K [f32] [res] [geglu]
prints, per warp role, the clock of every event relative to the first one.
codes: 1 producer stage issue(it) | 2 mma tile start(tc) | 3 mma stage consumed(it) | 4+g epilogue group g accumulator
ready(tc) | 8+g residual chunk landed(cnt) | 12+g chunk stored(cnt)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops, _lib
dev = torch.device("cuda")
M, N, K = (int(v) for v in sys.argv[1:4])
f32, res, geglu = "f32" in sys.argv, "res" in sys.argv, "geglu" in sys.argv
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
out = torch.empty(M, N // 2 if geglu else N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
r = torch.randn(M, N, device=dev) if res else None
L = _lib.lib()
fn = L.dcb_debug_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * (6 * 4096))()
kw = dict(geglu=True, block_n=256 if N % 256 == 0 else 128) if geglu else {}
for _ in range(4):
    ops.linear(a, w, residual=r, out=out, **kw)
assert fn(buf, 6 * 4096) > 0
ev = []
for role in range(6):
    n = int(buf[role * 4096])
    ev += [(buf[role * 4096 + i] >> 20, (buf[role * 4096 + i] >> 12) & 0xff, buf[role * 4096 + i] & 0xfff) for i in range(1, n)]
ev.sort()
t0 = ev[0][0]
names = {1: "prod", 2: "mma_tile", 3: "mma_stage"}
for g in range(4):
    names[4 + g] = f"epi{g}_acc"; names[8 + g] = f"epi{g}_res"; names[12 + g] = f"epi{g}_store"
print(f"{len(ev)} events, span {ev[-1][0] - t0} clk")
limit = int(os.environ.get("TRACE_LINES", "400"))
for t, c, arg in ev[:limit]:
    print(f"{t - t0:8d} {names.get(c, c):>10s} {arg}")
