"""In-kernel timeline of flash_tc (CTA 0): build with DCB_NVCC_EXTRA=-DDCB_TRACE; python tools/trace_flash.py [B H N]
codes: 1 producer issues K/V(j) | 2 mma: K(j) landed | 3 mma: S buffer free, issue S(j) | 4 mma: P(j) ready | 5 mma: V(j)
landed, issue PV(j) | 6 softmax: S(j) ready (start row max) | 7 softmax: P(j) written | 8 softmax: next row max done,
wait O(j) | 9 softmax: O(j) ready"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops, _lib
dev = torch.device("cuda")
B, H, N = (32, 5, 2560) if len(sys.argv) < 4 else tuple(int(v) for v in sys.argv[1:4])
inner = H * 64
qkv = torch.randn(B * N, 3 * inner, device=dev).to(torch.bfloat16)
out = torch.empty(B * N, inner, dtype=torch.bfloat16, device=dev)
for _ in range(3):
    ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], N, 1, 1.0)], B, H, N, 0.125, out=out)
fn = _lib.lib().dcb_debug_flash_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * (4 * 4096))()
assert fn(buf, 4 * 4096) > 0
names = {1: "prod_issue_kv", 2: "mma_k_landed", 3: "mma_issue_S", 4: "mma_p_ready", 5: "mma_issue_PV", 6: "sm_S_ready",
         7: "sm_P_written", 8: "sm_wait_O", 9: "sm_O_ready"}
ev = []
for role in range(4):
    n = int(buf[role * 4096])
    ev += [(buf[role * 4096 + i] >> 20, role, (buf[role * 4096 + i] >> 12) & 0xff, buf[role * 4096 + i] & 0xfff) for i in range(1, n)]
ev.sort()
t0 = ev[0][0]
print(f"{len(ev)} events, span {ev[-1][0] - t0} clk, {N // 128} tiles")
for t, role, c, arg in ev[:int(os.environ.get("TRACE_LINES", "200"))]:
    print(f"{t - t0:8d} r{role} {names.get(c, c):>14s} {arg}")
