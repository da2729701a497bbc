"""In-kernel timeline of flash_pp (CTA 0): build with DCB_NVCC_EXTRA=-DDCB_TRACE, run with DCB_FLASH_PP=1.
codes: softmax groups (roles 1, 2): 0 tile start, 1 S ready, 2 S in registers, 3 max / rescale done, 4 P published;
mma warp (role 0): 5 / 6 P of set A / B seen."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops, _lib
dev = torch.device("cuda")
B, H, N = 32, 5, 2560
inner = H * 64
qkv = torch.randn(B * N, 3 * inner, device=dev).to(torch.bfloat16)
for _ in range(3):
    ops.flash_attention(qkv[:, :inner], [(qkv[:, inner:2 * inner], qkv[:, 2 * inner:], N, 1, 1.0)], B, H, N, 0.125)
fn = _lib.lib().dcb_debug_pp_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_ulonglong * (3 * 4096))()
assert fn(buf, 3 * 4096) > 0
ev = []
for role in range(3):
    n = int(buf[role * 4096])
    ev += [(buf[role * 4096 + i] >> 20, role, (buf[role * 4096 + i] >> 12) & 0xff, buf[role * 4096 + i] & 0xfff) for i in range(1, n)]
ev.sort()
t0 = ev[0][0]
names = {0: "start", 1: "S_ready", 2: "S_in_regs", 3: "max_done", 4: "P_done", 5: "mma_P_A", 6: "mma_P_B",
         7: "mma_S_begin", 8: "mma_kv_ok", 9: "mma_PV_issued", 10: "mma_S_issued"}
print(f"{len(ev)} events, span {ev[-1][0] - t0} clk")
for t, role, c, arg in ev[:int(os.environ.get("TRACE_LINES", "260"))]:
    print(f"{t - t0:8d} r{role} {names.get(c, c):>10s} {arg}")
