import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dynamicrafter_b200 import ops, _lib
dev = torch.device("cuda")
G, S, C = 2, 40960, 320
x = torch.randn(G * S, C, device=dev)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(3):
    y = ops.group_norm(x, G, S, g, b, 1e-5)
torch.cuda.synchronize()
print("ok")
