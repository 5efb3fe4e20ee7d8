"""Three device-resident fused Chamfer calls (for ncu captures): python tools/run_chamfer.py [n] [m] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcu_b200 as pcu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = torch.Generator(device="cuda").manual_seed(0)
if batch:
    x = torch.rand((batch, n, 3), generator=g, device="cuda"); y = torch.rand((batch, m, 3), generator=g, device="cuda")
    for _ in range(3):
        c = pcu.batched_chamfer_distance(x, y)
    c = c.sum()
else:
    x = torch.rand((n, 3), generator=g, device="cuda"); y = torch.rand((m, 3), generator=g, device="cuda")
    for _ in range(3):
        c = pcu.chamfer_distance(x, y)
torch.cuda.synchronize()
print("ok", float(c))
