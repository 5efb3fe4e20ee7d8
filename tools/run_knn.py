"""One device-resident k-NN call (for ncu captures): python tools/run_knn.py [k] [n] [m]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcu_b200 as pcu
k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1000000
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((n, 3), generator=g, device="cuda"); y = torch.rand((m, 3), generator=g, device="cuda")
for _ in range(3):
    d, i = pcu.k_nearest_neighbors(x, y, k)
torch.cuda.synchronize()
print("ok", float(d.sum()))
