"""One call of an 8(f) row on device-resident inputs, for ncu captures: python tools/run_rows.py ball|dedup|voxel|sinkhorn"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcu_b200 as pcu
which = sys.argv[1]
n = 1000000
g = torch.Generator(device="cuda").manual_seed(0)
if which == "ball":
    sph = torch.nn.functional.normalize(torch.randn((n, 3), generator=g, device="cuda"), dim=1)
    for _ in range(2):
        out = pcu.estimate_point_cloud_normals_ball(sph, 12 * 4.0 / n)
elif which == "dedup":
    x = torch.rand((n, 3), generator=g, device="cuda"); x = torch.cat([x, x[: n // 4]])
    for _ in range(2):
        out = pcu.deduplicate_point_cloud(x, 1e-11)
elif which == "voxel":
    x = torch.rand((n, 3), generator=g, device="cuda")
    for _ in range(2):
        out = pcu.downsample_point_cloud_on_voxel_grid(0.01, x)
elif which == "sinkhorn":
    a = torch.rand((4, 2048, 3), generator=g, device="cuda"); b = torch.rand((4, 2048, 3), generator=g, device="cuda")
    M = pcu.pairwise_distances(a, b)
    w = torch.full((4, 2048), 1.0 / 2048, device="cuda")
    out = pcu.sinkhorn(w, w, M, eps=1e-2, max_iters=5, stop_thresh=0.0)
torch.cuda.synchronize()
print("ok", which)
