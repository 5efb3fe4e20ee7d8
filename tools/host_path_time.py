"""Where the time of one numpy-in call goes: python tools/host_path_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pcu_b200 as pcu
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
x = rng.random((n, 3), dtype=np.float32); y = rng.random((n, 3), dtype=np.float32)
xp = torch.from_numpy(x).pin_memory(); yp = torch.from_numpy(y).pin_memory()
xd, yd = xp.cuda(), yp.cuda()
def t(f, reps=30):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("device chamfer (enqueue+sync)  %.3f ms" % t(lambda: float(pcu.chamfer_distance(xd, yd))))
print("host pinned chamfer            %.3f ms" % t(lambda: float(pcu.chamfer_distance(xp.numpy(), yp.numpy()))))
print("host pageable chamfer          %.3f ms" % t(lambda: float(pcu.chamfer_distance(x, y))))
print("host pinned knn k=1            %.3f ms" % t(lambda: pcu.k_nearest_neighbors(xp.numpy(), yp.numpy(), 1)))
print("H2D 2x12MB pinned via torch    %.3f ms" % t(lambda: (xd.copy_(xp, non_blocking=True), yd.copy_(yp, non_blocking=True))))
print("host one-sided hausdorff       %.3f ms" % t(lambda: pcu.one_sided_hausdorff_distance(xp.numpy(), yp.numpy())))
# device-side stage times INSIDE the host call (events on the host path's own stream)
I = pcu._pcu_internal
I._set_profiling(0, None, True)
for name, f in (("chamfer pinned", lambda: float(pcu.chamfer_distance(xp.numpy(), yp.numpy()))),
                ("chamfer pageable", lambda: float(pcu.chamfer_distance(x, y)))):
    acc = {}
    host = 0.0
    for _ in range(10):
        t0 = time.perf_counter(); f(); host += (time.perf_counter() - t0) * 100
        for k, v in I._last_profile(0, None).items():
            acc[k] = acc.get(k, 0.0) + v / 10
    print(name, "host %.3f ms; device stages (us):" % host, {k: round(v * 1e3, 1) for k, v in acc.items()})
I._set_profiling(0, None, False)
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout)
