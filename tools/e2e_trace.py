"""Per-call wall time of the numpy-in Chamfer call on page-locked buffers, from a cold start:
python tools/e2e_trace.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pcu_b200 as pcu
I = pcu._pcu_internal
rng = np.random.default_rng(0)
n = 1000000
x = rng.random((n, 3), dtype=np.float32); y = rng.random((n, 3), dtype=np.float32)
xp = torch.from_numpy(x).pin_memory(); yp = torch.from_numpy(y).pin_memory()
xd, yd = xp.cuda(), yp.cuda()
for _ in range(33):
    c = float(pcu.chamfer_distance(xd, yd))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def trace(label, calls, prof=False):
    ts, stages = [], []
    if prof:
        I._set_profiling(0, None, True)
    for i in range(calls):
        t0 = time.perf_counter()
        float(pcu.chamfer_distance(xp.numpy(), yp.numpy()))
        ts.append((time.perf_counter() - t0) * 1e3)
        if prof:
            p = I._last_profile(0, None)
            stages.append(round(p.get("bbox+grid", 0) * 1e3))
    if prof:
        I._set_profiling(0, None, False)
    print(label, "ms per call:", " ".join("%.2f" % t for t in ts), flush=True)
    if prof:
        print("   H2D wait (us):", stages, flush=True)
trace("cold      ", 40)
trace("profiled  ", 20, prof=True)
trace("again     ", 20)
for _ in range(20):
    flush.fill_(1); c = float(pcu.chamfer_distance(xd, yd))
trace("after a device-only phase", 20)
time.sleep(0.3)
trace("after 0.3 s idle", 12)
