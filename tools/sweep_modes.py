"""C3 step and stage times over cell occupancies: python tools/sweep_modes.py [occupancies...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcu_b200 as pcu
I = pcu._pcu_internal
g = torch.Generator(device="cuda").manual_seed(0)
n = 1000000
x = torch.rand((n, 3), generator=g, device="cuda"); y = torch.rand((n, 3), generator=g, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
occs = [float(a) for a in sys.argv[1:]] or [0.0]
for occ in occs:
    I._set_defaults(cell_occupancy=occ)
    for _ in range(5):
        c = pcu.chamfer_distance(x, y)
    reps = 30
    tot = 0.0
    for _ in range(reps):       # step time without the profiling events
        flush.fill_(1)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); c = pcu.chamfer_distance(x, y); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    I._set_profiling(0, stream, True)
    acc = {}
    for _ in range(reps):
        flush.fill_(1)
        c = pcu.chamfer_distance(x, y)
        torch.cuda.synchronize()
        for k, v in I._last_profile(0, stream).items():
            acc[k] = acc.get(k, 0.0) + v / reps
    I._set_profiling(0, stream, False)
    print("occupancy %.2f  step %.4f ms  value %.9g  stages (us) %s" % (occ, tot / reps, float(c),
          {k: round(v * 1e3, 1) for k, v in acc.items()}))
I._set_defaults()
