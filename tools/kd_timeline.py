"""Timeline of the kd-tree replica's build (globaltimer stamps inside the cooperative kernel) and the replay stage:
python tools/kd_timeline.py [m]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pcu_b200 as pcu
I = pcu._pcu_internal
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dev = 0
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
d = torch.rand((m, 3), generator=g, device="cuda")
k = 16
for n, dup in ((1000000, 0), (10000000, 0), (1000000, 5000)):
    q = torch.rand((n, 3), generator=g, device="cuda")
    dd = d.clone()
    if dup:
        dd[:dup] = dd[dup:2 * dup]          # duplicated dataset points: many tied rows
    dists = torch.empty((n, k), dtype=torch.float32, device="cuda"); corrs = torch.empty((n, k), dtype=torch.int64, device="cuda")
    tied = torch.zeros(1, dtype=torch.int64, device="cuda")
    for mode in (0, 2):
        I._set_defaults(disable_tie_replay=mode)
        I._set_profiling(dev, stream, True)
        acc = 0.0
        for r in range(5):
            I._knn_device(False, q.data_ptr(), n, dd.data_ptr(), m, k, False, dists.data_ptr(), corrs.data_ptr(), tied.data_ptr(), 10, dev, stream)
            torch.cuda.synchronize()
            if r >= 2:
                acc += I._last_profile(dev, stream).get("finalize", 0.0) / 3
        I._set_profiling(dev, stream, False)
        t = I._debug_kd_times(dev, stream)
        lv = min(int(t[32]), 28)
        print("n=%8d dup=%5d mode %d: tied %6d replay stage %.3f ms | last build: set-up %.0f us, %d grid-wide levels %s = %.0f us, %d single-CTA subtrees %.0f us"
              % (n, dup, mode, int(tied.item()), acc, (t[1] - t[0]) / 1e3, lv, [round((t[2 + l] - t[1 + l]) / 1e3) for l in range(lv)],
                 (t[30] - t[0]) / 1e3, t[33], (t[31] - t[30]) / 1e3), flush=True)
    del q, dists, corrs
I._set_defaults()
