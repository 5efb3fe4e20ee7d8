"""Scratch: tie-replay cost (stage 'finalize') and flagged-row count for k-NN calls of several sizes,
pruned build (mode 0) vs full trees only (mode 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import pcu_b200 as pcu
I = pcu._pcu_internal
dev = 0
stream = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
m = 1000000
d = torch.rand((m, 3), generator=g, device="cuda")
for n, k in ((1000000, 16), (3000000, 16), (10000000, 16), (1000000, 4), (1000000, 32)):
    q = torch.rand((n, 3), generator=g, device="cuda")
    dists = torch.empty((n, k), dtype=torch.float32, device="cuda")
    corrs = torch.empty((n, k), dtype=torch.int64, device="cuda")
    tied = torch.zeros(1, dtype=torch.int64, device="cuda")
    for mode in (0, 2):
        I._set_defaults(disable_tie_replay=mode)
        I._set_profiling(dev, stream, True)
        acc = {}
        reps = 5
        for r in range(reps + 2):
            I._knn_device(False, q.data_ptr(), n, d.data_ptr(), m, k, False, dists.data_ptr(), corrs.data_ptr(), tied.data_ptr(), 10, dev, stream)
            torch.cuda.synchronize()
            if r >= 2:
                for key, val in I._last_profile(dev, stream).items():
                    acc[key] = acc.get(key, 0.0) + val / reps
        I._set_profiling(dev, stream, False)
        print("n=%8d k=%2d mode %d: tied %4d  search %.3f ms  replay %.3f ms" % (n, k, mode, int(tied.item()), acc.get("search", 0), acc.get("finalize", 0)), flush=True)
    del q, dists, corrs
I._set_defaults()
