"""Generates tests/golden/morton_*.npz from the reference's OWN MortonCode64 (oracle/_ref/libpcu_ref_morton.so,
built by oracle/Makefile from /root/reference/src/common/morton_code.cpp).  Run in the build container:

    python oracle/make_golden_morton.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

O.build()
assert O.have_morton_reference(), "needs /root/reference"
rng = np.random.default_rng(2024)
lim = 1 << 20
pts = np.concatenate([
    rng.integers(-lim, lim, (4000, 3)),
    rng.integers(0, 1000, (2000, 3)),                      # the reference's own test range (tests/test_examples.py:455-456)
    np.array([[0, 0, 0], [-1, -1, -1], [lim - 1, lim - 1, lim - 1], [-lim, -lim, -lim], [1, 0, 0], [0, 1, 0], [0, 0, 1],
              [-lim, lim - 1, 0], [5, -7, 11]]),
]).astype(np.int32)
codes = O.morton_encode(pts, impl="reference")
other = O.morton_encode(rng.integers(-1000, 1000, pts.shape).astype(np.int32), impl="reference")
data = rng.integers(0, 1000, (5000, 3)).astype(np.int32)
sorted_codes = np.sort(O.morton_encode(data, impl="reference"))
queries = O.morton_encode(np.concatenate([rng.integers(0, 1000, (600, 3)), rng.integers(-50, 1100, (200, 3)),
                                          np.array([[0, 0, 0], [999, 999, 999], [-lim, -lim, -lim], [lim - 1, lim - 1, lim - 1]])]).astype(np.int32),
                          impl="reference")
out = dict(pts=pts, codes=codes, decoded=O.morton_decode(codes, impl="reference"), other=other,
           added=O.morton_add(codes, other, impl="reference"), subtracted=O.morton_subtract(codes, other, impl="reference"),
           sorted_codes=sorted_codes, queries=queries)
for k in (1, 2, 7, 16):
    out["window_k%d" % k] = O.morton_knn(sorted_codes, queries, k, sort_dist=False, impl="reference")
tiny = np.sort(O.morton_encode(rng.integers(0, 1000, (10, 3)).astype(np.int32), impl="reference"))
out["tiny_codes"] = tiny
out["tiny_window_k15"] = O.morton_knn(tiny, queries, 15, sort_dist=False, impl="reference")     # k > n: (m, 10), test_examples.py:506-508
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "morton_ref.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
