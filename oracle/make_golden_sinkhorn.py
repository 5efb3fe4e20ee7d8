"""Generates tests/golden/sinkhorn_ref.npz by running the REFERENCE's own numpy code
(/root/reference/point_cloud_utils/_sinkhorn.py, loaded by path: the package __init__ needs the compiled module).

    python oracle/make_golden_sinkhorn.py
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_sinkhorn", "/root/reference/point_cloud_utils/_sinkhorn.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(77)
out = {}
cases = []
for tag, dtype, nb, n, m, eps in (("f64_single", np.float64, 0, 100, 100, 1e-3), ("f32_single", np.float32, 0, 64, 96, 1e-2),
                                  ("f64_batch", np.float64, 3, 50, 70, 5e-3), ("f32_batch", np.float32, 2, 128, 128, 2e-2)):
    shape_a = (n, 3) if nb == 0 else (nb, n, 3)
    shape_b = (m, 3) if nb == 0 else (nb, m, 3)
    a = rng.random(shape_a).astype(dtype)
    b = rng.random(shape_b).astype(dtype)
    M = ref.pairwise_distances(a, b)
    wa = rng.random(shape_a[:-1]).astype(dtype) + dtype(0.5)
    wb = rng.random(shape_b[:-1]).astype(dtype) + dtype(0.5)
    wa /= wa.sum(-1, keepdims=True)
    wb /= wb.sum(-1, keepdims=True)
    P = ref.sinkhorn(wa, wb, M, eps)
    out.update({tag + "_a": a, tag + "_b": b, tag + "_M": M, tag + "_wa": wa, tag + "_wb": wb, tag + "_P": P,
                tag + "_eps": np.float64(eps)})
    for p in (1, np.inf, -np.inf, 0, 3):
        out[tag + "_M_p%s" % str(p).replace("-", "neg")] = ref.pairwise_distances(a, b, p)
    cases.append(tag)
p = rng.random((80, 3)); q = rng.random((60, 3))          # test_examples.py:289-310 shapes, float64
emd, P = ref.earth_movers_distance(p, q, eps=1e-3)
out.update(emd_p=p, emd_q=q, emd_value=np.float64(emd), emd_P=P)
out["cases"] = np.array(cases)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sinkhorn_ref.npz"), **out)
print("written", len(out), "arrays; emd", emd)
