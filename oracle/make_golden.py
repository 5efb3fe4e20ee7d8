"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the reference's own nanoflann path.

Run in the build container (needs oracle/_ref/libpcu_ref.so, i.e. /root/reference mounted):

    python oracle/make_golden.py

Every fixture stores its inputs and the outputs of ``impl="reference"`` -- the reference's vendored
nanoflann.hpp compiled in place, driven exactly like src/point_cloud_distance.cpp:21-99 / :186-234
and point_cloud_utils/__init__.py:52-120 -- so the tests on the GPU box (where /root/reference does
not exist) compare against what the reference really returns, ties included.  The reference's own
test-suite has no golden vectors for this path (tests/test_examples.py:337-425 uses unseeded
random inputs and property assertions only), hence these.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def knn_case(name, q, d, ks, leafs=(10,), squared=(False,)):
    rec = {"query": q, "dataset": d, "ks": np.array(ks), "leafs": np.array(leafs), "squared": np.array(squared)}
    for k in ks:
        for leaf in leafs:
            for sq in squared:
                dist, idx = O.k_nearest_neighbors(q, d, k, sq, leaf, impl="reference")
                tag = "k%d_leaf%d_sq%d" % (k, leaf, int(sq))
                rec["dist_" + tag] = dist
                rec["idx_" + tag] = idx.astype(np.int32)  # widened back to int64 by the tests
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "ok")


def metric_case(name, x, y):
    rec = {"x": x, "y": y}
    for sq in (0, 1):
        a = O.one_sided_hausdorff_distance(x, y, True, bool(sq), impl="reference")
        b = O.one_sided_hausdorff_distance(y, x, True, bool(sq), impl="reference")
        h = O.hausdorff_distance(x, y, True, bool(sq), impl="reference")
        rec["one_sided_xy_sq%d" % sq] = np.array(a, dtype=np.float64)
        rec["one_sided_yx_sq%d" % sq] = np.array(b, dtype=np.float64)
        rec["hausdorff_sq%d" % sq] = np.array(h, dtype=np.float64)
    c, cxy, cyx = O.chamfer_distance(x, y, True, impl="reference")
    rec["chamfer"] = np.array(c)  # keeps numpy's result dtype (input dtype)
    rec["chamfer_f64"] = np.array(float(c), dtype=np.float64)
    rec["corrs_xy"] = cxy.astype(np.int32)
    rec["corrs_yx"] = cyx.astype(np.int32)
    for p in (1, np.inf):
        rec["chamfer_p%s" % ("inf" if p == np.inf else str(p))] = np.array(
            float(O.chamfer_distance(x, y, False, p_norm=p, impl="reference")))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(name, "ok")


def main():
    if not O.have_reference():
        O.build(quiet=False)
    assert O.have_reference(), "oracle/_ref is required to (re)generate goldens"
    os.makedirs(OUT, exist_ok=True)

    # BASELINE.json configs[0]: two 10k x 3 fp64 uniform clouds, seeds (0, 1)
    x = np.random.default_rng(0).random((10000, 3))
    y = np.random.default_rng(1).random((10000, 3))
    metric_case("c1_metrics_f64_10k", x, y)
    knn_case("c1_knn_f64_10k", x, y, ks=(1, 4))

    # fp32 uniform, dataset larger than query set, k in {1, 3, 16}
    q = np.random.default_rng(2).random((5000, 3), dtype=np.float32)
    d = np.random.default_rng(3).random((50000, 3), dtype=np.float32)
    knn_case("uniform_f32_5k_50k", q, d, ks=(1, 3, 16), squared=(False, True))
    x32 = np.random.default_rng(4).random((20000, 3), dtype=np.float32)
    y32 = np.random.default_rng(5).random((15000, 3), dtype=np.float32)
    metric_case("metrics_f32_20k_15k", x32, y32)

    rng = np.random.default_rng(6)
    for dt, sfx in ((np.float32, "f32"), (np.float64, "f64")):
        # duplicated points: systematic exact ties (order = kd-tree visit order)
        base = rng.random((700, 3)).astype(dt)
        dup = np.concatenate([base, base, base[:150]])
        qq = np.concatenate([base[:400], rng.random((400, 3)).astype(dt)])
        knn_case("duplicates_" + sfx, qq, dup, ks=(1, 2, 5, 16), leafs=(1, 10))
        # integer lattice: massive ties, all arithmetic exact
        g = np.stack(np.meshgrid(*[np.arange(10)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(dt)
        knn_case("lattice_self_" + sfx, g, g, ks=(1, 7, 27), leafs=(10,))
        knn_case("lattice_half_" + sfx, (g[::3] + 0.5).astype(dt), g, ks=(1, 8), leafs=(10, 3))
        # k > m: padded with -1 (src/point_cloud_distance.cpp:90-93)
        knn_case("k_gt_m_" + sfx, rng.random((50, 3)).astype(dt), rng.random((5, 3)).astype(dt), ks=(1, 5, 9))
        # degenerate boxes
        pl = rng.random((3000, 3)).astype(dt); pl[:, 2] = 0.25
        knn_case("planar_" + sfx, rng.random((600, 3)).astype(dt), pl, ks=(1, 6))
        ln = np.zeros((2000, 3), dt); ln[:, 1] = rng.random(2000).astype(dt)
        knn_case("collinear_" + sfx, rng.random((600, 3)).astype(dt), ln, ks=(1, 6))
        one = np.full((300, 3), 0.5, dt)
        knn_case("identical_" + sfx, np.concatenate([one[:10], rng.random((40, 3)).astype(dt)]), one, ks=(1, 4))
        # disjoint bounding boxes, single-point clouds, clustered (non-uniform) data
        knn_case("disjoint_" + sfx, rng.random((500, 3)).astype(dt),
                 (rng.random((4000, 3)) * 0.25 + np.array([3.0, -2.0, 7.0])).astype(dt), ks=(1, 4))
        knn_case("single_dataset_" + sfx, rng.random((64, 3)).astype(dt), rng.random((1, 3)).astype(dt), ks=(1, 2))
        knn_case("single_query_" + sfx, rng.random((1, 3)).astype(dt), rng.random((999, 3)).astype(dt), ks=(1, 16))
        cl = np.concatenate([rng.normal(0.2, 0.01, (3000, 3)), rng.normal(0.8, 0.03, (3000, 3)),
                             rng.random((500, 3))]).astype(dt)
        knn_case("clustered_" + sfx, rng.random((1500, 3)).astype(dt), cl, ks=(1, 10))
        metric_case("metrics_clustered_" + sfx, cl[::2].copy(), (cl[1::2] + np.array([0.05, 0, 0])).astype(dt))
        metric_case("metrics_duplicates_" + sfx, dup, qq)
        # large offsets (coordinates ~1e3, spacing ~1e-3): cell arithmetic under cancellation
        off = (rng.random((4000, 3)) * 0.5 + 1000.0).astype(dt)
        knn_case("offset_" + sfx, (rng.random((800, 3)) * 0.5 + 1000.0).astype(dt), off, ks=(1, 5))


if __name__ == "__main__":
    main()
