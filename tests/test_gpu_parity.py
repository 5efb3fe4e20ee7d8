"""GPU: the CUDA path (through the pybind11 module and the C ABI) against the golden vectors the
reference's own nanoflann path produced, and against the oracle on fresh seeded inputs.

Bars: int64 indices bit-exact, distances bit-exact (same rounded arithmetic), Chamfer / Hausdorff
scalars within 1e-6 relative (BASELINE.json)."""
import numpy as np
import pytest

from conftest import knn_cases, knn_golden_names, load_golden, metric_golden_names

pytestmark = pytest.mark.gpu
REL = 1e-6


@pytest.mark.parametrize("name", knn_golden_names())
def test_knn_matches_reference_goldens(pcu, name):
    g = load_golden(name)
    q, d = g["query"], g["dataset"]
    for k, leaf, sq, dist, idx in knn_cases(g):
        got_d, got_i = pcu.k_nearest_neighbors(q, d, k, squared_distances=sq, max_points_per_leaf=leaf)
        assert got_i.dtype == np.int64 and got_d.dtype == q.dtype
        assert got_i.shape == idx.shape and got_d.shape == dist.shape, (name, k)
        assert np.array_equal(got_d, dist), (name, k, leaf, sq, "distances")
        assert np.array_equal(got_i, idx), (name, k, leaf, sq, "indices")


@pytest.mark.parametrize("name", metric_golden_names())
def test_metrics_match_reference_goldens(pcu, name):
    g = load_golden(name)
    x, y = g["x"], g["y"]
    for sq in (0, 1):
        assert pcu.one_sided_hausdorff_distance(x, y, True, bool(sq)) == tuple(g["one_sided_xy_sq%d" % sq])
        assert pcu.one_sided_hausdorff_distance(y, x, True, bool(sq)) == tuple(g["one_sided_yx_sq%d" % sq])
        assert pcu.one_sided_hausdorff_distance(x, y, False, bool(sq)) == g["one_sided_xy_sq%d" % sq][0]
        assert pcu.hausdorff_distance(x, y, True, bool(sq)) == tuple(g["hausdorff_sq%d" % sq])
        assert pcu.hausdorff_distance(x, y, False, bool(sq)) == g["hausdorff_sq%d" % sq][0]
    c = pcu.chamfer_distance(x, y)
    assert c.dtype == x.dtype
    assert abs(float(c) - float(g["chamfer_f64"])) <= REL * float(g["chamfer_f64"])
    c2, cxy, cyx = pcu.chamfer_distance(x, y, return_index=True)
    assert abs(float(c2) - float(g["chamfer_f64"])) <= REL * float(g["chamfer_f64"])
    assert np.array_equal(cxy, g["corrs_xy"]) and np.array_equal(cyx, g["corrs_yx"])
    for p, key in ((1, "chamfer_p1"), (np.inf, "chamfer_pinf")):
        v = float(pcu.chamfer_distance(x, y, p_norm=p))
        assert abs(v - float(g[key])) <= REL * float(g[key])


def test_reference_test_suite_properties(pcu):
    """The assertions of /root/reference/tests/test_examples.py:337-425, with seeded inputs."""
    rng = np.random.default_rng(123)
    for _ in range(3):
        a, b = rng.random((1000, 3)), rng.random((500, 3))
        k = int(rng.integers(10)) + 1
        d, c = pcu.k_nearest_neighbors(a, b, k)
        assert d.shape == ((1000, k) if k > 1 else (1000,)) and c.shape == d.shape
        if k == 1:
            d, c = d[:, None], c[:, None]
        assert np.all(np.abs(np.linalg.norm(a[:, None, :] - b[c], axis=-1) - d) < 1e-5)
    with pytest.raises(ValueError):
        pcu.k_nearest_neighbors(rng.random((1000, 3)), rng.random((500, 3)), 0)
    a, b = rng.random((100, 3)), rng.random((50, 3))
    d1, c1 = pcu.k_nearest_neighbors(a, b, 3)
    d2, c2 = pcu.k_nearest_neighbors(a, b, 3, squared_distances=True)
    assert np.all(c1 == c2) and np.all(np.abs(d1 ** 2.0 - d2) < 1e-5)
    a, b = rng.random((1000, 3)), rng.random((500, 3))
    h_ab, ia1, ib1 = pcu.one_sided_hausdorff_distance(a, b, return_index=True)
    h_ba, ib2, ia2 = pcu.one_sided_hausdorff_distance(b, a, return_index=True)
    h, i1, i2 = pcu.hausdorff_distance(a, b, return_index=True)
    assert h == max(h_ab, h_ba)
    assert abs(h - np.linalg.norm(a[i1] - b[i2])) < 1e-7
    assert (i1, i2) == ((ia1, ib1) if h_ab > h_ba else (ia2, ib2))
    pcu.chamfer_distance(a[:100], b[:100])
    c, cab, cba = pcu.chamfer_distance(a[:100], b[:100], return_index=True)
    assert cab.shape == (100,) and cba.shape == (100,)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [1, 2, 7, 16, 32, 40])
def test_knn_against_oracle_fresh(pcu, oracle, dtype, k):
    rng = np.random.default_rng(1000 + k)
    q = rng.random((20000, 3)).astype(dtype)
    d = rng.random((30000, 3)).astype(dtype)
    got_d, got_i = pcu.k_nearest_neighbors(q, d, k)
    ref_d, ref_i = oracle.k_nearest_neighbors(q, d, k)
    assert np.array_equal(got_i, ref_i)
    assert np.array_equal(got_d, ref_d)


def test_storage_orders_and_views(pcu, oracle):
    rng = np.random.default_rng(5)
    q = np.asfortranarray(rng.random((3000, 3), dtype=np.float32))
    big = rng.random((8000, 6), dtype=np.float32)
    d = big[::2, 1:4]          # strided, non-contiguous view
    got = pcu.k_nearest_neighbors(q, d, 3)
    ref = oracle.k_nearest_neighbors(np.ascontiguousarray(q), np.ascontiguousarray(d), 3)
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])


def test_torch_cuda_inputs_stay_on_device(pcu, oracle):
    import torch
    rng = np.random.default_rng(9)
    x = rng.random((50000, 3), dtype=np.float32)
    y = rng.random((40000, 3), dtype=np.float32)
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    d, i = pcu.k_nearest_neighbors(xt, yt, 1)
    assert d.is_cuda and i.is_cuda and i.dtype == torch.int64 and d.shape == (50000,)
    rd, ri = oracle.k_nearest_neighbors(x, y, 1)
    assert np.array_equal(i.cpu().numpy(), ri) and np.array_equal(d.cpu().numpy(), rd)
    c = pcu.chamfer_distance(xt, yt)
    assert c.is_cuda and c.dtype == torch.float32 and c.dim() == 0
    ref = float(oracle.chamfer_distance(x, y))
    assert abs(float(c) - ref) <= REL * ref
    assert pcu.hausdorff_distance(xt, yt, return_index=True) == oracle.hausdorff_distance(x, y, return_index=True)
    d16, i16 = pcu.k_nearest_neighbors(xt.double(), yt.double(), 16, squared_distances=True)
    r16 = oracle.k_nearest_neighbors(x.astype(np.float64), y.astype(np.float64), 16, True)
    assert np.array_equal(i16.cpu().numpy(), r16[1]) and np.array_equal(d16.cpu().numpy(), r16[0])


def test_full_size_properties_c2_c3(pcu):
    """BASELINE configs[1] / [2] sizes (2 x 1e6 x 3 fp32): properties that need no CPU sweep."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((1000000, 3), generator=g, device="cuda")
    y = torch.rand((1000000, 3), generator=g, device="cuda")
    d, i = pcu.k_nearest_neighbors(x, y, 1)
    # (a) the returned distance is the distance to the returned index, in the reference's rounding
    diff = x - y[i]
    d2 = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    assert torch.equal(d, torch.sqrt(d2))
    # (b) no point of a random subset of y is closer (brute force on 2048 queries)
    sub = torch.randperm(1000000, generator=g, device="cuda")[:2048]
    brute = torch.cdist(x[sub].double(), y.double()).min(dim=1).values
    assert torch.all((d[sub].double() - brute).abs() <= 1e-6)
    # (c) self-query: every point is its own nearest neighbour at distance 0
    ds, is_ = pcu.k_nearest_neighbors(x, x, 1)
    assert torch.all(ds == 0) and torch.equal(x[is_], x)
    # (d) fused Chamfer equals the mean of the per-point distances from the KNN path
    c = pcu.chamfer_distance(x, y)
    d_yx, _ = pcu.k_nearest_neighbors(y, x, 1)
    ref = d.double().mean() + d_yx.double().mean()
    assert abs(float(c) - float(ref)) <= REL * float(ref)
    # (e) Hausdorff is the max of those distances, witnessed by its index pair
    h, hi, hj = pcu.hausdorff_distance(x, y, return_index=True)
    assert h == max(float(d.max()), float(d_yx.max()))
    assert abs(h - float(torch.linalg.vector_norm(x[hi] - y[hj]))) <= 1e-6


def _walk_equal(a, b, na, nb, path=""):
    """Recursively compare two kd-trees given as dicts of arrays (node numbering may differ)."""
    assert a["feat"][na] == b["feat"][nb], ("split dim", path)
    assert a["first"][na] == b["first"][nb] and a["last"][na] == b["last"][nb], ("range", path)
    if a["feat"][na] < 0:
        return 1
    assert a["div_lo"][na] == b["div_lo"][nb] and a["div_hi"][na] == b["div_hi"][nb], ("planes", path)
    return 1 + _walk_equal(a, b, a["kid0"][na], b["kid0"][nb], path + "L") + \
        _walk_equal(a, b, a["kid1"][na], b["kid1"][nb], path + "R")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_kd_replica_equals_reference_tree(pcu, oracle, dtype):
    """The GPU replica of the reference's kd-tree (used for tie replay): same permutation (vAcc), same
    split planes, same leaves as the oracle's build, which is pinned to nanoflann's."""
    import sys
    sys.setrecursionlimit(10000)
    rng = np.random.default_rng(77)
    base = rng.random((3000, 3)).astype(dtype)
    clouds = {
        "uniform": rng.random((20000, 3)).astype(dtype),
        "duplicates": np.concatenate([base, base, base[:500]]),
        "lattice": np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(dtype),
        "planar": np.concatenate([rng.random((4000, 2)), np.full((4000, 1), 0.25)], axis=1).astype(dtype),
        "identical": np.full((500, 3), 0.5, dtype),
        "clustered": np.concatenate([rng.normal(0.2, 0.01, (3000, 3)), rng.normal(0.8, 0.03, (3000, 3))]).astype(dtype),
        "tiny": rng.random((7, 3)).astype(dtype),
    }
    for name, pts in clouds.items():
        for leaf in (10, 1, 3):
            got = pcu._pcu_internal._debug_kd_tree(pts, leaf)
            ref = oracle.kd_tree(pts, leaf)
            assert np.array_equal(np.asarray(got["order"], dtype=np.int64), ref["order"]), (name, leaf, "order")
            n_ref = _walk_equal(got, ref, 0, 0)
            assert n_ref == len(ref["feat"]) == got["n_nodes"], (name, leaf)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_far_apart_and_clustered_clouds(pcu, oracle, dtype):
    """Queries far from every dataset point (disjoint boxes, tight clusters in a large void): the ring
    walk hands over to the occupancy-pyramid descent; results must not change."""
    rng = np.random.default_rng(31)
    x = rng.random((20000, 3)).astype(dtype)
    y = (rng.random((30000, 3)) * np.array([1.0, 2.0, 0.5]) + np.array([40.0, -25.0, 3.0])).astype(dtype)
    blobs = np.concatenate([rng.normal(c, 0.002, (5000, 3)) for c in ((0.1, 0.1, 0.1), (0.9, 0.2, 0.7), (0.5, 0.95, 0.3))])
    blobs = blobs.astype(dtype)
    for q, d in ((x, y), (y, x), (x, blobs), (blobs, x)):
        got_d, got_i = pcu.k_nearest_neighbors(q, d, 1)
        ref_d, ref_i = oracle.k_nearest_neighbors(q, d, 1)
        assert np.array_equal(got_i, ref_i) and np.array_equal(got_d, ref_d)
    for a, b in ((x, y), (x, blobs)):
        ref = float(oracle.chamfer_distance(a, b))
        assert abs(float(pcu.chamfer_distance(a, b)) - ref) <= REL * ref
        assert pcu.hausdorff_distance(a, b, return_index=True) == oracle.hausdorff_distance(a, b, return_index=True)
    st = pcu._pcu_internal._chamfer_stats(x, y)
    assert st[1]["n_far"] == len(x) and st[2]["n_far"] == len(y)      # every query took the slow path ...
    got = pcu.k_nearest_neighbors(x[:2000], blobs, 5)                 # ... and k > 1 still agrees
    ref = oracle.k_nearest_neighbors(x[:2000], blobs, 5)
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])


@pytest.mark.parametrize("k", [1, 16])
def test_million_point_clouds_against_oracle(pcu, oracle, k):
    """BASELINE configs[1] / [3] shapes at 1e6 x 1e6 fp32: every index and distance against the oracle
    (the host CPU finishes this in seconds with its OpenMP sweep)."""
    rng = np.random.default_rng(2024 + k)
    q = rng.random((1000000 if k == 1 else 250000, 3), dtype=np.float32)   # k = 16: a quarter of the queries keeps the CPU side short
    d = rng.random((1000000, 3), dtype=np.float32)
    got_d, got_i = pcu.k_nearest_neighbors(q, d, k)
    ref_d, ref_i = oracle.k_nearest_neighbors(q, d, k)
    assert np.array_equal(got_i, ref_i)
    assert np.array_equal(got_d, ref_d)
    if k == 1:
        ref = float(oracle.chamfer_distance(q, d))
        assert abs(float(pcu.chamfer_distance(q, d)) - ref) <= REL * ref
        assert pcu.hausdorff_distance(q, d, return_index=True) == oracle.hausdorff_distance(q, d, return_index=True)


@pytest.mark.parametrize("k", [3, 16, 40])
def test_topk_far_from_data(pcu, oracle, k):
    """k > 1 with queries far from every dataset point: ring passes hand over to the pyramid descent
    (k <= 32) or every query descends it (k > 32)."""
    rng = np.random.default_rng(77 + k)
    q = rng.random((3000, 3), dtype=np.float32)
    d = (rng.random((20000, 3)) * 0.3 + np.array([5.0, -3.0, 2.0])).astype(np.float32)
    blobs = np.concatenate([rng.normal(c, 0.003, (4000, 3)) for c in ((0.1, 0.1, 0.1), (0.9, 0.8, 0.7))]).astype(np.float32)
    for data in (d, blobs):
        got = pcu.k_nearest_neighbors(q, data, k)
        ref = oracle.k_nearest_neighbors(q, data, k)
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])


def test_concurrent_python_threads(pcu, oracle):
    """Several Python threads calling at once (the GIL is released inside the native call): results are
    the same as serial calls."""
    import threading
    rng = np.random.default_rng(99)
    clouds = [(rng.random((20000 + 1000 * i, 3), dtype=np.float32), rng.random((15000, 3), dtype=np.float32))
              for i in range(6)]
    expect = [oracle.k_nearest_neighbors(a, b, 3, max_points_per_leaf=10 + i) for i, (a, b) in enumerate(clouds)]
    got = [None] * len(clouds)

    def work(i):
        a, b = clouds[i]
        got[i] = pcu.k_nearest_neighbors(a, b, 3, max_points_per_leaf=10 + i)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(clouds))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for g, e in zip(got, expect):
        assert np.array_equal(g[1], e[1]) and np.array_equal(g[0], e[0])


def test_batched_chamfer_matches_per_pair(pcu, oracle):
    """Batched entry point: every pair equals the single-pair call (and the oracle), including cloud sizes
    whose per-pair offsets are not 16-byte aligned (plain-load path instead of the bulk tile copy)."""
    rng = np.random.default_rng(5)
    for n, m in ((1001, 777), (4096, 2048)):
        x = rng.random((5, n, 3), dtype=np.float32)
        y = rng.random((5, m, 3), dtype=np.float32)
        got = pcu.batched_chamfer_distance(x, y)
        assert got.shape == (5,) and got.dtype == np.float32
        for b in range(5):
            ref = float(oracle.chamfer_distance(x[b], y[b]))
            assert abs(float(got[b]) - ref) <= REL * ref
            assert abs(float(got[b]) - float(pcu.chamfer_distance(x[b], y[b]))) <= 1e-7 * ref


@pytest.mark.parametrize("binning", [1, 2])
def test_both_grid_builds_give_the_reference_results(pcu, oracle, binning):
    """The grid is built either by five grid-wide passes or by one CTA per cloud (counters in shared
    memory); `binning` forces one or the other.  Same inputs, both dtypes, k = 1 / k > 1 / the metrics /
    a batch, sizes on either side of the automatic switch -- every result must be the reference's."""
    internal = pcu._pcu_internal
    internal._set_defaults(binning=binning)
    try:
        rng = np.random.default_rng(4242)
        for dtype, n, m in ((np.float32, 1, 1), (np.float32, 37, 5), (np.float32, 3000, 70000), (np.float64, 50001, 9999),
                            (np.float32, 100000, 100000)):
            q = rng.random((n, 3)).astype(dtype)
            d = (rng.random((m, 3)) * np.array([1.0, 0.25, 2.0])).astype(dtype)
            for k in (1, 6):
                got = pcu.k_nearest_neighbors(q, d, k)
                ref = oracle.k_nearest_neighbors(q, d, k)
                assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0]), (binning, dtype, n, m, k)
            ref = float(oracle.chamfer_distance(q, d))
            assert abs(float(pcu.chamfer_distance(q, d)) - ref) <= REL * ref
            assert pcu.hausdorff_distance(q, d, return_index=True) == oracle.hausdorff_distance(q, d, return_index=True)
        # lattice points: massive ties, duplicate points, boundary cells
        lat = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
        lat = np.concatenate([lat, lat[:100]])
        got = pcu.k_nearest_neighbors(lat + np.float32(0.5), lat, 4)
        ref = oracle.k_nearest_neighbors(lat + np.float32(0.5), lat, 4)
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
        x = rng.random((130, 20000, 3), dtype=np.float32)
        y = rng.random((130, 9000, 3), dtype=np.float32)
        got = pcu.batched_chamfer_distance(x, y)
        for b in (0, 64, 129):
            ref = float(oracle.chamfer_distance(x[b], y[b]))
            assert abs(float(got[b]) - ref) <= REL * ref
    finally:
        internal._set_defaults()


@pytest.mark.parametrize("mode", [0, 2, 3])
def test_tie_replay_modes_agree_with_reference(pcu, oracle, mode):
    """A few queries with two exactly equidistant nearest neighbours (50 flagged rows: fewer than 128, so the
    replay builds the pruned reference tree, with mask bits in more than one word).  mode 0: pruned build; 2: full trees only; 3: pruned with
    zero slack, so walks run into stubs and the device-gated full rebuild answers.  All must give the
    reference's tie order."""
    internal = pcu._pcu_internal
    rng = np.random.default_rng(808)
    data = rng.random((200000, 3), dtype=np.float32)
    q = (np.floor(rng.random((50, 3)) * 1024) / 1024).astype(np.float32)      # multiples of 2^-10: q +- 2^-12 is exact
    delta = np.float32(2.0 ** -12)
    twins = np.concatenate([q + np.array([delta, 0, 0], np.float32), q - np.array([delta, 0, 0], np.float32),
                            q + np.array([0, delta, 0], np.float32)])
    data = np.concatenate([data, twins])
    data = data[rng.permutation(len(data))]
    queries = np.concatenate([rng.random((5000, 3), dtype=np.float32), q])
    internal._set_defaults(disable_tie_replay=mode)
    try:
        for k in (1, 2, 5):
            got = pcu.k_nearest_neighbors(queries, data, k)
            ref = oracle.k_nearest_neighbors(queries, data, k)
            assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0]), (mode, k)
    finally:
        internal._set_defaults()


def _sphere(rng, n, dtype, centre=(0.5, 0.5, 0.5), radius=0.45):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return (v * radius + np.array(centre)).astype(dtype)


def test_grid_refinement_for_surfaces_keeps_results(pcu, oracle):
    """Points on a surface fill few cells of a box-filling grid; from the second call with the same shapes
    on, the library refines the grid (feedback from the fill statistics of the previous call).  Every call,
    refined or not, must return the reference's results; box-filling clouds must not trigger it."""
    internal = pcu._pcu_internal
    rng = np.random.default_rng(606)
    x = _sphere(rng, 120000, np.float32)
    y = _sphere(rng, 90000, np.float32, centre=(0.52, 0.5, 0.49))
    ref_c = float(oracle.chamfer_distance(x, y))
    ref_h = oracle.hausdorff_distance(x, y, return_index=True)
    ref_k = oracle.k_nearest_neighbors(x, y, 4)
    seen = []
    for rep in range(4):
        assert abs(float(pcu.chamfer_distance(x, y)) - ref_c) <= REL * ref_c
        seen.append(internal._grid_refinement())
    assert seen[0] == (1.0, 1.0) and max(seen[-1]) > 2.0, seen
    assert pcu.hausdorff_distance(x, y, return_index=True) == ref_h
    for rep in range(3):
        got = pcu.k_nearest_neighbors(x, y, 4)
        assert np.array_equal(got[1], ref_k[1]) and np.array_equal(got[0], ref_k[0])
    assert internal._grid_refinement()[1] > 1.5
    # box-filling clouds of the same shapes right after: the refined grid is too fine for them and shrinks
    # back to the default over a few calls; results stay the reference's throughout
    u = rng.random((120000, 3), dtype=np.float32)
    w = rng.random((90000, 3), dtype=np.float32)
    ref_k = oracle.k_nearest_neighbors(u, w, 4)
    for rep in range(8):
        got = pcu.k_nearest_neighbors(u, w, 4)
        assert np.array_equal(got[1], ref_k[1]) and np.array_equal(got[0], ref_k[0])
    assert internal._grid_refinement() == (1.0, 1.0)
    for rep in range(3):
        pcu.k_nearest_neighbors(u, w, 16)
    assert internal._grid_refinement() == (1.0, 1.0)
