"""GPU: BASELINE.json configs[3] (C4: k = 16, 1e7 queries vs 1e6 points) and configs[4] (C5: 1024 pairs of
2 x 65536 points) at their FULL sizes, the C-ABI compute entry points driven through ctypes exactly as
INTEGRATION.md section 3 binds them, and the device-selection / stream-ordering rules of the host API.

Full-size checks use size-independent properties on the device plus the reference's own nanoflann path
(oracle/_ref through the `oracle` fixture) on random subsets that the host finishes in seconds."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-6
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_d2(q, p):
    """((qx-px)^2 + (qy-py)^2) + (qz-pz)^2 with every operation rounded separately (eager torch ops do not fuse)."""
    diff = q - p
    sq = diff * diff
    return (sq[..., 0] + sq[..., 1]) + sq[..., 2]


def test_c4_full_size_k16(pcu, oracle):
    """k_nearest_neighbors k = 16 on 1e7 x 3 queries vs 1e6 x 3 points, fp32 (1.92 GB of results)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(4)
    n, m, k = 10_000_000, 1_000_000, 16
    q = torch.rand((n, 3), generator=g, device="cuda")
    d = torch.rand((m, 3), generator=g, device="cuda")
    dist, idx = pcu.k_nearest_neighbors(q, d, k)
    assert dist.shape == (n, k) and idx.shape == (n, k) and idx.dtype == torch.int64
    assert int(idx.min()) >= 0 and int(idx.max()) < m
    # (a) every returned distance is the distance to the returned index, in the reference's rounding, and
    # (b) rows ascend; checked over all 1.6e8 entries, a slab of rows at a time
    step = 1_000_000
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        d2 = _reference_d2(q[lo:hi, None, :], d[idx[lo:hi]])
        assert torch.equal(dist[lo:hi], torch.sqrt(d2)), "distance != distance to the returned index (rows %d..%d)" % (lo, hi)
        assert bool((dist[lo:hi, 1:] >= dist[lo:hi, :-1]).all()), "row not ascending"
        same = idx[lo:hi, 1:] == idx[lo:hi, :-1]
        assert not bool(same.any()), "an index appears twice in a row"
    # (c) nothing nearer was missed: brute force in fp64 on random rows (the 16 smallest of 1e6 distances)
    rows = torch.randperm(n, generator=g, device="cuda")[:1024]
    brute = torch.cdist(q[rows].double(), d.double()).topk(k, dim=1, largest=False).values
    assert bool(((dist[rows].double() - brute).abs() <= 1e-6).all())
    # (d) the reference itself (nanoflann, same dataset) on a random 50 000-row subset: every index and distance
    sub = torch.randperm(n, generator=g, device="cuda")[:50_000]
    ref_d, ref_i = oracle.k_nearest_neighbors(q[sub].cpu().numpy(), d.cpu().numpy(), k)
    assert np.array_equal(idx[sub].cpu().numpy(), ref_i)
    assert np.array_equal(dist[sub].cpu().numpy(), ref_d)
    # rows at the very end of the output (row * k indexing beyond 2^27 elements)
    tail = torch.arange(n - 2000, n, device="cuda")
    ref_d, ref_i = oracle.k_nearest_neighbors(q[tail].cpu().numpy(), d.cpu().numpy(), k)
    assert np.array_equal(idx[tail].cpu().numpy(), ref_i) and np.array_equal(dist[tail].cpu().numpy(), ref_d)


def test_c5_full_size_batched_chamfer(pcu, oracle):
    """batched Chamfer on 1024 pairs of 2 x (65536 x 3) fp32: 1.6 GB of input, descriptor strides beyond 4 GB."""
    import importlib
    import torch
    bmod = importlib.import_module("point-cloud-utils_b200._batched")
    g = torch.Generator(device="cuda").manual_seed(5)
    B, n = 1024, 65536
    x = torch.rand((B, n, 3), generator=g, device="cuda")
    y = torch.rand((B, n, 3), generator=g, device="cuda")
    vals, total = bmod.batched_chamfer(x, y, return_sum=True)
    assert vals.shape == (B,) and vals.dtype == torch.float32 and total.dtype == torch.float64
    vh = vals.cpu().numpy()
    assert np.all(np.isfinite(vh)) and np.all(vh > 0)
    # per-pair value against the reference for 16 pairs, the first and the last among them
    rng = np.random.default_rng(55)
    pairs = sorted(set([0, B - 1] + [int(p) for p in rng.choice(B, 14, replace=False)]))
    for p in pairs:
        ref = float(oracle.chamfer_distance(x[p].cpu().numpy(), y[p].cpu().numpy()))
        assert abs(float(vh[p]) - ref) <= REL * ref, (p, float(vh[p]), ref)
        one = float(pcu.chamfer_distance(x[p], y[p]))
        assert abs(float(vh[p]) - one) <= 1e-7 * ref, (p, float(vh[p]), one)
    # the device-side fp64 sum (what a multi-GPU caller all-reduces) is the sum of the per-pair values
    expect = float(vh.astype(np.float64).sum())
    assert abs(float(total) - expect) <= 1e-12 * expect


def test_batched_sum_over_more_than_one_slice(pcu, oracle):
    """More pairs than one launch addresses (gridDim.y): the batch is cut in slices and the fp64 sum runs on."""
    import importlib
    import torch
    bmod = importlib.import_module("point-cloud-utils_b200._batched")
    g = torch.Generator(device="cuda").manual_seed(6)
    B = 20_000
    x = torch.rand((B, 24, 3), generator=g, device="cuda")
    y = torch.rand((B, 17, 3), generator=g, device="cuda")
    vals, total = bmod.batched_chamfer(x, y, return_sum=True)
    vh = vals.cpu().numpy()
    assert abs(float(total) - float(vh.astype(np.float64).sum())) <= 1e-12 * float(total)
    for p in (0, 16383, 16384, B - 1):
        ref = float(oracle.chamfer_distance(x[p].cpu().numpy(), y[p].cpu().numpy()))
        assert abs(float(vh[p]) - ref) <= REL * ref


# ---- the C ABI through ctypes, exactly as INTEGRATION.md section 3 shows it ---------------------------------
class NNStats(ctypes.Structure):
    _fields_ = [("sum_dist", ctypes.c_double), ("sum_sq_dist", ctypes.c_double), ("max_sq_dist", ctypes.c_double),
                ("argmax_query", ctypes.c_int64), ("argmax_data", ctypes.c_int64), ("n_queries", ctypes.c_int64),
                ("n_tied", ctypes.c_int64), ("n_far", ctypes.c_int64), ("witness_tied", ctypes.c_int64),
                ("pair_value", ctypes.c_double)]


def test_ctypes_binding_of_the_host_entry_points(pcu):
    from conftest import load_golden
    lib = ctypes.CDLL(os.path.join(ROOT, "point-cloud-utils_b200", "libpcu_b200.so"))
    lib.pcu_b200_last_error.restype = ctypes.c_char_p
    ws = ctypes.c_void_p()
    assert lib.pcu_b200_workspace_create(0, ctypes.byref(ws)) == 0, lib.pcu_b200_last_error()
    try:
        g = load_golden("metrics_f32_20k_15k")
        x = np.ascontiguousarray(g["x"], np.float32)
        y = np.ascontiguousarray(g["y"], np.float32)
        st = (NNStats * 2)()
        val = ctypes.c_float()
        rc = lib.pcu_b200_chamfer_host_f32(ws, x.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(x)),
                                           y.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(y)), st,
                                           ctypes.byref(val))
        assert rc == 0, lib.pcu_b200_last_error()
        ref = float(g["chamfer_f64"])
        assert abs(val.value - ref) <= REL * ref
        # the two records carry the one-sided Hausdorff results of both directions (squared distances)
        d_xy, i_xy, j_xy = g["one_sided_xy_sq1"]
        d_yx, i_yx, j_yx = g["one_sided_yx_sq1"]
        assert (np.float32(st[0].max_sq_dist), st[0].argmax_query, st[0].argmax_data) == (np.float32(d_xy), int(i_xy), int(j_xy))
        assert (np.float32(st[1].max_sq_dist), st[1].argmax_query, st[1].argmax_data) == (np.float32(d_yx), int(i_yx), int(j_yx))
        assert st[0].n_queries == len(x) and st[1].n_queries == len(y)
        # k nearest neighbours through the same handle
        g = load_golden("uniform_f32_5k_50k")
        q = np.ascontiguousarray(g["query"], np.float32)
        d = np.ascontiguousarray(g["dataset"], np.float32)
        k = int(g["ks"][-1])
        leaf = int(g["leafs"][0])
        dist = np.empty((len(q), k), np.float32)
        idx = np.empty((len(q), k), np.int64)
        tied = ctypes.c_int64(-1)

        class Options(ctypes.Structure):
            _fields_ = [("max_points_per_leaf", ctypes.c_int), ("cell_occupancy", ctypes.c_float),
                        ("disable_tie_replay", ctypes.c_int), ("binning", ctypes.c_int), ("host_staging", ctypes.c_int)]
        opt = Options(leaf, 0.0, 0, 0, 0)
        assert lib.pcu_b200_workspace_set_options(ws, ctypes.byref(opt)) == 0
        rc = lib.pcu_b200_knn_host_f32(ws, q.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(q)),
                                       d.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(d)), k, 0,
                                       dist.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.byref(tied))
        assert rc == 0, lib.pcu_b200_last_error()
        tag = "k%d_leaf%d_sq0" % (k, leaf)
        assert np.array_equal(idx.squeeze(), g["idx_" + tag]) and np.array_equal(dist.squeeze(), g["dist_" + tag])
        assert tied.value >= 0
        # argument errors come back as status 1 with a message, never as a crash
        rc = lib.pcu_b200_knn_host_f32(ws, q.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(q)),
                                       d.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(d)), 0, 0,
                                       dist.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p), None)
        assert rc == 1 and b"Invalid value for k" in lib.pcu_b200_last_error()
        lib.pcu_b200_workspace_device.restype = ctypes.c_int
        assert lib.pcu_b200_workspace_device(ws) == 0
    finally:
        assert lib.pcu_b200_workspace_destroy(ws) == 0


# ---- which GPU the numpy path runs on; what a call leaves behind ------------------------------------------
def test_numpy_path_follows_the_current_device(pcu, oracle):
    """The reference's interface has no device argument: numpy calls run on torch's current device (or the one
    named by the keyword-only `device=`), never silently on GPU 0."""
    import torch
    internal = pcu._pcu_internal
    rng = np.random.default_rng(17)
    x = rng.random((30000, 3), dtype=np.float32)
    y = rng.random((20000, 3), dtype=np.float32)
    ref = float(oracle.chamfer_distance(x, y))
    ref_knn = oracle.k_nearest_neighbors(x, y, 3)
    ndev = pcu.device_count()
    target = 1 if ndev >= 2 else 0
    internal._release_workspaces()
    with torch.cuda.device(target):
        assert pcu.current_device() == target
        c = pcu.chamfer_distance(x, y)
        assert abs(float(c) - ref) <= REL * ref
        assert internal._workspace_exists(target) and internal._workspace_bytes(target) > 0
        if ndev >= 2:
            assert not internal._workspace_exists(0), "the numpy path ran on GPU 0 although cuda:1 is current"
    assert torch.cuda.current_device() == 0
    # explicit keyword, whatever is current
    internal._release_workspaces()
    got = pcu.k_nearest_neighbors(x, y, 3, device=target)
    assert np.array_equal(got[1], ref_knn[1]) and np.array_equal(got[0], ref_knn[0])
    assert internal._workspace_exists(target) and (ndev < 2 or not internal._workspace_exists(0))
    assert pcu.hausdorff_distance(x, y, return_index=True, device="cuda:%d" % target) == \
        oracle.hausdorff_distance(x, y, return_index=True)
    assert abs(float(pcu.batched_chamfer_distance(x[None], y[None], device=torch.device("cuda", target))[0]) - ref) <= REL * ref
    if ndev >= 2:
        # tensors on cuda:1 while cuda:0 is current: runs on cuda:1 and leaves cuda:0 current
        xt, yt = torch.from_numpy(x).to("cuda:1"), torch.from_numpy(y).to("cuda:1")
        d, i = pcu.k_nearest_neighbors(xt, yt, 3)
        assert d.device.index == 1 and torch.cuda.current_device() == 0
        assert np.array_equal(i.cpu().numpy(), ref_knn[1])
        with pytest.raises(ValueError, match="contradicts"):
            pcu.k_nearest_neighbors(xt, yt, 3, device=0)


def test_device_call_then_numpy_call_without_synchronising(pcu, oracle):
    """A CUDA-tensor call returns without synchronising; a numpy call issued right behind it uses a workspace
    of its own (its stream does not order against the caller's), so neither disturbs the other's scratch."""
    import torch
    rng = np.random.default_rng(23)
    x = rng.random((400000, 3), dtype=np.float32)
    y = rng.random((300000, 3), dtype=np.float32)
    a = rng.random((50000, 3), dtype=np.float32)
    b = rng.random((60000, 3), dtype=np.float32)
    ref_big = oracle.k_nearest_neighbors(x, y, 8)
    ref_small = oracle.hausdorff_distance(a, b, return_index=True)
    ref_cham = float(oracle.chamfer_distance(a, b))
    xt, yt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    torch.cuda.synchronize()
    for _ in range(3):
        d, i = pcu.k_nearest_neighbors(xt, yt, 8)          # enqueued on torch's default stream, still running ...
        h = pcu.hausdorff_distance(a, b, return_index=True)  # ... while the numpy path copies, bins and sweeps
        c = pcu.chamfer_distance(a, b)
        assert h == ref_small and abs(float(c) - ref_cham) <= REL * ref_cham
        assert np.array_equal(i.cpu().numpy(), ref_big[1]) and np.array_equal(d.cpu().numpy(), ref_big[0])


# ---- prepared clouds: a fixed cloud binned once -------------------------------------------------------------
def test_prepared_cloud_gives_the_plain_results(pcu, oracle):
    import torch
    rng = np.random.default_rng(41)
    for dtype in (np.float32, np.float64):
        y = rng.random((150000, 3)).astype(dtype)
        target = pcu.prepare_cloud(y)
        assert len(target) == 150000 and target.dtype == dtype
        yt = torch.from_numpy(y).cuda()
        target_dev = pcu.prepare_cloud(yt)
        for n in (90000, 1, 150000):
            x = (rng.random((n, 3)) * np.array([1.0, 0.8, 1.2])).astype(dtype)
            ref_c = float(oracle.chamfer_distance(x, y))
            ref_h = oracle.hausdorff_distance(x, y, return_index=True)
            ref_o = oracle.one_sided_hausdorff_distance(x, y)
            for t, xin in ((target, x), (target_dev, torch.from_numpy(x).cuda())):
                c = pcu.chamfer_distance(xin, t)
                assert abs(float(c) - ref_c) <= REL * ref_c and abs(float(c) - float(pcu.chamfer_distance(x, y))) <= 1e-7 * ref_c
                assert pcu.hausdorff_distance(xin, t, return_index=True) == ref_h
                assert pcu.one_sided_hausdorff_distance(xin, t) == ref_o
                assert pcu.one_sided_hausdorff_distance(xin, t, False, True) == oracle.one_sided_hausdorff_distance(x, y, False, True)
        # duplicates in the target: the Hausdorff witness is decided by tie order -> the replay runs against the handle's copy
        yd = np.concatenate([y[:5000], y[:5000]])
        td = pcu.prepare_cloud(yd)
        xq = (y[:300] + dtype(0.25)).astype(dtype)
        assert pcu.one_sided_hausdorff_distance(xq, td) == oracle.one_sided_hausdorff_distance(xq, yd)
        assert pcu.hausdorff_distance(torch.from_numpy(xq).cuda(), pcu.prepare_cloud(torch.from_numpy(yd).cuda()), return_index=True) == \
            oracle.hausdorff_distance(xq, yd, return_index=True)
        with pytest.raises(ValueError):
            pcu.chamfer_distance(x.astype(np.float64 if dtype == np.float32 else np.float32), target)
        with pytest.raises(ValueError):
            pcu.chamfer_distance(x, target, return_index=True)
        target.close()
        with pytest.raises(ValueError, match="closed"):
            pcu.chamfer_distance(x, target)


@pytest.mark.gpu
def test_pageable_inputs_through_the_pinned_ring(pcu, oracle):
    """Large ordinary numpy arrays are staged through the workspace's pinned ring by copy threads (csrc/staging.h):
    same results as the driver's own staging, for sizes around the chunk (2 MB) and ring (32 MB) boundaries, odd
    tails, both precisions, back-to-back calls that reuse ring slots still in flight, and a forced ring for small
    inputs."""
    import torch
    I = pcu._pcu_internal
    rng = np.random.default_rng(21)
    try:
        for n, m, dtype in ((349525, 349526, np.float32), (174763, 1500000, np.float32), (3000001, 700000, np.float32),
                            (400000, 380000, np.float64)):
            x = rng.random((n, 3)).astype(dtype); y = rng.random((m, 3)).astype(dtype)
            results = []
            for mode in (2, 0, 1):                      # never / automatic / always
                I._set_defaults(host_staging=mode)
                c = pcu.chamfer_distance(x, y)
                h = pcu.hausdorff_distance(x, y, return_index=True)
                d, i = pcu.k_nearest_neighbors(x[:200000], y, 3)
                results.append((float(c), h, d.copy(), i.copy()))
            for r in results[1:]:
                # (the Chamfer value's last bits depend on the order in which the far pass met its queries)
                assert abs(r[0] - results[0][0]) <= 1e-6 * results[0][0] and r[1] == results[0][1]
                assert np.array_equal(r[2], results[0][2]) and np.array_equal(r[3], results[0][3])
        # the same values as from page-locked inputs (which are never staged)
        I._set_defaults(host_staging=0)
        xp = torch.from_numpy(x).pin_memory(); yp = torch.from_numpy(y).pin_memory()
        assert abs(float(pcu.chamfer_distance(xp.numpy(), yp.numpy())) - results[0][0]) <= 1e-6 * results[0][0]
        # small inputs forced through the ring
        I._set_defaults(host_staging=1)
        a = rng.random((1000, 3)).astype(np.float32); b = rng.random((777, 3)).astype(np.float32)
        ref = oracle.chamfer_distance(a, b)
        assert abs(float(pcu.chamfer_distance(a, b)) - ref) <= 1e-6 * ref
    finally:
        I._set_defaults()


@pytest.mark.gpu
def test_prepared_dataset_for_knn_gives_the_plain_results(pcu, oracle):
    """k_nearest_neighbors(query, prepare_cloud(dataset, k=...)): the handle owns the binned dataset AND the reference
    tree; a call only bins its queries, searches, and replays its tied rows on the handle's tree.  Bit-identical to the
    plain call and to the reference -- also where tie order decides (lattice, duplicated points), for other k than the
    one the handle was prepared for, for another leaf size (the call then builds its own tree), for a handle prepared
    without k, for CUDA tensors."""
    import torch
    rng = np.random.default_rng(77)
    for dtype in (np.float32, np.float64):
        d = rng.random((60000, 3)).astype(dtype)
        d[:3000] = d[3000:6000]                                        # duplicates: ties for every query near them
        q = np.concatenate([rng.random((40000, 3)), d[:2000]]).astype(dtype)
        for k_prep, leaf in ((16, 10), (1, 10), (4, 3)):
            h = pcu.prepare_cloud(d, k=k_prep, max_points_per_leaf=leaf)
            assert h.k == k_prep and h.max_points_per_leaf == leaf
            for k in (1, 4, 16):
                ref = oracle.k_nearest_neighbors(q, d, k, max_points_per_leaf=leaf)
                got = pcu.k_nearest_neighbors(q, h, k, max_points_per_leaf=leaf)
                assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
                plain = pcu.k_nearest_neighbors(q, d, k, max_points_per_leaf=leaf)
                assert np.array_equal(got[1], plain[1]) and np.array_equal(got[0], plain[0])
            # another leaf size than the handle's tree: the call builds its own, results follow the call's leaf size
            ref = oracle.k_nearest_neighbors(q, d, 8, True, 25)
            got = pcu.k_nearest_neighbors(q, h, 8, True, 25)
            assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
            h.close()
        # a handle prepared without k (statistics calls) serves k-NN too
        h0 = pcu.prepare_cloud(d)
        assert h0.k is None
        ref = oracle.k_nearest_neighbors(q, d, 5)
        got = pcu.k_nearest_neighbors(q, h0, 5)
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])
        assert abs(float(pcu.chamfer_distance(q, h0)) - float(oracle.chamfer_distance(q, d))) <= REL * float(oracle.chamfer_distance(q, d))
        # CUDA tensors
        ht = pcu.prepare_cloud(torch.from_numpy(d).cuda(), k=8)
        gd, gi = pcu.k_nearest_neighbors(torch.from_numpy(q).cuda(), ht, 8)
        ref = oracle.k_nearest_neighbors(q, d, 8)
        assert gd.is_cuda and np.array_equal(gi.cpu().numpy(), ref[1]) and np.array_equal(gd.cpu().numpy(), ref[0])
        # ... and the metrics against a k-NN handle
        assert pcu.one_sided_hausdorff_distance(q, ht) == oracle.one_sided_hausdorff_distance(q, d)
        with pytest.raises(ValueError):
            pcu.k_nearest_neighbors(q.astype(np.float64 if dtype == np.float32 else np.float32), h0, 3)
        with pytest.raises(ValueError, match="greater than 0"):
            pcu.k_nearest_neighbors(q, h0, 0)
    lat = np.stack(np.meshgrid(*[np.arange(20)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) / 20
    hl = pcu.prepare_cloud(lat, k=8)
    ql = (lat[::3] + np.float32(0.025)).astype(np.float32)            # cell centres: 8 equidistant lattice points each
    ref = oracle.k_nearest_neighbors(ql, lat, 8)
    got = pcu.k_nearest_neighbors(ql, hl, 8)
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[0], ref[0])


@pytest.mark.gpu
def test_pinned_empty_is_a_dma_source(pcu, oracle):
    rng = np.random.default_rng(3)
    x = pcu.pinned_empty(50000); y = pcu.pinned_empty(40000, np.float64)
    assert x.shape == (50000, 3) and x.dtype == np.float32 and y.dtype == np.float64 and x.flags.c_contiguous
    x[:] = rng.random((50000, 3)); y[:] = rng.random((40000, 3))
    y32 = pcu.pinned_empty(40000); y32[:] = y
    ref = float(oracle.chamfer_distance(np.array(x), np.array(y32)))
    assert abs(float(pcu.chamfer_distance(x, y32)) - ref) <= REL * ref
    with pytest.raises(ValueError):
        pcu.pinned_empty(10, np.int32)


@pytest.mark.gpu
def test_ctypes_host_forms_of_voxel_grid_and_duplicate_removal(pcu, oracle):
    """pcu_b200_voxel_downsample_host_* / pcu_b200_deduplicate_host_* (numpy memory in and out, what an npe binding of
    src/sample_point_cloud.cpp:336-367 / src/remove_duplicates.cpp:108-176 would call) give the Python layer's results."""
    lib = ctypes.CDLL(os.path.join(ROOT, "point-cloud-utils_b200", "libpcu_b200.so"))
    lib.pcu_b200_last_error.restype = ctypes.c_char_p
    ws = ctypes.c_void_p()
    assert lib.pcu_b200_workspace_create(0, ctypes.byref(ws)) == 0, lib.pcu_b200_last_error()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    try:
        rng = np.random.default_rng(12)
        pts = rng.random((40000, 3)).astype(np.float32)
        col = rng.random((40000, 2)).astype(np.float64)
        # voxel grid, with a float64 attribute of two columns and the per-voxel counts
        size = (ctypes.c_double * 3)(0.05, 0.05, 0.05)
        lo64 = pts.min(0).astype(np.float64) - 0.025; hi64 = pts.max(0).astype(np.float64) + 0.025    # the wrapper's default bounds
        lo = (ctypes.c_double * 3)(*lo64); hi = (ctypes.c_double * 3)(*hi64)
        out_p = np.empty((40000, 3), np.float32); out_a = np.empty((40000, 2), np.float64); out_c = np.empty(40000, np.int32)
        rows = ctypes.c_int64(-1)
        rc = lib.pcu_b200_voxel_downsample_host_f32(ws, vp(pts), ctypes.c_int64(40000), vp(col), 2, 1, size, lo, hi, 2,
                                                    vp(out_p), vp(out_a), vp(out_c), ctypes.byref(rows))
        assert rc == 0, lib.pcu_b200_last_error()
        ref_p, ref_a, ref_c = oracle.downsample_point_cloud_on_voxel_grid(0.05, pts, col, min_bound=lo64, max_bound=hi64,
                                                                          min_points_per_voxel=2, return_counts=True)
        m = rows.value
        assert m == len(ref_p) and 0 < m < 40000
        # rows come in the order of each voxel's first point, in the library and in the oracle alike
        assert np.array_equal(out_c[:m], ref_c)
        assert np.allclose(out_p[:m], ref_p, rtol=0, atol=2e-6)          # the oracle sums in float32, the kernel in fp64
        assert np.allclose(out_a[:m], ref_a, rtol=0, atol=1e-12)
        # duplicate removal of a mesh (int64 faces), then of the bare cloud in double precision
        v = np.concatenate([pts[:20000], pts[:5000]]).astype(np.float32)
        f = rng.integers(0, len(v), (30000, 3)).astype(np.int64)
        o_p = np.empty((len(v), 3), np.float32); svi = np.empty(len(v), np.int32); svj = np.empty(len(v), np.int32)
        o_f = np.empty_like(f); counts = (ctypes.c_int64 * 3)()
        rc = lib.pcu_b200_deduplicate_host_f32(ws, vp(v), ctypes.c_int64(len(v)), ctypes.c_double(1e-11), vp(f), ctypes.c_int64(len(f)), 3, 1,
                                               vp(o_p), vp(svi), vp(svj), vp(o_f), counts)
        assert rc == 0, lib.pcu_b200_last_error()
        rv, rf, rsvi, rsvj = oracle.deduplicate_mesh_vertices(v, f, 1e-11)
        assert counts[0] == len(rv) == 20000 and counts[1] == len(rf) and counts[2] == 0
        assert np.array_equal(o_p[:counts[0]], rv) and np.array_equal(svi[:counts[0]], rsvi) and np.array_equal(svj, rsvj)
        assert np.array_equal(o_f[:counts[1]], rf)
        v64 = v.astype(np.float64)
        o_p64 = np.empty((len(v), 3), np.float64)
        rc = lib.pcu_b200_deduplicate_host_f64(ws, vp(v64), ctypes.c_int64(len(v64)), ctypes.c_double(0.0), None, ctypes.c_int64(0), 0, 0,
                                               vp(o_p64), vp(svi), vp(svj), None, counts)
        assert rc == 0, lib.pcu_b200_last_error()
        rv, rsvi, rsvj = oracle.deduplicate_point_cloud(v64, 0.0)
        assert counts[0] == len(rv) and np.array_equal(o_p64[:counts[0]], rv) and np.array_equal(svj, rsvj)
    finally:
        assert lib.pcu_b200_workspace_destroy(ws) == 0
