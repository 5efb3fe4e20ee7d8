"""Point-cloud normals from k nearest neighbours (SURVEY.md 8f, N1;
/root/reference/src/point_cloud_normals.cpp:115-173, :375-411, point_cloud_utils/_pointcloud_normals.py:4-54).

What is bit-exact: which points are kept and, through the k-NN path, every neighbour set.  What is a tolerance:
the unit normal -- 1 - |n_gpu . n_ref| <= 1e-6 wherever the two smallest singular values of the neighbourhood are
separated (relative gap > 1e-3); without view directions the sign is free, as it is in the reference.
The oracle's SVD is numpy's (Eigen's JacobiSVD is not in the reference tree): parity of the vector is unpinned."""
import numpy as np
import pytest

TOL = 1e-6


def _surface(rng, n, dtype, noise=0.01):
    """Noisy samples of the unit sphere with outward view directions."""
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    pts = v * (1.0 + noise * rng.normal(size=(n, 1)))
    return pts.astype(dtype), v.astype(dtype)


def _gap_ok(points, idx, k, oracle):
    """Rows whose two smallest singular values are well separated (the normal is well defined there)."""
    _, nn = oracle.k_nearest_neighbors(points, points, k, True)
    off = (points[nn] - points[:, None, :]).astype(np.float64)
    s = np.linalg.svd(off, compute_uv=False)
    ok = (s[:, 1] - s[:, 2]) > 1e-3 * s[:, 0]
    return ok[idx]


def test_oracle_conventions(oracle):
    rng = np.random.default_rng(3)
    pts, dirs = _surface(rng, 3000, np.float64, noise=0.002)
    idx, nrm = oracle.estimate_point_cloud_normals_knn(pts, 12)
    assert idx.dtype == np.int64 and np.array_equal(idx, np.arange(3000)) and nrm.shape == (3000, 3)   # tests/test_examples.py:437-438
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-12)
    assert np.mean(np.abs(np.einsum("ij,ij->i", nrm, dirs)) > 0.95) > 0.95       # it is the surface normal
    idx2, nrm2 = oracle.estimate_point_cloud_normals_knn(pts, 12, dirs, np.deg2rad(3.0))
    assert 0 < len(idx2) < 3000 and np.all(np.einsum("ij,ij->i", nrm2, dirs[idx2]) >= np.cos(np.deg2rad(3.0)) - 1e-12)
    assert len(oracle.estimate_point_cloud_normals_knn(pts[:5], 12)[0]) == 0    # fewer than k points: all dropped (:139-142)
    with pytest.raises(ValueError):
        oracle.estimate_point_cloud_normals_knn(pts, 0)


def test_argument_errors(pcu):
    pts = np.random.rand(50, 3)
    with pytest.raises(ValueError, match="number of neighbors"):
        pcu.estimate_point_cloud_normals_knn(pts, 0)
    with pytest.raises(ValueError, match="shape"):
        pcu.estimate_point_cloud_normals_knn(pts[:, :2], 5)
    with pytest.raises(ValueError, match="view directions"):
        pcu.estimate_point_cloud_normals_knn(pts, 5, view_directions=pts[:10])
    with pytest.raises(ValueError):
        pcu.estimate_point_cloud_normals_knn(pts.astype(np.float32), 5, view_directions=pts)
    with pytest.raises(ValueError, match="NumPy array"):
        pcu.estimate_point_cloud_normals_knn(pts.tolist(), 5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("k", [3, 12, 20])
def test_normals_match_the_oracle(pcu, oracle, dtype, k):
    rng = np.random.default_rng(100 + k)
    pts, dirs = _surface(rng, 20000, dtype)
    ref_i, ref_n = oracle.estimate_point_cloud_normals_knn(pts, k)
    got_i, got_n = pcu.estimate_point_cloud_normals_knn(pts, k)
    assert got_i.dtype == np.int64 and got_n.dtype == dtype and got_n.shape == (20000, 3)
    assert np.array_equal(got_i, ref_i)
    assert np.allclose(np.linalg.norm(got_n.astype(np.float64), axis=1), 1.0, atol=1e-6 if dtype == np.float32 else 1e-12)
    ok = _gap_ok(pts, ref_i, k, oracle)
    dots = np.abs(np.einsum("ij,ij->i", got_n.astype(np.float64), ref_n.astype(np.float64)))
    assert ok.mean() > 0.9
    assert np.all(1.0 - dots[ok] <= (TOL if dtype == np.float64 else 1e-5)), float((1.0 - dots[ok]).max())
    # with view directions: oriented towards the sensor, filtered by angle -- same kept set, same vectors
    thr = np.deg2rad(25.0)
    ref_i, ref_n = oracle.estimate_point_cloud_normals_knn(pts, k, dirs, thr)
    got_i, got_n = pcu.estimate_point_cloud_normals_knn(pts, k, dirs, thr)
    cosang = np.einsum("ij,ij->i", got_n.astype(np.float64), dirs[got_i].astype(np.float64))
    assert np.all(cosang >= np.cos(thr) - 1e-6)
    # points within rounding of the threshold may fall on either side; everything else agrees exactly
    sym = np.setxor1d(got_i, ref_i)
    if len(sym):
        allow_i, allow_n = oracle.estimate_point_cloud_normals_knn(pts, k, dirs, np.pi)
        c = np.einsum("ij,ij->i", allow_n.astype(np.float64), dirs[allow_i].astype(np.float64))
        lookup = dict(zip(allow_i.tolist(), c.tolist()))
        assert all(abs(np.arccos(np.clip(lookup[int(i)], -1, 1)) - thr) < 1e-4 for i in sym), sym[:10]
    both = np.intersect1d(got_i, ref_i)
    gn = got_n[np.searchsorted(got_i, both)].astype(np.float64)
    rn = ref_n[np.searchsorted(ref_i, both)].astype(np.float64)
    ok2 = _gap_ok(pts, both, k, oracle)
    assert np.all(1.0 - np.einsum("ij,ij->i", gn, rn)[ok2] <= (TOL if dtype == np.float64 else 1e-5))   # signed now


@pytest.mark.gpu
def test_normals_cuda_tensors_and_small_clouds(pcu, oracle):
    import torch
    rng = np.random.default_rng(7)
    pts, dirs = _surface(rng, 50000, np.float32)
    pt, dt = torch.from_numpy(pts).cuda(), torch.from_numpy(dirs).cuda()
    i, n = pcu.estimate_point_cloud_normals_knn(pt, 12)
    assert i.is_cuda and n.is_cuda and i.dtype == torch.int64 and n.shape == (50000, 3)
    gi, gn = pcu.estimate_point_cloud_normals_knn(pts, 12)
    assert np.array_equal(i.cpu().numpy(), gi) and np.array_equal(n.cpu().numpy(), gn)
    i, n = pcu.estimate_point_cloud_normals_knn(pt, 12, dt, np.deg2rad(15.0))
    gi, gn = pcu.estimate_point_cloud_normals_knn(pts, 12, dirs, np.deg2rad(15.0))
    assert 0 < len(gi) < 50000 and np.array_equal(i.cpu().numpy(), gi) and np.array_equal(n.cpu().numpy(), gn)
    # fewer points than neighbours: everything is dropped; a planar cloud: the normal is the plane's
    i, n = pcu.estimate_point_cloud_normals_knn(pts[:7], 12)
    assert i.shape == (0,) and n.shape == (0, 3)
    flat = np.concatenate([rng.random((4000, 2)), np.full((4000, 1), 0.25)], axis=1)
    i, n = pcu.estimate_point_cloud_normals_knn(flat, 8)
    assert np.array_equal(i, np.arange(4000)) and np.allclose(np.abs(n[:, 2]), 1.0, atol=1e-9)


# ---- the radius variant (src/point_cloud_normals.cpp:48-113, :303-370) -------------------------------------------
def _ball_gap_ok(points, idx, nbrs):
    ok = np.zeros(len(idx), dtype=bool)
    for t, i in enumerate(idx):
        off = (points[nbrs[i]] - points[i]).astype(np.float64)
        s = np.linalg.svd(off, compute_uv=False)
        ok[t] = len(s) == 3 and (s[1] - s[2]) > 1e-3 * s[0]
    return ok


def test_ball_oracle_conventions(oracle):
    rng = np.random.default_rng(5)
    pts, dirs = _surface(rng, 2500, np.float64, noise=0.002)
    # the search radius is a SQUARED distance: radius 0.01 -> points within 0.1
    nbrs, d2 = oracle.ball_neighbourhoods(pts, 0.01)
    i = 17
    true = np.nonzero(np.sum((pts - pts[i]) ** 2, axis=1) < 0.01 * (1 - 1e-12))[0]
    assert set(true.tolist()) <= set(nbrs[i].tolist()) and i in nbrs[i] and np.all(d2[i] < 0.01)
    idx, nrm = oracle.estimate_point_cloud_normals_ball(pts, 0.01)
    assert np.mean(np.abs(np.einsum("ij,ij->i", nrm, dirs[idx])) > 0.95) > 0.95
    counts = np.array([len(j) for j in nbrs])
    cut = int(np.median(counts)) + 1
    few = oracle.estimate_point_cloud_normals_ball(pts, 0.01, min_pts_per_ball=cut)[0]
    assert np.array_equal(few, np.nonzero(counts >= cut)[0]) and 0 < len(few) < 2500
    with pytest.raises(ValueError):
        oracle.estimate_point_cloud_normals_ball(pts, 0.0)
    with pytest.raises(ValueError):
        oracle.estimate_point_cloud_normals_ball(pts, 0.1, min_pts_per_ball=2)


def test_ball_argument_errors(pcu):
    pts = np.random.rand(50, 3)
    with pytest.raises(ValueError, match="radius"):
        pcu.estimate_point_cloud_normals_ball(pts, 0.0)
    with pytest.raises(ValueError, match="min_pts_per_ball"):
        pcu.estimate_point_cloud_normals_ball(pts, 0.1, min_pts_per_ball=2)
    with pytest.raises(ValueError, match="max_pts_per_ball"):
        pcu.estimate_point_cloud_normals_ball(pts, 0.1, max_pts_per_ball=2)
    with pytest.raises(ValueError, match="weight_function"):
        pcu.estimate_point_cloud_normals_ball(pts, 0.1, weight_function="gauss")
    with pytest.raises(ValueError, match="shape"):
        pcu.estimate_point_cloud_normals_ball(pts[:, :2], 0.1)
    with pytest.raises(ValueError, match="view directions"):
        pcu.estimate_point_cloud_normals_ball(pts, 0.1, view_directions=pts[:10])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("weight", ["constant", "rbf"])
def test_ball_normals_match_the_oracle(pcu, oracle, dtype, weight):
    rng = np.random.default_rng(11)
    pts, dirs = _surface(rng, 6000, dtype)
    radius = 0.01                                   # reach sqrt(0.01) = 0.1: ~ 15 neighbours on this sphere
    nbrs, _ = oracle.ball_neighbourhoods(pts, radius)
    counts = np.array([len(j) for j in nbrs])
    min_pts = int(np.percentile(counts, 30))        # a threshold that really drops points
    ref_i, ref_n = oracle.estimate_point_cloud_normals_ball(pts, radius, min_pts_per_ball=min_pts, weight_function=weight)
    got_i, got_n = pcu.estimate_point_cloud_normals_ball(pts, radius, min_pts_per_ball=min_pts, weight_function=weight)
    assert got_i.dtype == np.int64 and got_n.dtype == dtype
    assert 0 < len(ref_i) < 6000 and np.array_equal(got_i, ref_i)          # neighbour counts are exact
    assert np.array_equal(ref_i, np.nonzero(counts >= min_pts)[0])
    ok = _ball_gap_ok(pts, ref_i, nbrs)
    dots = np.abs(np.einsum("ij,ij->i", got_n.astype(np.float64), ref_n.astype(np.float64)))
    assert ok.mean() > 0.8
    assert np.all(1.0 - dots[ok] <= (TOL if dtype == np.float64 else 1e-5)), float((1.0 - dots[ok]).max())
    # view directions: oriented and filtered
    thr = np.deg2rad(30.0)
    ref_i, ref_n = oracle.estimate_point_cloud_normals_ball(pts, radius, dirs, thr, min_pts, weight_function=weight)
    got_i, got_n = pcu.estimate_point_cloud_normals_ball(pts, radius, dirs, thr, min_pts, weight_function=weight)
    cosang = np.einsum("ij,ij->i", got_n.astype(np.float64), dirs[got_i].astype(np.float64))
    assert np.all(cosang >= np.cos(thr) - 1e-6)
    sym = np.setxor1d(got_i, ref_i)
    assert len(sym) <= 3                            # only points within rounding of the threshold may differ
    both = np.intersect1d(got_i, ref_i)
    gn = got_n[np.searchsorted(got_i, both)].astype(np.float64)
    rn = ref_n[np.searchsorted(ref_i, both)].astype(np.float64)
    ok2 = _ball_gap_ok(pts, both, nbrs)
    assert np.all(1.0 - np.einsum("ij,ij->i", gn, rn)[ok2] <= (TOL if dtype == np.float64 else 1e-5))


@pytest.mark.gpu
def test_ball_normals_radius_extremes_and_offsets(pcu, oracle):
    """Radius far below the point spacing (every ball holds the point alone: all dropped), radius covering the whole
    cloud (every point sees every point), a cloud far from the origin (outward-rounded cell box) and exact ties on the
    ball's boundary (strict `<` as in RadiusResultSet::addPoint)."""
    rng = np.random.default_rng(2)
    pts = rng.random((3000, 3)).astype(np.float32)
    i, n = pcu.estimate_point_cloud_normals_ball(pts, 1e-12)
    assert i.shape == (0,) and n.shape == (0, 3)
    i, n = pcu.estimate_point_cloud_normals_ball(pts[:400], 10.0)
    ri, rn = oracle.estimate_point_cloud_normals_ball(pts[:400], 10.0)
    assert np.array_equal(i, ri) and np.array_equal(i, np.arange(400))
    # every point fits the same matrix up to its own centre: compare a few
    assert np.all(1.0 - np.abs(np.einsum("ij,ij->i", n.astype(np.float64), rn.astype(np.float64))) < 1e-4)
    far = (pts + np.float32(4096.0)).astype(np.float32)          # ulp(4096) = 4.9e-4, comparable to the reach below
    for radius in (1e-5, 3e-4):
        nbrs, _ = oracle.ball_neighbourhoods(far, radius)
        counts = np.array([len(j) for j in nbrs])
        got = pcu.estimate_point_cloud_normals_ball(far, radius, min_pts_per_ball=3)[0]
        assert np.array_equal(got, np.nonzero(counts >= 3)[0])
    # lattice: neighbours at squared distance exactly 1/16 must be excluded by radius = 1/16, included just above it
    g = np.stack(np.meshgrid(*[np.arange(8)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float64) * 0.25
    inner = np.all((g > 0.1) & (g < 1.6), axis=1)
    got = pcu.estimate_point_cloud_normals_ball(g, 0.0625, min_pts_per_ball=3)[0]
    assert got.shape == (0,)                                       # only the point itself is strictly inside
    got = pcu.estimate_point_cloud_normals_ball(g, np.nextafter(0.0625, 1.0), min_pts_per_ball=7)[0]
    assert np.array_equal(got, np.nonzero(inner)[0])               # 6 face neighbours + itself


@pytest.mark.gpu
def test_ball_normals_subset_and_tensors(pcu, oracle):
    """max_pts_per_ball: a random subset of exactly that many neighbours is fitted (reproducible through NumPy's seed);
    the normals stay close to the full fit on a smooth surface.  CUDA tensors give the numpy results."""
    import torch
    rng = np.random.default_rng(9)
    pts, dirs = _surface(rng, 20000, np.float32, noise=0.001)
    radius = 0.01
    full_i, full_n = pcu.estimate_point_cloud_normals_ball(pts, radius, min_pts_per_ball=10)
    np.random.seed(123)
    a_i, a_n = pcu.estimate_point_cloud_normals_ball(pts, radius, min_pts_per_ball=10, max_pts_per_ball=8)
    np.random.seed(123)
    b_i, b_n = pcu.estimate_point_cloud_normals_ball(pts, radius, min_pts_per_ball=10, max_pts_per_ball=8)
    c_i, c_n = pcu.estimate_point_cloud_normals_ball(pts, radius, min_pts_per_ball=10, max_pts_per_ball=8)
    assert np.array_equal(a_i, full_i) and np.array_equal(a_i, b_i) and np.array_equal(a_n, b_n)
    assert not np.array_equal(a_n, c_n)                            # another seed, another subset
    assert not np.array_equal(a_n, full_n)
    d = np.abs(np.einsum("ij,ij->i", a_n.astype(np.float64), full_n.astype(np.float64)))
    assert np.median(d) > 0.99
    # a cap above every neighbourhood size changes nothing
    big_i, big_n = pcu.estimate_point_cloud_normals_ball(pts, radius, min_pts_per_ball=10, max_pts_per_ball=100000)
    assert np.array_equal(big_i, full_i) and np.array_equal(big_n, full_n)
    pt, dt = torch.from_numpy(pts).cuda(), torch.from_numpy(dirs).cuda()
    ti, tn = pcu.estimate_point_cloud_normals_ball(pt, radius, dt, np.deg2rad(1.0), 10, weight_function="rbf")
    gi, gn = pcu.estimate_point_cloud_normals_ball(pts, radius, dirs, np.deg2rad(1.0), 10, weight_function="rbf")
    assert ti.is_cuda and 0 < len(gi) <= 20000
    assert np.array_equal(ti.cpu().numpy(), gi) and np.array_equal(tn.cpu().numpy(), gn)
