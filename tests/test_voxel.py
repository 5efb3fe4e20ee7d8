"""Voxel-grid down-sampling (/root/reference/point_cloud_utils/__init__.py:123-200,
/root/reference/src/sample_point_cloud.cpp:163-244, :336-367).

Exact: the voxel of every point, the set of output voxels, the number of points in each, the number of rows.
Tolerance: the means -- the reference sums in the array's precision point after point (the oracle does the same), the
kernels accumulate in fp64: 1e-5 relative to the voxel size for float32, 1e-12 for float64.  Row order: the reference's
is that of a std::unordered_map walk (unspecified); rows are compared as sets keyed by their voxel."""
import numpy as np
import pytest


def _by_voxel(rows, voxel_size, lo, oracle):
    """Sort output rows by the voxel their mean lies in (means of a voxel's points lie inside the voxel or on its faces)."""
    key = oracle.voxel_indices(rows.astype(np.float64), voxel_size, lo)
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
    return order


def test_oracle_conventions(oracle):
    rng = np.random.default_rng(1)
    p = rng.random((4000, 3))
    out = oracle.downsample_point_cloud_on_voxel_grid(0.25, p)
    assert out.shape[1] == 3 and 0 < out.shape[0] <= 5 ** 3                      # bounds are padded by half a voxel
    v, a, cnt = oracle.downsample_point_cloud_on_voxel_grid(0.25, p, p * 2.0, return_counts=True)
    assert np.allclose(a, 2.0 * v) and cnt.sum() == 4000
    v3, c3 = oracle.downsample_point_cloud_on_voxel_grid(0.05, p, min_points_per_voxel=3, return_counts=True)
    assert c3.min() >= 3
    with pytest.raises(ValueError):
        oracle.downsample_point_cloud_on_voxel_grid(-1.0, p)
    with pytest.raises(ValueError):
        oracle.downsample_point_cloud_on_voxel_grid(0.1, p, min_bound=(0, 0, 0), max_bound=(0, 1, 1))


def test_argument_errors(pcu):
    p = np.random.rand(100, 3)
    with pytest.raises(ValueError, match="numpy array"):
        pcu.downsample_point_cloud_on_voxel_grid(0.1, p.tolist())
    with pytest.raises(ValueError, match="3-tuple"):
        pcu.downsample_point_cloud_on_voxel_grid((0.1, 0.1), p)
    with pytest.raises(ValueError, match="first dimension"):
        pcu.downsample_point_cloud_on_voxel_grid(0.1, p, p[:50])
    with pytest.raises(ValueError, match="max_bound must be greater"):
        pcu.downsample_point_cloud_on_voxel_grid(0.1, p, min_bound=(0, 0, 0), max_bound=(1, 0, 1))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_downsample_matches_the_oracle(pcu, oracle, dtype):
    from importlib import import_module
    mod = import_module("point-cloud-utils_b200")
    rng = np.random.default_rng(11)
    p = (rng.random((200000, 3)) * np.array([2.0, 1.0, 0.5]) - 0.3).astype(dtype)
    colors = rng.random((200000, 3)).astype(np.float32)
    scalar = rng.random((200000,)).astype(np.float64)
    for size, min_pts in ((0.05, 1), ((0.1, 0.03, 0.07), 2), (0.011, 1)):
        vs = np.array([size] * 3) if np.isscalar(size) else np.array(size)
        lo = np.min(p, axis=0) - vs * 0.5
        got = pcu.downsample_point_cloud_on_voxel_grid(size, p, colors, scalar, min_points_per_voxel=min_pts)
        ref = oracle.downsample_point_cloud_on_voxel_grid(size, p, colors, scalar, min_points_per_voxel=min_pts, return_counts=True)
        assert len(got) == 3 and got[0].dtype == dtype and got[1].dtype == np.float32 and got[2].dtype == np.float64
        assert got[0].shape == ref[0].shape and got[1].shape == ref[1].shape and got[2].shape == ref[2].shape
        # same order by construction (first point of each voxel), so rows can be compared directly ...
        tol = (1e-5 if dtype == np.float32 else 1e-12) * float(vs.max()) * 50        # the reference's own float sums drift with the count
        assert np.abs(got[0].astype(np.float64) - ref[0].astype(np.float64)).max() <= tol
        assert np.abs(got[1].astype(np.float64) - ref[1].astype(np.float64)).max() <= 5e-5
        assert np.abs(got[2] - ref[2]).max() <= 1e-12
        # ... and the integer part exactly: every output row lies in a distinct voxel, the voxels are the oracle's, with its counts
        _, _, counts = mod._voxel_internal(p, np.zeros([0, 0]), vs, lo, np.max(p, axis=0) + vs * 0.5, min_pts, None, return_counts=True)
        assert np.array_equal(counts, ref[3])
        full = oracle.voxel_indices(p, vs, lo)
        uniq, cnt = np.unique(full, axis=0, return_counts=True)
        assert got[0].shape[0] == int((cnt >= min_pts).sum())
    # explicit bounds that do not contain the cloud (no clipping in the reference: indices simply go negative / large)
    got = pcu.downsample_point_cloud_on_voxel_grid(0.1, p, min_bound=(0.0, 0.0, 0.0), max_bound=(1.0, 1.0, 1.0))
    ref = oracle.downsample_point_cloud_on_voxel_grid(0.1, p, min_bound=(0.0, 0.0, 0.0), max_bound=(1.0, 1.0, 1.0))
    assert got.shape == ref.shape and np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= 1e-4
    with pytest.raises(ValueError, match="negative"):
        pcu.downsample_point_cloud_on_voxel_grid(-0.1, p, min_bound=(0.0, 0.0, 0.0), max_bound=(1.0, 1.0, 1.0))
    with pytest.raises(ValueError, match="too small"):
        pcu.downsample_point_cloud_on_voxel_grid(1e-12, p)


@pytest.mark.gpu
def test_downsample_cuda_tensors_and_duplicates(pcu, oracle):
    import torch
    rng = np.random.default_rng(12)
    base = rng.random((5000, 3)).astype(np.float32)
    p = np.concatenate([base, base, base[:1000]])                 # exact duplicates share voxels
    pt = torch.from_numpy(p).cuda()
    v = pcu.downsample_point_cloud_on_voxel_grid(0.02, pt)
    vn = pcu.downsample_point_cloud_on_voxel_grid(0.02, p)
    assert v.is_cuda and np.array_equal(v.cpu().numpy(), vn)
    ref = oracle.downsample_point_cloud_on_voxel_grid(0.02, p)
    assert vn.shape == ref.shape and np.abs(vn - ref).max() <= 1e-5
    v, a = pcu.downsample_point_cloud_on_voxel_grid(0.02, pt, pt * 3.0)
    assert torch.allclose(a, v * 3.0, atol=1e-5)
    one = pcu.downsample_point_cloud_on_voxel_grid(10.0, p)       # everything in one voxel: the centroid
    assert one.shape == (1, 3) and np.allclose(one[0], p.astype(np.float64).mean(0), atol=1e-5)
