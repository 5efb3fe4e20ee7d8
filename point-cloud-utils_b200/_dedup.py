"""deduplicate_point_cloud / deduplicate_mesh_vertices -- the reference's names, arguments and return conventions
(/root/reference/src/remove_duplicates.cpp:85-176), computed by csrc/dedup.cuh (keys -> radix sort -> run heads).

numpy arrays in -> numpy arrays out (staged on the GPU through torch: device memory and streams only); CUDA torch
tensors in -> CUDA tensors out.  float32 / float64 points, int32 / int64 faces."""
import numpy as _np

from . import _pcu_internal


def _torch():
    import importlib
    return importlib.import_module("torch")


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _run(points, faces, epsilon, device, pname, fname):
    torch = _torch()
    is_np = not _is_tensor(points)
    if is_np:
        points = _np.asarray(points)
        if points.dtype not in (_np.float32, _np.float64):
            raise ValueError("Invalid scalar type (%s) for argument '%s'. Expected one of ['float32', 'float64']." % (points.dtype, pname))
        if points.ndim != 2 or points.shape[1] != 3:                       # validate_point_cloud, src/common/common.h:68-73
            raise ValueError("Only 3D inputs are supported: v must have shape (n, 3) (n > 0). Got points.shape =" + str(points.shape) + ".")
        if faces is not None:
            faces = _np.asarray(faces)
            if faces.dtype not in (_np.int32, _np.int64):
                raise ValueError("Invalid scalar type (%s) for argument '%s'. Expected one of ['int32', 'int64']." % (faces.dtype, fname))
            if faces.ndim != 2:
                raise ValueError("Invalid shape for faces, must be (m, c) but got " + str(faces.shape))
        dev = torch.device("cuda", _pcu_internal._current_device() if device < 0 else device)
        pts = torch.from_numpy(_np.ascontiguousarray(points)).to(dev)
        fcs = torch.from_numpy(_np.ascontiguousarray(faces)).to(dev) if faces is not None else None
    else:
        if not points.is_cuda or points.dtype not in (torch.float32, torch.float64):
            raise ValueError("points must be a float32 / float64 numpy array or CUDA tensor")
        if points.dim() != 2 or points.shape[1] != 3:
            raise ValueError("Only 3D inputs are supported: v must have shape (n, 3) (n > 0). Got points.shape =" + str(tuple(points.shape)) + ".")
        if device >= 0 and points.device.index != device:
            raise ValueError("device=%d was given but the tensors live on %s" % (device, points.device))
        pts = points.detach().contiguous()
        fcs = None
        if faces is not None:
            if not _is_tensor(faces) or faces.device != pts.device or faces.dtype not in (torch.int32, torch.int64) or faces.dim() != 2:
                raise ValueError("faces must be an (m, c) int32 / int64 tensor on the device of the vertices")
            fcs = faces.detach().contiguous()
    n = pts.shape[0]
    if n == 0:          # the reference accepts an empty cloud and returns empty arrays
        empty_i = torch.zeros(0, dtype=torch.int32, device=pts.device)
        out = (pts.clone(), fcs.clone() if fcs is not None else None, empty_i, empty_i.clone())
    else:
        nf, cols = (fcs.shape[0], fcs.shape[1]) if fcs is not None else (0, 0)
        out_p = torch.empty((n, 3), dtype=pts.dtype, device=pts.device)
        svi = torch.empty(n, dtype=torch.int32, device=pts.device)
        svj = torch.empty(n, dtype=torch.int32, device=pts.device)
        out_f = torch.empty((nf, cols), dtype=fcs.dtype, device=pts.device) if fcs is not None else None
        counts = torch.empty(3, dtype=torch.int64, device=pts.device)
        _pcu_internal._deduplicate_device(pts.dtype == torch.float64, pts.data_ptr(), n, float(epsilon),
                                          fcs.data_ptr() if (fcs is not None and nf) else 0, nf, cols,
                                          bool(fcs is not None and fcs.dtype == torch.int64), out_p.data_ptr(), svi.data_ptr(),
                                          svj.data_ptr(), out_f.data_ptr() if (out_f is not None and nf) else 0, counts.data_ptr(),
                                          pts.device.index or 0, torch.cuda.current_stream(pts.device).cuda_stream)
        u, kept_f, bad = (int(v) for v in counts.tolist())
        if bad:
            raise ValueError("%d faces refer to vertices outside [0, %d)" % (bad, n))
        out = (out_p[:u], out_f[:kept_f] if out_f is not None else None, svi[:u], svj)
    if is_np:
        return tuple(None if t is None else t.cpu().numpy() for t in out)
    return out


def deduplicate_point_cloud(points, epsilon, return_index=True, *, device=None):
    """
    Removes duplicated points from a point cloud where two points are considered the same if their distance is below
    some threshold

    Args:
        x : #x by 3 Matrix of 3D positions
        epsilon: threshold below which two points are considered equal
        return_index: If true, return indices to map between input and output

    Returns:
        x_new : #x_new x 3 Point cloud with duplicates removed
        if return indices is set, this function also returns:
            svi : #x_new x 1 indices so that x_new = x[svi]
            svj : #x x 1 indices so that x = x_new[svj]

    Mirrors /root/reference/src/remove_duplicates.cpp:85-126 (libigl round + unique_rows): with epsilon > 0 two points
    are merged when round(p / epsilon) agrees on every axis, otherwise when they are equal; unique rows come in ascending
    lexicographic order.  svi names the FIRST input row of each cluster (libigl's choice among equal rows is unspecified).
    `device` (keyword only): see `k_nearest_neighbors`.
    """
    x_new, _, svi, svj = _run(points, None, epsilon, -1 if device is None else int(device), "points", "f")
    if return_index:
        return x_new, svi, svj
    return x_new


def deduplicate_mesh_vertices(v, f, epsilon, return_index=True, *, device=None):
    """
    Removes duplicated vertices from a triangle mesh two vertices are considered the same if their distance is below
    some threshold

    Args:
        v : #v by 3 Matrix of mesh vertex 3D positions
        f : #f by 3 Matrix of face (triangle) indices
        epsilon: threshold below which two points are considered equal
        return_index: If true, return indices to map between input and output

    Returns:
        v_out : #v x 3 array of mesh vertices with duplicates removed
        f_out : #f x 3 array of mesh faces corresponding to the deduplicated mesh
        svi : #v_out x 1 indices so that v_out = v[svi] (only returned if return_index is True)
        svj : #v x 1 indices so that v = v_out[svj] (only returned if return_index is True)

    Mirrors /root/reference/src/remove_duplicates.cpp:129-176 and :36-77: faces are re-indexed through svj; a face two of
    whose corners end up on the same vertex is dropped; the others keep their order.
    """
    if f is None:
        raise ValueError("f must be an (m, 3) array of face indices")
    v_out, f_out, svi, svj = _run(v, f, epsilon, -1 if device is None else int(device), "v", "f")
    if return_index:
        return v_out, f_out, svi, svj
    return v_out, f_out
