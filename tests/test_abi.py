"""CPU: the C-ABI library loads, exports every symbol include/pcu_b200.h declares, and the host
logic of the binding (argument validation, error classes) behaves like the reference's binding.
No compute call is made here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pcu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcu_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pcu):
    lib = ctypes.CDLL(os.path.join(ROOT, "point-cloud-utils_b200", "libpcu_b200.so"))
    names = declared_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.pcu_b200_abi_version.restype = ctypes.c_int
    assert lib.pcu_b200_abi_version() == 3
    lib.pcu_b200_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.pcu_b200_last_error(), bytes)


def test_stats_struct_layout(pcu):
    assert pcu._pcu_internal._stats_nbytes() == 80


def test_no_cpu_fallback_without_gpu(pcu):
    if pcu.device_count() > 0:
        pytest.skip("a GPU is present")
    a = np.random.rand(10, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pcu.k_nearest_neighbors(a, a, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pcu.chamfer_distance(a, a)


def test_argument_errors_are_value_errors(pcu):
    a = np.random.rand(10, 3)
    b = np.random.rand(5, 3)
    with pytest.raises(ValueError, match="Invalid value for k"):
        pcu.k_nearest_neighbors(a, b, 0)                      # point_cloud_distance.cpp:133-135
    with pytest.raises(ValueError, match="zero elements"):
        pcu.k_nearest_neighbors(a, b[:0], 1)                  # :136-141
    with pytest.raises(ValueError, match="Only 3D inputs"):
        pcu.k_nearest_neighbors(a[:, :2], b, 1)               # :143-149
    with pytest.raises(ValueError, match="scalar type"):
        pcu.k_nearest_neighbors(a.astype(np.float32), b, 1)   # npe_matches
    with pytest.raises(ValueError, match="scalar type"):
        pcu.k_nearest_neighbors(a.astype(np.int64), b.astype(np.int64), 1)
    with pytest.raises(ValueError, match="zero elements"):
        pcu.one_sided_hausdorff_distance(a[:0], b)            # :195-200
    with pytest.raises(ValueError, match="Only 3D inputs"):
        pcu.one_sided_hausdorff_distance(a, b[:, :1])         # :202-208
    with pytest.raises(ValueError):
        pcu.hausdorff_distance(a, b.astype(np.float32))
    with pytest.raises(ValueError):
        pcu.chamfer_distance(a[:, :2], b)


def _reference_parameters(fn):
    """Positional parameters (the reference's) and keyword-only extras of a public function."""
    import inspect
    sig = inspect.signature(fn)
    positional = [p for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD]
    extras = [p for p in sig.parameters.values() if p.kind == p.KEYWORD_ONLY]
    assert len(positional) + len(extras) == len(sig.parameters)
    return sig, [p.name for p in positional], [p.name for p in extras]


def test_signatures_match_the_reference(pcu):
    """Names, order and defaults of the reference's parameters; the only addition is the keyword-only
    `device` (the reference is CPU code and has no such notion), which never shifts a positional argument."""
    sig, names, extras = _reference_parameters(pcu.k_nearest_neighbors)
    assert names == ["query_points", "dataset_points", "k", "squared_distances", "max_points_per_leaf", "num_threads"]
    assert extras == ["device"] and sig.parameters["device"].default is None
    assert sig.parameters["squared_distances"].default is False
    assert sig.parameters["max_points_per_leaf"].default == 10 and sig.parameters["num_threads"].default == -1
    sig, names, extras = _reference_parameters(pcu.one_sided_hausdorff_distance)
    assert names == ["source", "target", "return_index", "squared_distances", "max_points_per_leaf"]
    assert extras == ["device"]
    assert sig.parameters["return_index"].default is True     # point_cloud_distance.cpp:189
    sig, names, extras = _reference_parameters(pcu.hausdorff_distance)
    assert names == ["x", "y", "return_index", "squared_distances", "max_points_per_leaf"] and extras == ["device"]
    assert sig.parameters["return_index"].default is False    # __init__.py:52
    sig, names, extras = _reference_parameters(pcu.chamfer_distance)
    assert names == ["x", "y", "return_index", "p_norm", "max_points_per_leaf"] and extras == ["device"]
    assert sig.parameters["p_norm"].default == 2


def test_device_argument_is_validated_on_the_host(pcu):
    a = np.random.rand(10, 3)
    for bad in ("cpu", "cuda:x", 1.5, -2):
        with pytest.raises(ValueError, match="device"):
            pcu.chamfer_distance(a, a, device=bad)
    with pytest.raises(ValueError, match="out of range"):
        pcu.k_nearest_neighbors(a, a, 1, device=pcu.device_count() + 7)


def test_current_device_rule(pcu):
    """PCU_B200_DEVICE / LOCAL_RANK only count when they name a visible device; with no GPU the answer is 0."""
    import os
    old = {k: os.environ.pop(k, None) for k in ("PCU_B200_DEVICE", "LOCAL_RANK")}
    try:
        assert pcu.current_device() == 0 or pcu.device_count() > 1
        os.environ["LOCAL_RANK"] = "5"
        os.environ["PCU_B200_DEVICE"] = "not-a-number"
        d = pcu.current_device()
        assert d == (5 if pcu.device_count() > 5 else d) and 0 <= d < max(1, pcu.device_count())
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
