mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2h_pytest.log 2>&1
tail -12 gpurun_out/r2h_pytest.log
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import pcu_b200 as pcu
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand((1000000, 3), generator=g, device="cuda")
for k in (12,):
    for _ in range(3): i, n = pcu.estimate_point_cloud_normals_knn(x, k)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): i, n = pcu.estimate_point_cloud_normals_knn(x, k)
    torch.cuda.synchronize(); print("normals k=%d on 1e6 points: %.3f ms" % (k, (time.perf_counter() - t0) / 5 * 1e3))
# k-NN occupancy sweep for k = 16 (2e6 queries vs 1e6 points)
I = pcu._pcu_internal
q = torch.rand((2000000, 3), generator=g, device="cuda"); d = torch.rand((1000000, 3), generator=g, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for occ in (3.0, 4.0, 6.0, 8.0, 12.0):
    I._set_defaults(cell_occupancy=occ)
    for _ in range(2): pcu.k_nearest_neighbors(q, d, 16)
    I._set_profiling(0, stream, True)
    acc = {}
    for _ in range(3):
        pcu.k_nearest_neighbors(q, d, 16); torch.cuda.synchronize()
        for kk, v in I._last_profile(0, stream).items(): acc[kk] = acc.get(kk, 0.0) + v / 3
    I._set_profiling(0, stream, False)
    print("k=16 occupancy %.1f:" % occ, {kk: round(v, 3) for kk, v in acc.items()})
I._set_defaults()
PY
