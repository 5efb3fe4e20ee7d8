#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 nearest-neighbour path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c4|c5]

Metric (BASELINE.json): Chamfer query-points/second on 2 x (1e6 x 3) fp32 uniform-random clouds
(config "C3"): one step = one fused bidirectional Chamfer (+ Hausdorff statistics) of
one pair, grid build included; value = (n + m) * pairs / time.

* N = 1: one pair, inputs resident in HBM before the timed region (`value`), and the same call
  through the numpy-facing API from pinned host buffers with the H2D / D2H copies inside the timed
  region (`e2e`).
* N > 1 (torchrun, one rank per GPU): the path shards over independent pairs -- every rank owns one
  pair of the same shape (weak scaling) and the only exchange is an NCCL all-reduce of the fp64 sum
  of the per-pair Chamfer values.  Timing: CUDA events on the launching stream, max over ranks.
* L2: the 24 MB of inputs fit the 126 MB L2, so a 256 MiB buffer is overwritten before every timed
  step (outside the per-step event interval).
* `roofline`: the dominant kernel (the fused search sweep) timed with CUDA events recorded inside the
  library on the launching stream (pcu_b200_workspace_set_profiling), algorithmic bytes = 24 B per
  query point, peak = MEASURED_PEAKS.json hbm_gbs.
* `cpu_baseline` / `--impl reference`: the reference's own nanoflann path (oracle/_ref, compiled from
  /root/reference/external/nanoflann in place) driven like point_cloud_utils.chamfer_distance (three
  tree builds per direction, OpenMP sweep, numpy norm/mean) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "c3": dict(name="C3: chamfer_distance (+Hausdorff stats) fused, 2x(1e6x3) fp32 uniform", n=1000000, m=1000000, k=1),
    "c2": dict(name="C2: k_nearest_neighbors k=1, 1e6x3 query vs 1e6x3 dataset fp32 uniform", n=1000000, m=1000000, k=1),
    "c4": dict(name="C4: k_nearest_neighbors k=16, 1e7x3 query vs 1e6x3 dataset fp32 uniform", n=10000000, m=1000000, k=16),
    "c5": dict(name="C5: batched chamfer, 1024 pairs of 2x(65536x3) fp32 uniform, pair-sharded", n=65536, m=65536, k=1,
               batch=1024),
}
ALGO_BYTES_PER_QPT = {"c3": 24.0, "c2": 36.0, "c4": 205.2, "c5": 24.0}

# Test hook (tests/test_bench_cpu.py): PCU_BENCH_SCALE=0.01 shrinks every cloud so that the reference arm's
# JSON contract can be checked in seconds.  Never set for a measurement.
_SCALE = float(os.environ.get("PCU_BENCH_SCALE", "1"))
if _SCALE != 1.0:
    for _w in WORKLOADS.values():
        _w["n"] = max(16, int(_w["n"] * _SCALE))
        _w["m"] = max(16, int(_w["m"] * _SCALE))
        if "batch" in _w:
            _w["batch"] = max(2, int(_w["batch"] * _SCALE))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_clouds(seed, n, m, batch=None):
    rng = np.random.default_rng(seed)
    if batch:
        return rng.random((batch, n, 3), dtype=np.float32), rng.random((batch, m, 3), dtype=np.float32)
    return rng.random((n, 3), dtype=np.float32), rng.random((m, 3), dtype=np.float32)


# ------------------------------------------------------------------------------------------------
def run_reference(args, wl_key):
    """The reference's CPU path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    O.build()
    kind = "reference" if O.have_reference() else "port"
    wl = WORKLOADS[wl_key]
    n, m = wl["n"], wl["m"]
    total_steps = args.steps + args.warmup
    # bounded sample: the full C3 pair costs ~3 s per step on the reference path (six serial tree builds)
    if wl_key == "c4":
        n = 200000
    if wl_key == "c5":
        sample_pairs = 2
    else:
        sample_pairs = 1
    if total_steps > 40 and wl_key in ("c3", "c2"):
        n = m = 250000
    x, y = make_clouds(1234, n, m)
    cores = O.hardware_threads(kind)

    def step():
        if wl_key in ("c3", "c5"):
            acc = 0.0
            for _ in range(sample_pairs):
                acc += float(O.chamfer_distance(x, y, impl=kind))
            return acc
        d, i = O.k_nearest_neighbors(x, y, wl["k"], impl=kind)
        return float(d.sum())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    units = ((n + m) if wl_key in ("c3", "c5") else n) * sample_pairs
    value = units * args.steps / dt
    sample = "%d pair(s) of 2x(%dx3) fp32, reference-faithful (3 tree builds per direction, OpenMP sweep, numpy norm/mean)" \
        % (sample_pairs, n) if wl_key in ("c3", "c5") else "%d queries vs %d points, k=%d" % (n, m, wl["k"])
    line = {
        "impl": "reference", "metric": metric_name(wl_key), "value": value, "unit": unit_name(wl_key),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "host": "CPU only (nanoflann kd-tree, %d threads)" % cores},
        "cpu_baseline": {"value": value, "unit": unit_name(wl_key), "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit_name(wl_key), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def metric_name(wl_key):
    return "chamfer_query_points_per_sec" if wl_key in ("c3", "c5") else "knn_queries_per_sec"


def unit_name(wl_key):
    return "query-points/s" if wl_key in ("c3", "c5") else "queries/s"


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl_key = args.workload
    if args.impl == "reference":
        return run_reference(args, wl_key)

    import torch
    import torch.distributed as dist
    import pcu_b200 as pcu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or pcu.device_count() == 0:
        raise SystemExit("bench.py needs a B200: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[wl_key]
    n, m, k = wl["n"], wl["m"], wl["k"]
    batch = wl.get("batch")
    if batch:
        from importlib import import_module
        bmod = import_module("point-cloud-utils_b200._batched")
        lo, hi = bmod.shard_bounds(batch, world, rank)
        local_batch = hi - lo
        xh, yh = make_clouds(1000 + rank, n, m, local_batch)
    else:
        local_batch = 1
        xh, yh = make_clouds(1000 + rank, n, m)
    # pinned host copies (e2e path) and device-resident copies (value path)
    xp = torch.from_numpy(xh).pin_memory()
    yp = torch.from_numpy(yh).pin_memory()
    xd, yd = xp.to(dev, non_blocking=True), yp.to(dev, non_blocking=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def device_step():
        if wl_key == "c3":
            res = pcu.chamfer_distance(xd, yd)                       # 0-dim CUDA tensor, no synchronisation
        elif wl_key == "c5":
            _, res = bmod.batched_chamfer(xd, yd, return_sum=True)   # fp64 sum of this rank's pairs
        else:
            pcu.k_nearest_neighbors(xd, yd, k)
            return None
        if world > 1:   # the path's only exchange: one fp64 scalar per rank
            res = res.to(torch.float64).reshape(1)
            dist.all_reduce(res, op=dist.ReduceOp.SUM)
        return res

    def host_step():
        if wl_key == "c3":
            return float(pcu.chamfer_distance(xp.numpy(), yp.numpy()))
        if wl_key == "c5":
            return float(bmod.batched_chamfer(xp.numpy(), yp.numpy(), return_sum=True)[1])
        d, i = pcu.k_nearest_neighbors(xp.numpy(), yp.numpy(), k)
        return float(d[0].sum())

    units_per_step_rank = ((n + m) if wl_key in ("c3", "c5") else n) * local_batch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm ------------------------------------------------------------------
    for _ in range(args.warmup):
        flush.fill_(1)
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = pcu.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        flush.fill_(s & 0xff)          # evict the 24 MB of inputs / scratch from L2 (outside the interval)
        starts[s].record()
        device_step()
        stops[s].record()
    barrier()
    launches = pcu.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = sum(a.elapsed_time(b) for a, b in zip(starts, stops))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    units = torch.tensor([float(units_per_step_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    units_per_step = float(units.item())
    ms_per_step = total_ms / args.steps
    value = units_per_step / (ms_per_step * 1e-3)

    # ---- roofline pass: per-stage CUDA events inside the library, same workload -----------------
    pcu._pcu_internal._set_profiling(local, stream, True)
    stage_ms = {}
    reps = min(args.steps, 30)
    for s in range(reps):
        flush.fill_(s & 0xff)
        device_step()
        torch.cuda.synchronize(dev)
        for key, val in pcu._pcu_internal._last_profile(local, stream).items():
            stage_ms[key] = stage_ms.get(key, 0.0) + val / reps
    pcu._pcu_internal._set_profiling(local, stream, False)
    peak, peak_src = measured_peak()
    algo_bytes = ALGO_BYTES_PER_QPT[wl_key] * units_per_step_rank
    search_ms = stage_ms.get("search", 0.0)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            traffic = json.load(f).get(wl_key, {}).get("search_bytes_per_launch")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "nn1_kernel (fused search sweep, both directions in one launch)"
        if wl_key in ("c3", "c5") else ("nn1_kernel" if k == 1 else "knn_warp_kernel"),
        "achieved": algo_bytes / (search_ms * 1e-3) / 1e9 if search_ms > 0 else None, "peak": peak, "unit": "GB/s",
        "frac": (algo_bytes / (search_ms * 1e-3) / 1e9 / peak) if search_ms > 0 else None, "traffic": traffic,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": search_ms,
        "stage_ms": {kk: round(v, 5) for kk, v in stage_ms.items()},
        "step_frac": algo_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
    }

    # ---- end-to-end arm: numpy-facing API, pinned host buffers, copies inside the timed region ---
    e2e_steps = args.steps if wl_key in ("c3", "c2") else min(args.steps, 10)
    for _ in range(3):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step / (float(te.item()) / e2e_steps)
    h2d = int(xh.nbytes + yh.nbytes)
    d2h = 2 * 72 + 4 if wl_key == "c3" else (4 * local_batch + 8 if wl_key == "c5" else int(n * k * 12 + 8))

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        kind = "reference" if O.have_reference() else "port"
        cores = O.hardware_threads(kind)
        if wl_key in ("c3", "c5"):
            cn = 1000000 if wl_key == "c3" else 65536
            reps_cpu = 1 if wl_key == "c3" else 8
            cx, cy = make_clouds(4321, cn, cn)
            t0 = time.perf_counter()
            for _ in range(reps_cpu):
                O.chamfer_distance(cx, cy, impl=kind)
            dt = time.perf_counter() - t0
            cpu = {"value": 2 * cn * reps_cpu / dt, "unit": "query-points/s", "cores": cores, "kind": kind,
                   "sample": "%d x chamfer_distance on 2x(%dx3) fp32 exactly as point_cloud_utils does it (three kd-tree "
                             "builds per direction, OpenMP query sweep, numpy gather/norm/mean): %.2f s" % (reps_cpu, cn, dt)}
        else:
            cn = 1000000 if wl_key == "c2" else 300000
            cx, cy = make_clouds(4321, cn, m)
            t0 = time.perf_counter()
            O.k_nearest_neighbors(cx, cy, k, impl=kind)
            dt = time.perf_counter() - t0
            cpu = {"value": cn / dt, "unit": "queries/s", "cores": cores, "kind": kind,
                   "sample": "%d of the queries vs the full %d-point dataset, k=%d: %.2f s" % (cn, m, k, dt)}

    if rank == 0:
        line = {
            "metric": metric_name(wl_key), "value": value, "unit": unit_name(wl_key), "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if wl_key != "c5" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": wl["name"], "pairs_per_gpu": local_batch, "seed": 1000,
                "parallelism": "pair-sharded, one rank per GPU, NCCL all-reduce of one fp64 sum" if world > 1 else "1 GPU",
                "l2": "inputs (24 MB) < L2: a 256 MiB buffer is overwritten before every timed step, outside the interval",
                "timing": "sum of per-step CUDA-event intervals on the launching stream, max over ranks",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": unit_name(wl_key), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(te.item()) / e2e_steps * 1e3, "steps": e2e_steps,
                    "api": "pcu.chamfer_distance(numpy, numpy) -> pcu_b200_chamfer_host_f32 (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
