"""Which host placement feeds this GPU fastest?  For every NUMA node: bind the process to its CPUs, allocate fresh
pinned buffers there, time H2D copies and the numpy-facing Chamfer call on them.  python tools/numa_probe.py"""
import glob, os, re, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pcu_b200 as pcu


def parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


nodes = {}
for path in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    nodes[int(re.search(r"node(\d+)", path).group(1))] = parse_cpulist(open(path).read())
print("NUMA nodes:", {k: "%d cpus (%d..%d)" % (len(v), v[0], v[-1]) if v else "no cpus" for k, v in nodes.items()})
allowed = os.sched_getaffinity(0)
print("allowed cpus:", len(allowed))
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
    aff = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
    print("NVML affinity of GPU 0: %d cpus (%d..%d)" % (len(aff), aff[0], aff[-1]))
    try:
        print("NVML numa node id:", pynvml.nvmlDeviceGetNumaNodeId(h))
    except Exception as e:
        print("no nvmlDeviceGetNumaNodeId:", e)
except Exception as e:
    print("NVML unavailable:", e)
for f in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    pass
print(subprocess.run("nvidia-smi topo -m | head -14", shell=True, capture_output=True, text=True).stdout)

n = 1000000
rng = np.random.default_rng(0)
x = rng.random((n, 3), dtype=np.float32); y = rng.random((n, 3), dtype=np.float32)
xd = torch.empty((n, 3), device="cuda"); yd = torch.empty((n, 3), device="cuda")
keep = []


def h2d_ms(xp, yp, reps=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        a.record(); xd.copy_(xp, non_blocking=True); yd.copy_(yp, non_blocking=True); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def call_ms(a, b, reps=20):
    for _ in range(3):
        float(pcu.chamfer_distance(a, b))
    t0 = time.perf_counter()
    for _ in range(reps):
        float(pcu.chamfer_distance(a, b))
    return (time.perf_counter() - t0) / reps * 1e3


def probe(label):
    xp = torch.empty((n, 3), dtype=torch.float32).pin_memory(); yp = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    xp.copy_(torch.from_numpy(x)); yp.copy_(torch.from_numpy(y))
    keep.extend([xp, yp])        # never hand the blocks back to the caching host allocator
    xs, ys = np.array(x, copy=True), np.array(y, copy=True)
    print("%-28s H2D 24 MB pinned %.3f ms (%.1f GB/s)   chamfer(pinned numpy) %.3f ms   chamfer(pageable) %.3f ms"
          % (label, h2d_ms(xp, yp), 24e-3 / h2d_ms(xp, yp), call_ms(xp.numpy(), yp.numpy()), call_ms(xs, ys)), flush=True)


probe("as started")
for node, cpus in nodes.items():
    cpus = [c for c in cpus if c in allowed]
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    probe("bound to node %d" % node)
os.sched_setaffinity(0, allowed)
probe("all cpus again")
