"""Does the H2D rate ramp up with traffic (PCIe link power management)?  python tools/pcie_ramp.py"""
import subprocess, time
import torch
def link():
    return subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pstate", "--format=csv,noheader"],
                          capture_output=True, text=True).stdout.strip()
x = torch.empty(24 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(24 << 20, dtype=torch.uint8, device="cuda")
a = torch.rand((4096, 4096), device="cuda")
print("link at start:", link(), flush=True)
for phase in ("cold", "after 0.5 s of compute only", "after 0.5 s idle"):
    if phase.startswith("after 0.5 s of compute"):
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.5:
            (a @ a).sum().item()
    if phase.startswith("after 0.5 s idle"):
        time.sleep(0.5)
    ts = []
    for i in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); d.copy_(x, non_blocking=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-28s 24 MB H2D ms: first %s ... copies 20-29 %s ... last %s | link %s" %
          (phase, ["%.2f" % t for t in ts[:6]], ["%.2f" % t for t in ts[20:24]], ["%.2f" % t for t in ts[-3:]], link()), flush=True)
