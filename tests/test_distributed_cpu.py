"""CPU, world_size 2, gloo: the host-side plumbing of the pair-sharded batched Chamfer
(shard ownership, the scalar all-reduce, the per-pair all-gather).  The per-shard compute itself needs
a GPU and is covered by the -m gpu tests; here each rank contributes known per-pair values."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import importlib
import torch, torch.distributed as dist
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
import pcu_b200
b = importlib.import_module("point-cloud-utils_b200._batched")
rank, world = dist.get_rank(), dist.get_world_size()
for batch in (1, 2, 5, 8, 1024):
    lo, hi = b.shard_bounds(batch, world, rank)
    vals = torch.arange(lo, hi, dtype=torch.float32) * 0.5 + 1.0        # pair p is worth 0.5 p + 1
    total = b.reduce_sum(vals.double().sum())
    expect = sum(0.5 * p + 1.0 for p in range(batch))
    assert abs(float(total) - expect) < 1e-9, (batch, float(total), expect)
    allv = b.gather_values(vals, batch)
    assert allv.shape == (batch,), allv.shape
    assert torch.equal(allv, torch.arange(batch, dtype=torch.float32) * 0.5 + 1.0)
try:
    b.gather_values(torch.zeros(4), 5)      # wrong shard length (ranks own 3 and 2) is rejected on every rank
    raise SystemExit("expected ValueError")
except ValueError:
    pass
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_partition():
    import importlib
    sys.path.insert(0, ROOT)
    import pcu_b200  # noqa: F401
    b = importlib.import_module("point-cloud-utils_b200._batched")
    for batch in (1, 2, 7, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            owned = []
            for r in range(world):
                lo, hi = b.shard_bounds(batch, world, r)
                assert 0 <= lo <= hi <= batch
                owned.extend(range(lo, hi))
                assert (hi - lo) in (batch // world, batch // world + 1)
            assert owned == list(range(batch))


def test_two_rank_gloo_plumbing(pcu, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, out)
        assert "rank %d ok" % rank in out
