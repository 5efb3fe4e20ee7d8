"""Batched Chamfer over independent (x_b, y_b) pairs, and its multi-GPU form.

Pairs are independent, so the multi-GPU scheme is pure data parallelism: each rank
(one process per GPU, `torch.distributed`) runs the single-GPU batched path on its contiguous shard
of pairs and the only traffic is one all-reduce of a single fp64 (the batch sum) -- or one
all-gather of the per-pair values when the caller wants them all.
"""
import numpy as _np

from . import _pcu_internal


def _torch():
    import importlib
    return importlib.import_module("torch")


def batched_chamfer(x, y, max_points_per_leaf=10, return_sum=False):
    """x: (B, n, 3), y: (B, m, 3) float32, numpy or CUDA tensors -> (B,) Chamfer distances
    (and, with return_sum, their fp64 sum as computed on the device)."""
    if type(x).__module__.startswith("torch"):
        torch = _torch()
        if not isinstance(y, torch.Tensor) or x.dtype != torch.float32 or y.dtype != torch.float32:
            raise ValueError("batched_chamfer_distance: x and y must both be float32 tensors")
        if x.dim() != 3 or y.dim() != 3 or x.shape[2] != 3 or y.shape[2] != 3 or x.shape[0] != y.shape[0]:
            raise ValueError("batched_chamfer_distance: x and y must have shape (B, n, 3) and (B, m, 3)")
        if x.shape[0] == 0 or x.shape[1] == 0 or y.shape[1] == 0:
            raise ValueError("Invalid input set with zero elements: x and y must have shape (B, n, 3) and (B, m, 3)")
        if not x.is_cuda:
            out, total = _pcu_internal._batched_chamfer(x.numpy(), y.numpy(), int(max_points_per_leaf))
            out = torch.from_numpy(_np.asarray(out))
            return (out, total) if return_sum else out
        if x.device != y.device:
            raise ValueError("x and y must live on the same device")
        xs, ys = x.detach().contiguous(), y.detach().contiguous()
        B = xs.shape[0]
        out = torch.empty(B, dtype=torch.float32, device=xs.device)
        total = torch.zeros((), dtype=torch.float64, device=xs.device)
        want_sum = return_sum and B <= 16384
        _pcu_internal._batched_chamfer_device(xs.data_ptr(), ys.data_ptr(), B, xs.shape[1], ys.shape[1],
                                              out.data_ptr(), total.data_ptr() if want_sum else 0,
                                              int(max_points_per_leaf), xs.device.index or 0,
                                              torch.cuda.current_stream(xs.device).cuda_stream)
        if return_sum:
            return out, (total if want_sum else out.double().sum())
        return out
    out, total = _pcu_internal._batched_chamfer(_np.asarray(x), _np.asarray(y), int(max_points_per_leaf))
    return (out, total) if return_sum else out


def shard_bounds(batch, world_size, rank):
    """Contiguous block of pairs owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_sum(local, group=None):
    """All-reduce (sum) of one fp64 scalar per rank: the only exchange of the sharded batched path."""
    torch = _torch()
    import torch.distributed as dist
    if not isinstance(local, torch.Tensor):
        local = torch.tensor(float(local), dtype=torch.float64)
    local = local.to(torch.float64).reshape(1).clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(local, op=dist.ReduceOp.SUM, group=group)
    return local.reshape(())


def gather_values(vals, batch, group=None):
    """All-gather of the per-pair values of every rank's shard (see shard_bounds) into the (batch,) vector."""
    torch = _torch()
    import torch.distributed as dist
    if not isinstance(vals, torch.Tensor):
        vals = torch.from_numpy(_np.asarray(vals))
    if not (dist.is_available() and dist.is_initialized()):
        return vals
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_bounds(batch, world, rank)
    if vals.numel() != hi - lo:
        raise ValueError("rank %d owns pairs [%d, %d) but passed %d values" % (rank, lo, hi, vals.numel()))
    width = (int(batch) + world - 1) // world
    padded = torch.zeros(width, dtype=vals.dtype, device=vals.device)
    padded[: vals.numel()] = vals
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)
    parts = []
    for r in range(world):
        rlo, rhi = shard_bounds(batch, world, r)
        parts.append(gathered[r][: rhi - rlo])
    return torch.cat(parts)


def distributed_batched_chamfer_sum(x_shard, y_shard, group=None, max_points_per_leaf=10):
    """Every rank passes ITS shard of pairs; returns the fp64 sum of the Chamfer distances of all
    pairs of all ranks (a 0-dim tensor on the shard's device).  One all-reduce of one scalar."""
    _, local = batched_chamfer(x_shard, y_shard, max_points_per_leaf, return_sum=True)
    return reduce_sum(local, group)


def distributed_batched_chamfer(x_shard, y_shard, batch, group=None, max_points_per_leaf=10):
    """Every rank passes its shard (see shard_bounds); returns the (batch,) per-pair values on every
    rank.  One all-gather of at most ceil(batch / world) floats per rank."""
    return gather_values(batched_chamfer(x_shard, y_shard, max_points_per_leaf), batch, group)
