"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over NVLink /
NVSwitch on the B200 box, gloo in CPU tests).

The path shards without a data-path exchange: every configuration / waypoint / IK seed is independent
and the model and cost tables (< 4 KB) are replicated.  Rows are partitioned into contiguous blocks
with whole trajectories kept on one rank, so a [b, h] cost never crosses ranks.  The only collectives
are the ones the callers of the path need:
  * all_reduce(SUM) of the scalar / per-term total cost;
  * all_gather of each rank's k best IK solutions followed by a local top-k;
  * all_reduce(MIN) of an "every seed valid" flag (emulates robot_tree.py:356 without a host sync
    per iteration when called every few dozen iterations).
Payloads are bytes to KB: latency-bound, so they are issued once per batch, never per kernel.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) block of `n_rows` units (trajectories, seeds, configurations) for `rank`;
    the first n_rows % world_size ranks get one extra row."""
    base, extra = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_trajectories(x: torch.Tensor, rank: int | None = None, world_size: int | None = None) -> torch.Tensor:
    """Slice of a [b, h, d] (or [n, d]) batch owned by this rank; whole trajectories stay together."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(x.shape[0], rank, world_size)
    return x[lo:hi]


def total_cost(local_cost: torch.Tensor) -> torch.Tensor:
    """Sum of the cost over every rank's shard (fp64 accumulation of the per-rank fp32 sums)."""
    s = local_cost.double().sum().reshape(1)
    if world()[1] > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return s[0]


def total_cost_from_partials(block_sums: torch.Tensor) -> torch.Tensor:
    """The same from the per-block partial sums of DeviceCost.objective_sum (the kernel did the first stage of the reduction):
    a fixed-order fp64 sum over <= 8192 values + the all-reduce of the scalar."""
    s = block_sums.sum(dtype=torch.float64).reshape(1)
    if world()[1] > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return s[0]


def total_cost_terms(local_terms: torch.Tensor) -> torch.Tensor:
    """Per-term totals [4] (self, object, ws, ee) over all ranks."""
    s = local_terms.double().reshape(-1, local_terms.shape[-1]).sum(0)
    if world()[1] > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return s


def all_valid(local_all_valid: bool, device) -> bool:
    flag = torch.tensor([1 if local_all_valid else 0], dtype=torch.int32, device=device)
    if world()[1] > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def first_all_valid_iteration(valid_count: torch.Tensor, local_batch: int) -> int:
    """The iteration at which the reference's IK loop would stop on the WHOLE (sharded) batch - the first one at which every
    seed of every rank is valid (robot_tree.py:356) - or -1.  valid_count [n_iters]: this rank's per-iteration count of valid
    seeds (DeviceModel.ik_adam(...)["valid_count"]); one all-reduce(SUM) of the count vector and of the batch sizes replaces a
    flag exchange per iteration."""
    v = torch.cat([valid_count.to(torch.int64).reshape(-1), torch.tensor([int(local_batch)], dtype=torch.int64, device=valid_count.device)])
    if world()[1] > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    hit = torch.nonzero(v[:-1] == v[-1])
    return int(hit[0]) if hit.numel() else -1


def gather_best_solutions(q_local: torch.Tensor, err_local: torch.Tensor, valid_local: torch.Tensor, k: int, local_keys=None):
    """Global k best IK solutions.

    q_local [S, D], err_local [S] (SE(3) error), valid_local [S] bool on this rank.  Each rank selects its
    k best valid seeds (invalid ones are ranked last), the [k, D+2] blocks are all-gathered and every rank
    takes the global top-k.  Returns (q [k, D], err [k], valid [k] bool), identical on all ranks and equal
    to a single-process top-k over the concatenated seeds (ties broken by global seed order).
    local_keys: the per-block candidate keys of DeviceModel.ik_adam(best_k >= k) (error bits << 32 | seed index, int64): the
    local selection then runs over blocks * best_k candidates instead of all S seeds."""
    rank, w = world()
    S, D = q_local.shape
    kk = min(k, S)
    inf = torch.full((), float("inf"), dtype=err_local.dtype, device=err_local.device)
    if local_keys is not None:
        order = torch.topk(local_keys, kk, largest=False, sorted=True).values & 0xFFFFFFFF
    else:
        order = _smallest_k(torch.where(valid_local, err_local, inf), kk)
    # [q | err | valid] rows of the selected seeds (a dozen tiny launches matter when the whole solve takes 2 ms)
    block = torch.cat([q_local.index_select(0, order), err_local.index_select(0, order)[:, None],
                       valid_local.index_select(0, order).to(q_local.dtype)[:, None]], dim=1)
    if kk < k:
        pad = torch.full((k - kk, D + 2), float("inf"), dtype=q_local.dtype, device=q_local.device)
        pad[:, D + 1] = 0
        block = torch.cat([block, pad], 0)
    if w > 1:
        allb = torch.empty((w * k, D + 2), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(allb, block.contiguous())
    else:
        allb = block
    top = _smallest_k(torch.where(allb[:, D + 1] > 0, allb[:, D], inf), k)
    sel = allb.index_select(0, top)
    return sel[:, :D], sel[:, D], sel[:, D + 1] > 0


def _smallest_k(key: torch.Tensor, k: int) -> torch.Tensor:
    """Indices of the k smallest entries of a non-negative fp32 key, ascending, ties by index (what a stable argsort gives) -
    as one top-k over 64-bit integers (key bits << 32 | index: for non-negative floats the bit pattern orders like the value,
    +inf last) instead of a full sort of the batch."""
    n = key.shape[0]
    if key.dtype != torch.float32 or n >= (1 << 31):
        return torch.argsort(key, stable=True)[:k]
    # (an SE(3) error can round to a tiny negative value: its bit pattern is a negative integer and still ranks first; no host
    # synchronisation is spent on checking)
    packed = (key.contiguous().view(torch.int32).to(torch.int64) << 32) | torch.arange(n, device=key.device, dtype=torch.int64)
    vals = torch.topk(packed, k, largest=False, sorted=True).values
    return vals & 0xFFFFFFFF


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index`, so that page-locked
    host buffers allocated afterwards (first touch) and the copy threads sit on the GPU's NUMA node.  Matters for
    the host-buffer path when several ranks share one host: without it, ranks contend for one socket's memory
    and the inter-socket link.  Returns False when NVML is unavailable (the call is then a no-op)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(int(device_index)))
        return True
    except Exception:
        return False
