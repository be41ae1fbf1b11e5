"""Multi-GPU plumbing: one process per GPU, environments sharded across ranks.

The reference is single-device.  Here every rank owns ``num_envs`` envs, its
slice of the rollout buffer and a replica of the policy; the only exchanges are

* one gradient all-reduce (SUM of the flat fp32 gradient buffer, ~0.5 MB) per optimizer step
  -- gradients are already divided by the GLOBAL minibatch size, so the sum is the mean;
* a handful of scalar all-reduces that keep the global statistics identical to a
  single-process run over all envs: advantage moments (sum, sum of squares, count), the
  reward-scale moments, collect statistics (sums + min/max) and the per-minibatch loss sums.

All helpers are no-ops when ``torch.distributed`` is not initialised, and work on CPU tensors
with the gloo backend (tests/test_cpu_parallel.py) exactly as on CUDA tensors with NCCL.
"""

from __future__ import annotations

import math
import os

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init_from_env(backend: str = "nccl") -> tuple[int, int, int]:
    """Initialise the default process group from torchrun's environment (RANK, LOCAL_RANK,
    WORLD_SIZE, MASTER_ADDR, MASTER_PORT); returns ``(rank, local_rank, world_size)``."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rk, local, world


def shard_envs(total_envs: int, world: int, rk: int) -> tuple[int, int]:
    """Contiguous env range ``[begin, end)`` of rank ``rk`` (SURVEY.md §8e partitioning)."""
    base, extra = divmod(total_envs, world)
    begin = rk * base + min(rk, extra)
    return begin, begin + base + (1 if rk < extra else 0)


def sync_replicas(model: torch.nn.Module, *extra: torch.Tensor) -> None:
    """Make every rank's replica identical to rank 0's (SURVEY.md §8e "identical init (broadcast
    once)"): the flat parameter buffer of a fused model in one broadcast, else every parameter and
    buffer of the module, plus any ``extra`` tensors (optimizer moments).  A no-op on one rank.
    Without it replicas seeded differently would apply the same averaged gradient to different
    weights forever, with no error."""
    if world_size() <= 1:
        return
    flat = getattr(model, "_flat", None)
    with torch.no_grad():
        if flat is not None:
            dist.broadcast(flat, src=0)
        else:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0)
        for t in extra:
            dist.broadcast(t, src=0)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(t)
    return t


def reduce_collect_acc_(acc: torch.Tensor) -> torch.Tensor:
    """Reduce the 16-double accumulator of ``rl8_collect_stats`` over all ranks, in place: slots 0..5 are sums,
    6 / 8 minima, 7 / 9 maxima (include/rl8_b200.h).  ONE collective (an all-gather of the 16 doubles, combined
    locally in rank order, so every rank gets bit-identical results) instead of a SUM and a MIN all-reduce."""
    world = world_size()
    if world > 1:
        gathered = [torch.empty_like(acc) for _ in range(world)]
        dist.all_gather(gathered, acc)
        parts = torch.stack(gathered)
        acc[:6] = parts[:, :6].sum(0)
        acc[6], acc[8] = parts[:, 6].min(), parts[:, 8].min()
        acc[7], acc[9] = parts[:, 7].max(), parts[:, 9].max()
    return acc


def mean_std(s: float, s2: float, n: float) -> tuple[float, float]:
    """Mean and unbiased std from a sum, a sum of squares and a count."""
    mean = s / n
    if n <= 1:
        return mean, float("nan")
    var = (s2 - s * mean) / (n - 1)
    return mean, math.sqrt(max(var, 0.0))
