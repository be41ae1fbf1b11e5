"""CPU, world_size 2 (gloo): the host-side collectives of the env-sharded data-parallel path.

Every rank holds a shard of the envs; after the reductions in rl8_b200/parallel.py the
global statistics must equal the single-process ones, and summed per-rank gradients that
each carry 1 / (M * world) must equal the single-process mean gradient."""

from __future__ import annotations

import math
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from rl8_b200 import parallel

    rk, local, w = parallel.init_from_env("gloo")
    assert (rk, w) == (rank, world) and parallel.world_size() == world and parallel.rank() == rank
    gen = torch.Generator().manual_seed(0)
    N, T = 10, 6
    rewards = torch.randn(T, N, generator=gen, dtype=torch.float64)
    rdr = torch.randn(T, N, generator=gen, dtype=torch.float64) * 3 + 1
    grads_full = torch.randn(4, 33, generator=gen)  # 4 "rows" of per-sample gradients
    b, e = parallel.shard_envs(N, world, rank)
    r, d = rewards[:, b:e], rdr[:, b:e]
    R = r.sum(0)
    acc = torch.zeros(16, dtype=torch.float64)
    acc[0], acc[1], acc[2], acc[3] = r.sum(), (r * r).sum(), R.sum(), (R * R).sum()
    acc[4], acc[5] = d.sum(), (d * d).sum()
    acc[6], acc[7], acc[8], acc[9] = r.min(), r.max(), R.min(), R.max()
    parallel.reduce_collect_acc_(acc)
    a = acc.tolist()
    mean, std = parallel.mean_std(a[0], a[1], float(N * T))
    assert math.isclose(mean, float(rewards.mean()), rel_tol=1e-12)
    assert math.isclose(std, float(rewards.std()), rel_tol=1e-10)
    Rm, Rs = parallel.mean_std(a[2], a[3], float(N))
    assert math.isclose(Rs, float(rewards.sum(0).std()), rel_tol=1e-10)
    assert math.isclose(parallel.mean_std(a[4], a[5], float(N * T))[1], float(rdr.std()), rel_tol=1e-10)
    assert a[6] == float(rewards.min()) and a[7] == float(rewards.max())
    assert a[8] == float(rewards.sum(0).min()) and a[9] == float(rewards.sum(0).max())
    # gradients: each rank averages its rows over the GLOBAL count, the sum is the global mean
    rows = grads_full[rank * 2 : rank * 2 + 2]
    g = rows.sum(0) / 4.0
    parallel.all_reduce_sum_(g)
    torch.testing.assert_close(g, grads_full.mean(0))
    # replicas: ranks seeded differently must end up with rank 0's parameters (and buffers)
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.BatchNorm1d(5))
    net[1].running_mean.normal_()
    moments = torch.randn(7)

    class Flat(torch.nn.Module):  # a fused model: every parameter is a view of ONE flat buffer
        def __init__(self) -> None:
            super().__init__()
            self._flat = torch.randn(12)
            self.w = torch.nn.Parameter(self._flat[:12].view(3, 4))

    flat = Flat()
    before = [t.detach().clone() for t in net.state_dict().values()] + [flat._flat.clone(), moments.clone()]
    parallel.sync_replicas(net, moments)
    parallel.sync_replicas(flat)
    after = [t.detach().clone() for t in net.state_dict().values()] + [flat._flat.clone(), moments.clone()]
    changed: list[bool] = []
    for b4, now in zip(before, after):
        if now.dtype != torch.float32:
            continue
        both = [torch.empty_like(now) for _ in range(world)]
        dist.all_gather(both, now)
        assert all(torch.equal(both[0], x) for x in both), "replicas differ after sync_replicas"
        if rank == 0:
            assert torch.equal(b4, now), "rank 0 must keep its own values"
        else:
            changed.append(not torch.equal(b4, now))
    if rank != 0:  # Linear weight / bias, running_mean, the flat buffer and the moments were seeded per rank
        assert sum(changed) >= 5, "test is vacuous: ranks started identical"
    assert flat.w.data_ptr() == flat._flat.data_ptr()  # still a view of the flat buffer
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_statistics_and_gradients_match_single_process(tmp_path) -> None:  # noqa: ANN001
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_envs_partitions_everything() -> None:
    from rl8_b200 import parallel

    for total, world in ((1 << 20, 8), (10, 3), (7, 8)):
        spans = [parallel.shard_envs(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert parallel.world_size() == 1 and parallel.rank() == 0
