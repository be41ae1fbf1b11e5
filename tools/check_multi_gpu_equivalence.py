"""W ranks x N/W envs == 1 process x N envs, against vectors recorded from the UNMODIFIED single-process reference.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        tools/check_multi_gpu_equivalence.py [golden case] [cuda_core]

Every rank owns the env shard ``[r N/W, (r+1) N/W)`` of the golden case (default ``ff_cartpole``: N = 64, T = 16,
two rounds, full-batch updates, so the sharded minibatch is the reference's minibatch -- SURVEY.md §8e), injected
with the shard's recorded initial state and sampling noise.  Asserted on every rank: the shard's rollout buffers
equal the recorded ones (discrete actions bit-exact); the GLOBAL statistics (collect stats, reward scale, losses,
KL) and the updated parameters equal the single-process reference's at the golden tolerances; replicas stay
bit-identical.  Started from per-rank random weights on purpose: ``Algorithm.__init__`` broadcasts rank 0's.
"""

from __future__ import annotations

import os
import sys

import pytest
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.conftest import Golden  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "ff_cartpole"
if len(sys.argv) > 2 and sys.argv[2] == "cuda_core":
    os.environ["RL8_FP32_SIMT"] = "1"

import rl8_b200.env as E  # noqa: E402
from rl8_b200 import AlgorithmConfig  # noqa: E402
from rl8_b200 import distributions as Dm  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
DEV = "cuda"

g = Golden(case)
meta = dict(g.meta)
env_cls = {"discrete_dummy": E.DiscreteDummyEnv, "continuous_dummy": E.ContinuousDummyEnv, "cartpole": E.CartPole,
           "mountain_car": E.MountainCar, "pendulum": E.Pendulum}[meta.pop("env")]
dist_cls = {"categorical": Dm.Categorical, "normal": Dm.Normal, "squashed_normal": Dm.SquashedNormal}[meta.pop("dist")]
N, T, rounds = meta.pop("N"), meta.pop("T"), meta.pop("rounds")
sub = meta.pop("subsample")
for k in [k for k in meta if k.startswith("min_relu_input")]:
    meta.pop(k)
assert N % world == 0
assert meta.get("sgd_minibatch_size") is None, "equivalence needs full-batch updates (SURVEY.md §8e)"
n = N // world
sl = slice(rank * n, (rank + 1) * n)
states: list[torch.Tensor] = []
noises: list[torch.Tensor] = []
dummy = "Dummy" in env_cls.__name__


class InjEnv(env_cls):  # type: ignore[misc, valid-type]
    def reset(self, *, config=None):  # noqa: ANN001, ANN202
        obs = super().reset(config=config)
        if not states:
            return obs
        obs = self.set_state(states.pop(0).to(DEV))
        return self.state if dummy else obs


class InjDist(dist_cls):  # type: ignore[misc, valid-type]
    @classmethod
    def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
        if not noises:
            return super().draw_noise(steps, num, width, device)
        z = noises.pop(0).to(device)
        return z.reshape(steps, num, width) if cls.rl8_kind == 0 else z.reshape(steps, num)


InjEnv.__name__ = env_cls.__name__
torch.manual_seed(1234 + rank)  # different initial weights per rank: the constructor must broadcast rank 0's
algo = AlgorithmConfig(num_envs=n, horizon=T, distribution_cls=InjDist, **meta).build(InjEnv)
algo.policy.model.load_state_dict(g.group("param0"))


def close(a: torch.Tensor, b: torch.Tensor, rtol: float, atol: float, msg: str) -> None:
    torch.testing.assert_close(a.detach().cpu().float(), b.float(), rtol=rtol, atol=atol, msg=lambda m: f"{msg}: {m}")


def same_everywhere(t: torch.Tensor) -> bool:
    ts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(ts, t.contiguous())
    return all(torch.equal(ts[0], x) for x in ts)


hper = g.meta["horizons_per_env_reset"]
for rnd in range(rounds):
    p = f"r{rnd}"
    will_reset = (algo.state.horizons % hper) == 0
    if will_reset:
        s0 = g[f"{p}/state0"]
        states.append(s0[sl] if dummy else s0[:, sl])
    noises.append(g[f"{p}/noise"][:, sl])
    cstats = algo.collect()
    ref_actions = g[f"{p}/collect/actions"][sl]
    if ref_actions.dtype == torch.int64:
        assert torch.equal(algo.buffer["actions"].cpu(), ref_actions), "discrete actions of the shard must be bit-exact"
    for k in ("obs", "rewards", "logp", "values", "reversed_discounted_returns"):
        close(algo.buffer[k], g[f"{p}/collect/{k}"][sl], 2e-5, 5e-6, f"rank {rank} {k}")
    for k, v in g.group(f"{p}/collect_stats").items():
        assert cstats[k] == pytest.approx(float(v), rel=1e-5, abs=1e-6), (k, cstats[k], float(v))
    assert algo.state.reward_scale == pytest.approx(g.scalar(f"{p}/reward_scale"), rel=1e-5)
    sstats = algo.step()
    for k, v in g.group(f"{p}/step_stats").items():
        assert sstats[k] == pytest.approx(float(v), rel=2e-5, abs=2e-6), (k, sstats[k], float(v))
    sd = algo.policy.model.state_dict()
    for k, ref in g.group(f"{p}/param1").items():
        close(sd[k].flatten()[::sub], ref, 1e-5, 2e-6, f"param {k}")
    assert same_everywhere(algo.policy.model.flat_params), "replicas diverged"
    if rank == 0:
        print(f"{case} round {rnd}: {world} ranks x {n} envs reproduce the single-process reference "
              f"(losses/total {sstats['losses/total']:.6f}, returns/mean {cstats['returns/mean']:.5f})")
if rank == 0:
    print("EQUIVALENCE OK")
dist.destroy_process_group()
