"""Multi-GPU invariants, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py

Env-sharded data parallelism must keep the replicas identical: after a few Trainer.step() iterations on
DIFFERENT env shards, every rank holds bit-identical parameters and reports identical (globally reduced)
statistics -- for the fused default model (fp32 and bf16 paths) and for a user-defined torch model.
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import rl8_b200.env as E  # noqa: E402
from rl8_b200 import AlgorithmConfig, Trainer  # noqa: E402
from rl8_b200.models import GenericModel  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


class Small(GenericModel):
    def __init__(self, observation_spec, action_spec, /) -> None:  # noqa: ANN001
        super().__init__(observation_spec, action_spec)
        self.body = nn.Sequential(nn.Linear(observation_spec.shape[0], 64), nn.Tanh())
        self.pi, self.vf = nn.Linear(64, action_spec.space.n), nn.Linear(64, 1)

    def forward(self, batch):  # noqa: ANN001, ANN201
        self._z = self.body(batch["obs"])
        return {"logits": self.pi(self._z).unsqueeze(1)}

    def value_function(self):  # noqa: ANN201
        return self._z.new_zeros(0) if self._z is None else self.vf(self._z)


def same_everywhere(t: torch.Tensor) -> bool:
    ts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(ts, t.contiguous())
    return all(torch.equal(ts[0], x) for x in ts)


for name, kw in (("fused fp32", {}), ("fused bf16", {"enable_amp": True}), ("user-defined model", {"model_cls": Small})):
    torch.manual_seed(0)  # identical initial weights on every rank
    algo = AlgorithmConfig(num_envs=2048, horizon=16, sgd_minibatch_size=8192, **kw).build(E.CartPole)
    torch.manual_seed(100 + rank)  # different env states and sampling noise per rank
    trainer = Trainer(algo)
    for _ in range(3):
        stats = trainer.step()
    flat = torch.cat([p.detach().flatten().float() for p in algo.policy.model.parameters()])
    keys = ("losses/total", "monitors/kl_div", "returns/mean", "rewards/std")
    st = torch.tensor([stats[k] for k in keys], dtype=torch.float64, device="cuda")
    ok_p, ok_s = same_everywhere(flat), same_everywhere(st)
    if rank == 0:
        print(f"{name}: parameters identical on {world} ranks: {ok_p}; global statistics identical: {ok_s};"
              f" losses/total {stats['losses/total']:.5f}")
    assert ok_p and ok_s, name
dist.destroy_process_group()
