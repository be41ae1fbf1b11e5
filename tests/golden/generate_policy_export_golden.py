"""Policy export round trip: weights + deterministic samples of the UNMODIFIED reference ``Policy``.
Container-only tool::

    PYTHONPATH=oracle/refshim:/root/reference/src TORCHDYNAMO_DISABLE=1 \
        python tests/golden/generate_policy_export_golden.py

For each case it builds the upstream ``rl8.policies.Policy`` (src/rl8/policies/_feedforward.py:20-190) with random
weights, samples it deterministically on fixed observations (``Policy.sample(..., kind="last", deterministic=True)``,
:66-176) and stores the ``state_dict`` with the observations, actions, log-probabilities and values in
``tests/golden/policy_export.npz``.  The GPU test loads the state_dict into this engine's ``Policy`` and must reproduce
the outputs -- i.e. a checkpoint trained on one side deploys on the other (``Policy.save`` / ``load_state_dict``).
"""

from __future__ import annotations

import os

import numpy as np
import torch
from rl8.data import DataKeys  # upstream
from rl8.distributions import SquashedNormal  # upstream
from rl8.policies import Policy  # upstream
from tensordict import TensorDict  # refshim
from torchrl.data import Categorical, Unbounded  # refshim

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # case: (obs dim, action spec, distribution_cls)
    "discrete_cartpole": (5, Categorical(3, shape=torch.Size([1])), None),
    "continuous_pendulum_normal": (3, Unbounded(shape=torch.Size([1])), None),
    "continuous_pendulum_squashed": (3, Unbounded(shape=torch.Size([1])), SquashedNormal),
}
B = 384  # three 128-row tiles


def main() -> None:
    out: dict[str, np.ndarray] = {}
    for seed, (name, (d, act, dist_cls)) in enumerate(CASES.items()):
        torch.manual_seed(100 + seed)
        policy = Policy(Unbounded(shape=torch.Size([d])), act, distribution_cls=dist_cls)
        with torch.no_grad():  # the reference initialises the heads near zero: spread them so argmax / means are decided
            for p in policy.model.parameters():
                if p.abs().max() < 1e-2:
                    p.uniform_(-0.3, 0.3)
        obs = torch.randn(B, 1, d) * 1.5
        batch = TensorDict({DataKeys.OBS: obs}, batch_size=[B, 1])
        sample = policy.sample(batch, kind="last", deterministic=True, return_actions=True, return_logp=True,
                               return_values=True)
        for k, v in policy.model.state_dict().items():
            out[f"{name}/param/{k}"] = v.detach().numpy()
        out[f"{name}/obs"] = obs.numpy()
        out[f"{name}/actions"] = sample[DataKeys.ACTIONS].detach().numpy()
        out[f"{name}/logp"] = sample[DataKeys.LOGP].detach().numpy()
        out[f"{name}/values"] = sample[DataKeys.VALUES].detach().numpy()
        print(name, {k: tuple(out[f"{name}/{k}"].shape) for k in ("obs", "actions", "logp", "values")})
    np.savez_compressed(os.path.join(HERE, "policy_export.npz"), **out)


if __name__ == "__main__":
    main()
