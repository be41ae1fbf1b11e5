"""Parameter names and shapes of the reference's default models (policy export round trip).
Container-only tool::

    PYTHONPATH=oracle/refshim:/root/reference/src TORCHDYNAMO_DISABLE=1 \
        python tests/golden/generate_state_dict_golden.py

Builds the UNMODIFIED upstream default models (src/rl8/models/_feedforward.py:234-383,
src/rl8/models/_recurrent.py:163-341) for the bundled envs' specs and writes
``tests/golden/state_dict_shapes.json``: ``{case: {parameter name: shape}}``.  A ``state_dict()`` of this
engine's models must carry exactly these entries so that weights trained here load into the reference's
``Policy`` / ``RecurrentPolicy`` with ``load_state_dict`` (and the other way round).
"""

from __future__ import annotations

import json
import os

import torch
from rl8.models import (  # upstream
    DefaultContinuousModel,
    DefaultContinuousRecurrentModel,
    DefaultDiscreteModel,
    DefaultDiscreteRecurrentModel,
)
from torchrl.data import Categorical, Unbounded  # refshim

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # case: (model class, obs dim, action spec)
    "ff_discrete_cartpole": (DefaultDiscreteModel, 5, Categorical(3, shape=torch.Size([1]))),
    "ff_discrete_dummy": (DefaultDiscreteModel, 1, Categorical(2, shape=torch.Size([1]))),
    "ff_continuous_pendulum": (DefaultContinuousModel, 3, Unbounded(shape=torch.Size([1]))),
    "rec_discrete_cartpole": (DefaultDiscreteRecurrentModel, 5, Categorical(3, shape=torch.Size([1]))),
    "rec_continuous_pendulum": (DefaultContinuousRecurrentModel, 3, Unbounded(shape=torch.Size([1]))),
}


def main() -> None:
    out = {}
    for name, (cls, d, act) in CASES.items():
        model = cls(Unbounded(shape=torch.Size([d])), act)
        out[name] = {k: list(v.shape) for k, v in model.state_dict().items()}
    path = os.path.join(HERE, "state_dict_shapes.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
