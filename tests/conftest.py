"""Shared pytest config: the ``gpu`` marker and golden-vector helpers."""

from __future__ import annotations

import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu)")


def pytest_collection_modifyitems(config: pytest.Config, items: list[pytest.Item]) -> None:
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """A golden case: ``g["r0/collect/obs"]`` -> torch tensor; ``g.meta`` -> dict."""

    def __init__(self, name: str) -> None:
        self.name = name
        self._z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
        self.meta = ast.literal_eval(str(self._z["meta"])) if "meta" in self._z.files else {}

    def __contains__(self, key: str) -> bool:
        return key in self._z.files

    def __getitem__(self, key: str) -> torch.Tensor:
        return torch.from_numpy(np.array(self._z[key]))

    def scalar(self, key: str) -> float:
        return float(self._z[key])

    def group(self, prefix: str) -> dict[str, torch.Tensor]:
        prefix = prefix.rstrip("/") + "/"
        return {k[len(prefix):]: self[k] for k in self._z.files if k.startswith(prefix)}


GOLDEN_CASES = [
    "ff_discrete_dummy",
    "ff_continuous_dummy_normal",
    "ff_continuous_dummy_squashed",
    "ff_cartpole",
    "ff_mountain_car",
    "ff_pendulum_squashed",
    "ff_pendulum_normal",
    # round 2: rows crossing the 128 / 256-row tiles of the tensor-core kernels with ragged minibatches,
    # normalize_rewards=False, and an early stop (target_kl_div) that triggers inside the reference
    "ff_cartpole_n320",
    "ff_pendulum_n300_raw_rewards",
    "ff_discrete_dummy_early_stop",
]


@pytest.fixture(params=GOLDEN_CASES)
def golden_case(request: pytest.FixtureRequest) -> Golden:
    return Golden(request.param)


RECURRENT_GOLDEN_CASES = [
    "rec_discrete_dummy",
    "rec_cartpole",
    "rec_pendulum_squashed",
    "rec_continuous_dummy_normal",
]


@pytest.fixture(params=RECURRENT_GOLDEN_CASES)
def recurrent_golden_case(request: pytest.FixtureRequest) -> Golden:
    return Golden(request.param)


@pytest.fixture
def kat() -> Golden:
    return Golden("kat")
