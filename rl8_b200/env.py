"""Tensor-batched environment protocol and the bundled environments.

``Env`` keeps the reference's protocol (src/rl8/env.py:16-128): ``env_cls(num_envs, horizon,
*, device)``, ``reset(*, config) -> obs [N, ...]``, ``step(action) -> {"obs", "rewards"}``.
The bundled environments (dummy envs from src/rl8/env.py:154-259 and CartPole / MountainCar
/ Pendulum from the reference's examples/) run on hand-written sm_100a kernels
(``rl8_env_reset`` / ``rl8_env_step``): state is a struct-of-arrays ``[S, N]`` tensor, one
env per lane, 128-bit coalesced loads and stores.  They also carry ``rl8_kind`` /
``rl8_cfg()`` so :class:`rl8_b200.Algorithm` can fuse them into its rollout kernel.
"""

from __future__ import annotations

import math
from abc import ABC, abstractmethod
from dataclasses import asdict, dataclass
from typing import Any, ClassVar, Mapping

import torch

from . import _lib
from .data import DataKeys, Device
from .specs import Categorical, TensorSpec, Unbounded


class Env(ABC):
    """Protocol of a highly parallel (IsaacGym-like) environment."""

    action_spec: TensorSpec
    observation_spec: TensorSpec
    device: Device
    horizon: None | int
    max_horizon: ClassVar[int]
    max_num_envs: ClassVar[int]
    num_envs: int

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        if hasattr(self, "max_horizon") and horizon is not None and horizon > self.max_horizon:
            raise ValueError(f"{type(self).__name__} `horizon` must be <= {self.max_horizon}.")
        if hasattr(self, "max_num_envs") and num_envs > self.max_num_envs:
            raise ValueError(f"{type(self).__name__} `num_envs` must be <= {self.max_num_envs}.")
        self.num_envs = num_envs
        self.horizon = horizon
        self.device = device

    @abstractmethod
    def reset(self, *, config: None | dict[str, Any] = None) -> torch.Tensor:
        """Reset every environment; returns the initial observation ``[N, ...]``."""

    @abstractmethod
    def step(self, action: torch.Tensor) -> Mapping[str, torch.Tensor]:
        """Apply ``action [N, ...]``; returns ``{"obs": [N, ...], "rewards": [N, 1]}``."""


class KernelEnv(Env):
    """Base of the environments implemented as CUDA kernels.

    Subclasses define ``rl8_kind``, the SoA state height ``S``, the observation width
    ``D`` and ``rl8_cfg()``; the noise their reset consumes is drawn with torch's device
    generator (standard normal or U[0, 1), see ``reset_noise``).
    """

    rl8_kind: ClassVar[int]
    S: ClassVar[int]
    D: ClassVar[int]
    reset_noise: ClassVar[str]  # "normal" | "uniform"

    #: SoA state ``[S, N]`` (f32, contiguous), advanced in place by ``step``.
    state: torch.Tensor

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        super().__init__(num_envs, horizon, device=device)
        if torch.device(device).type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__} runs on CUDA only (device={device!r}); rl8_b200 has no"
                " CPU path."
            )
        self._lib = _lib.load()
        n = num_envs
        self.state = torch.zeros(self.S, n, device=device)
        # obs [D, N] SoA; exposed as the [N, D] transposed view the reference returns
        # (examples/cartpole/env.py:56-59).
        self._obs = self.state if self._obs_aliases_state() else torch.zeros(self.D, n, device=device)
        self._reward = torch.zeros(n, 1, device=device)

    def _obs_aliases_state(self) -> bool:
        return False

    def rl8_cfg(self) -> _lib.EnvCfg:
        raise NotImplementedError

    def draw_reset_noise(self) -> torch.Tensor:
        """Noise ``[S, N]`` the reset kernel turns into the initial state."""
        if self.reset_noise == "normal":
            return torch.randn(self.S, self.num_envs, device=self.device)
        return torch.rand(self.S, self.num_envs, device=self.device)

    def _reset_kernel(self) -> torch.Tensor:
        noise = self.draw_reset_noise()
        cfg = self.rl8_cfg()
        rc = self._lib.rl8_env_reset(
            self.rl8_kind, cfg, _lib.ptr(noise), _lib.ptr(self.state), _lib.ptr(self._obs),
            1, self.num_envs, self.num_envs, _lib.stream(),
        )
        _lib.check(rc, "rl8_env_reset")
        return self._obs.T

    def set_state(self, state: torch.Tensor) -> torch.Tensor:
        """Install ``state`` (``[S, N]``, or ``[N, 1]`` for the dummy envs) and return its
        observation -- e.g. to replay a recorded initial condition."""
        self.state.copy_(state.reshape(self.state.shape))
        rc = self._lib.rl8_env_observe(
            self.rl8_kind, _lib.ptr(self.state), _lib.ptr(self._obs), 1, self.num_envs,
            self.num_envs, _lib.stream(),
        )
        _lib.check(rc, "rl8_env_observe")
        return self._obs.T

    def step(self, action: torch.Tensor) -> dict[str, torch.Tensor]:
        dtype = torch.int64 if isinstance(self.action_spec, Categorical) else torch.float32
        a = action.reshape(-1)
        if a.dtype != dtype or not a.is_contiguous() or a.numel() != self.num_envs:
            if a.numel() != self.num_envs:
                raise ValueError(f"expected {self.num_envs} actions, got {a.numel()}")
            a = a.to(dtype).contiguous()
        _lib.require_cuda(a, "action")
        cfg = self.rl8_cfg()
        rc = self._lib.rl8_env_step(
            self.rl8_kind, cfg, _lib.ptr(self.state), _lib.ptr(a), _lib.ptr(self._obs),
            1, self.num_envs, _lib.ptr(self._reward), self.num_envs, _lib.stream(),
        )
        _lib.check(rc, "rl8_env_step")
        return {DataKeys.OBS: self._obs.T, DataKeys.REWARDS: self._reward}


def _cfg(*values: float) -> _lib.EnvCfg:
    cfg = _lib.EnvCfg()
    for i, v in enumerate(values):
        cfg.p[i] = v  # one rounding double -> f32, like a torch scalar operand
    return cfg


class DummyEnv(KernelEnv):
    """1-D position driven towards the origin (src/rl8/env.py:154-203)."""

    S = 1
    D = 1
    reset_noise = "uniform"

    #: Initial states are drawn from U(-bounds, bounds).
    bounds: float

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        super().__init__(num_envs, horizon, device=device)
        self.observation_spec = Unbounded(1, device=device)
        self.bounds = 100.0
        # The reference keeps the state as [N, 1] and returns it aliased as the observation.
        self.state = self.state.view(num_envs, 1)
        self._obs = self.state.view(1, num_envs)

    def _obs_aliases_state(self) -> bool:
        return True

    def rl8_cfg(self) -> _lib.EnvCfg:
        return _cfg(self.bounds)

    def reset(self, *, config: None | dict[str, Any] = None) -> torch.Tensor:
        config = config or {}
        self.bounds = config.get("bounds", self.bounds)
        self._reset_kernel()
        return self.state

    def step(self, action: torch.Tensor) -> dict[str, torch.Tensor]:
        out = super().step(action)
        out[DataKeys.OBS] = self.state
        return out


class ContinuousDummyEnv(DummyEnv):
    """``state += action`` (src/rl8/env.py:206-230)."""

    rl8_kind = _lib.ENV_CONTINUOUS_DUMMY

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        super().__init__(num_envs, horizon, device=device)
        self.action_spec = Unbounded(shape=torch.Size([1]), device=device)


class DiscreteDummyEnv(DummyEnv):
    """``state += 2 * action - 1`` with two actions (src/rl8/env.py:233-259)."""

    rl8_kind = _lib.ENV_DISCRETE_DUMMY

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        super().__init__(num_envs, horizon, device=device)
        self.action_spec = Categorical(2, shape=torch.Size([1]), device=device)


# ---------------------------------------------------------------------------------------
# Classic-control environments (the reference's examples/)
# ---------------------------------------------------------------------------------------


@dataclass
class CartPoleConfig:
    """examples/cartpole/env.py:67-98 (derived fields are recomputed)."""

    cart_mass: float = 1.0
    force_mag: float = 5.0
    gravity: float = 9.8
    kinematics_integrator: str = "euler"
    length: float = 0.5
    pole_mass: float = 0.1
    pole_mass_length: float = 0.05
    total_mass: float = 1.1
    tau: float = 0.02

    def __post_init__(self) -> None:
        self.pole_mass_length = self.pole_mass * self.length
        self.total_mass = self.cart_mass + self.pole_mass


class CartPole(KernelEnv):
    """Cart-pole swing/balance with a dense reward (examples/cartpole/env.py:101-150)."""

    rl8_kind = _lib.ENV_CARTPOLE
    S = 4
    D = 5
    reset_noise = "normal"
    max_horizon = 128

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        super().__init__(num_envs, horizon, device=device)
        self.observation_spec = Unbounded(5, device=device, dtype=torch.float32)
        self.action_spec = Categorical(3, shape=torch.Size([1]), device=device)
        self._config = CartPoleConfig()

    def rl8_cfg(self) -> _lib.EnvCfg:
        c = self._config
        return _cfg(
            c.force_mag, c.gravity, c.length, c.pole_mass, c.pole_mass_length, c.total_mass,
            c.tau, 4.0 / 3.0, 0.0 if c.kinematics_integrator == "euler" else 1.0,
        )

    def reset(self, *, config: None | dict[str, Any] = None) -> torch.Tensor:
        self._config = CartPoleConfig(**(config or {}))
        return self._reset_kernel()


@dataclass
class MountainCarConfig:
    """examples/mountain_car/env.py:41-62."""

    force_mag: float = 0.001
    goal_position: float = 0.5
    goal_velocity: float = 0.0
    gravity: float = 0.0025
    max_position: float = 0.6
    max_speed: float = 0.07
    min_position: float = -1.2


class MountainCar(KernelEnv):
    """Mountain car with a dense distance reward (examples/mountain_car/env.py:65-116)."""

    rl8_kind = _lib.ENV_MOUNTAIN_CAR
    S = 2
    D = 2
    reset_noise = "normal"
    max_horizon = 512

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        super().__init__(num_envs, horizon, device=device)
        self.observation_spec = Unbounded(2, device=device, dtype=torch.float32)
        self.action_spec = Categorical(3, shape=torch.Size([1]), device=device)
        self._config = MountainCarConfig()
        self._obs = self.state  # obs == state (examples/mountain_car/env.py:36-37)

    def rl8_cfg(self) -> _lib.EnvCfg:
        c = self._config
        return _cfg(
            c.force_mag, c.goal_position, c.goal_velocity, c.gravity, c.max_position,
            c.max_speed, c.min_position,
        )

    def reset(self, *, config: None | dict[str, Any] = None) -> torch.Tensor:
        self._config = MountainCarConfig(**(config or {}))
        return self._reset_kernel()


@dataclass
class PendulumConfig:
    """examples/pendulum/env.py:42-60."""

    dt: float = 0.05
    g: float = 10.0
    l: float = 1.0  # noqa: E741
    m: float = 1.0
    max_speed: float = 8.0
    max_torque: float = 2.0


class Pendulum(KernelEnv):
    """Pendulum swing-up (examples/pendulum/env.py:63-118)."""

    rl8_kind = _lib.ENV_PENDULUM
    S = 2
    D = 3
    reset_noise = "uniform"
    max_horizon = 512

    def __init__(self, num_envs: int, /, horizon: None | int = None, *, device: Device = "cpu"):
        super().__init__(num_envs, horizon, device=device)
        self.observation_spec = Unbounded(3, device=device, dtype=torch.float32)
        self.action_spec = Unbounded(shape=torch.Size([1]), device=device, dtype=torch.float32)
        self._config = PendulumConfig()

    def rl8_cfg(self) -> _lib.EnvCfg:
        c = self._config
        return _cfg(
            c.dt, 3 * c.g / (2 * c.l), 3.0 / (c.m * c.l**2), c.max_speed, c.max_torque,
            math.pi, 2 * math.pi,
        )

    def reset(self, *, config: None | dict[str, Any] = None) -> torch.Tensor:
        self._config = PendulumConfig(**(config or {}))
        return self._reset_kernel()


__all__ = [
    "Env", "KernelEnv", "DummyEnv", "ContinuousDummyEnv", "DiscreteDummyEnv", "CartPole",
    "CartPoleConfig", "MountainCar", "MountainCarConfig", "Pendulum", "PendulumConfig", "asdict",
]
