"""CPU oracle for the rl8 PPO rollout-and-update hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-tensor, single-threaded-semantics restatement (torch fp32 on the
CPU, no TensorDict, no classes from the product) of what the upstream reference
computes on the path SURVEY.md §8 scopes.  It is *the checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``rl8_b200/`` may import, call or link it; the
product fails loudly when its CUDA library is missing instead of routing here.

Pinning: ``tests/golden/generate_golden.py`` runs the *unmodified* upstream source
(``/root/reference/src/rl8`` + ``/root/reference/examples``) behind ``oracle/refshim``
in the build container, with noise injected through the reference's own
``distribution_cls`` plug-in point, and stores inputs + outputs under
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function here
against those vectors and against the reference's own known-answer test
(``tests/test_nn/test_functional.py:14-49``).  Parity status: **pinned**.

All citations are ``path:line`` relative to the upstream repository root.

Conventions
-----------
* ``N`` envs, ``T`` horizon, buffers are env-major ``[N, T+1, ...]`` like the reference.
* Model parameters are a ``dict[str, Tensor]`` keyed like the reference models'
  ``state_dict()`` (``feature_model.0.0.weight`` ...).
* Injected noise: categorical ``q ~ Exp(1)`` of shape ``[N, 1, A]`` per step; normal
  ``z ~ N(0, 1)`` of shape ``[N, 1]`` per step (SURVEY.md §8c "RNG injection").
"""

from __future__ import annotations

import math
from typing import Any, Callable

import torch
import torch.nn.functional as F

Params = dict[str, torch.Tensor]

# --------------------------------------------------------------------------------------
# Environments
# --------------------------------------------------------------------------------------


def dummy_discrete_step(state: torch.Tensor, action: torch.Tensor):
    """src/rl8/env.py:253-259 — ``state += 2a - 1`` (in place), obs aliases state."""
    state += 2 * action - 1
    return state, state, -state.abs()


def dummy_continuous_step(state: torch.Tensor, action: torch.Tensor):
    """src/rl8/env.py:224-230 — ``state += a`` (in place), reward ``-|state|``."""
    state += action
    return state, state, -state.abs()


CARTPOLE_DEFAULTS = dict(
    cart_mass=1.0,
    force_mag=5.0,
    gravity=9.8,
    kinematics_integrator="euler",
    length=0.5,
    pole_mass=0.1,
    tau=0.02,
)


def cartpole_config(**over: Any) -> dict[str, Any]:
    """examples/cartpole/env.py:67-98 — derived fields are recomputed in __post_init__."""
    cfg = {**CARTPOLE_DEFAULTS, **over}
    cfg["pole_mass_length"] = cfg["pole_mass"] * cfg["length"]
    cfg["total_mass"] = cfg["cart_mass"] + cfg["pole_mass"]
    return cfg


def cartpole_obs(state: torch.Tensor) -> torch.Tensor:
    """examples/cartpole/env.py:133-136 — obs ``[N, 5]`` = (x, x', cos th, sin th, th')."""
    x, xd, th, thd = state
    return torch.vstack((x, xd, torch.cos(th), torch.sin(th), thd)).T


def cartpole_step(state: torch.Tensor, action: torch.Tensor, cfg: None | dict = None):
    """examples/cartpole/env.py:12-64 (eager; ``TORCHDYNAMO_DISABLE=1``).

    ``state`` is SoA ``[4, N]``; ``action`` int64 ``[N, 1]`` in {0,1,2}.
    Returns ``(new_state [4,N], obs [N,5] (strided), reward [N,1])``.
    """
    c = cfg or cartpole_config()
    x, xd, th, thd = state
    push = (action.flatten() - 1) * c["force_mag"]
    cth = torch.cos(th)
    sth = torch.sin(th)
    tmp = (push + c["pole_mass_length"] * thd**2 * sth) / c["total_mass"]
    th_acc = (c["gravity"] * sth - cth * tmp) / (
        c["length"] * (4.0 / 3.0 - c["pole_mass"] * cth**2 / c["total_mass"])
    )
    x_acc = tmp - c["pole_mass_length"] * th_acc * cth / c["total_mass"]
    tau = c["tau"]
    if c["kinematics_integrator"] == "euler":
        x = x + tau * xd
        xd = xd + tau * x_acc
        th = th + tau * thd
        thd = thd + tau * th_acc
    else:
        xd = xd + tau * x_acc
        x = x + tau * xd
        thd = thd + tau * th_acc
        th = th + tau * thd
    new_state = torch.vstack((x, xd, th, thd))
    cos_new = torch.cos(th)
    sin_new = torch.sin(th)
    obs = torch.vstack((x, xd, cos_new, sin_new, thd))
    # reward: -( |cos-1| + |sin| ) - ( |x| + |x'| + |th'| ), summed in that order
    ang = torch.vstack((cos_new - 1.0, sin_new - 0.0)).abs().sum(dim=0, keepdim=True).T
    oth = torch.vstack((x, xd, thd)).abs().sum(dim=0, keepdim=True).T
    return new_state, obs.T, -(ang + oth)


MOUNTAIN_CAR_DEFAULTS = dict(
    force_mag=0.001,
    goal_position=0.5,
    goal_velocity=0.0,
    gravity=0.0025,
    max_position=0.6,
    max_speed=0.07,
    min_position=-1.2,
)


def mountain_car_step(state: torch.Tensor, action: torch.Tensor, cfg: None | dict = None):
    """examples/mountain_car/env.py:12-38.  ``state`` ``[2, N]`` is updated in place
    (the reference mutates the position/velocity views) and re-stacked."""
    c = {**MOUNTAIN_CAR_DEFAULTS, **(cfg or {})}
    p, v = state
    v += (action.flatten() - 1) * c["force_mag"] - c["gravity"] * torch.cos(3 * p)
    v = v.clip_(-c["max_speed"], c["max_speed"])
    p += v
    p = p.clip_(c["min_position"], c["max_position"])
    v[(p == c["min_position"]) & (v < 0)] = 0.0
    r = (p - c["goal_position"]).abs_()
    r *= -1
    r[(p >= c["goal_position"]) & (v >= c["goal_velocity"])] = 1.0
    new_state = torch.vstack((p, v))
    return new_state, new_state.T, r.reshape(-1, 1)


PENDULUM_DEFAULTS = dict(dt=0.05, g=10.0, l=1.0, m=1.0, max_speed=8.0, max_torque=2.0)


def pendulum_obs(state: torch.Tensor) -> torch.Tensor:
    """examples/pendulum/env.py:104-106."""
    th, thd = state
    return torch.vstack((torch.cos(th), torch.sin(th), thd)).T


def pendulum_step(state: torch.Tensor, action: torch.Tensor, cfg: None | dict = None):
    """examples/pendulum/env.py:12-39.  Cost is from the *old* state; the angle is only
    wrapped inside the cost (Python-sign remainder)."""
    c = {**PENDULUM_DEFAULTS, **(cfg or {})}
    th, thd = state
    u = torch.clip(action.flatten(), -c["max_torque"], c["max_torque"])
    cost = (
        (((th + torch.pi) % (2 * torch.pi)) - torch.pi) ** 2
        + 0.1 * thd**2
        + 0.001 * (u**2)
    )
    g, l, m, dt = c["g"], c["l"], c["m"], c["dt"]
    new_thd = thd + (3 * g / (2 * l) * torch.sin(th) + 3.0 / (m * l**2) * u) * dt
    new_thd = new_thd.clip_(-c["max_speed"], c["max_speed"])
    new_th = th + new_thd * dt
    new_state = torch.vstack((new_th, new_thd))
    obs = torch.vstack((torch.cos(new_th), torch.sin(new_th), new_thd)).T
    return new_state, obs, (-cost).reshape(-1, 1)


class OracleEnv:
    """Tiny adapter giving every oracle env the same reset/step surface.

    ``reset_fn(num_envs, generator) -> state`` draws from the env's reset distribution
    with torch's CPU generator exactly like the reference env classes do
    (src/rl8/env.py:197-203, examples/*/env.py reset methods).
    """

    SPECS = {
        # name: (obs_dim, action kind, action cardinality, state rows (SoA) or 0 for [N,1])
        "discrete_dummy": (1, "discrete", 2),
        "continuous_dummy": (1, "continuous", 1),
        "cartpole": (5, "discrete", 3),
        "mountain_car": (2, "discrete", 3),
        "pendulum": (3, "continuous", 1),
    }

    def __init__(self, name: str, num_envs: int, config: None | dict = None) -> None:
        self.name = name
        self.num_envs = num_envs
        self.obs_dim, self.action_kind, self.num_actions = self.SPECS[name]
        self.config = config
        self.state: torch.Tensor = torch.empty(0)
        if name == "cartpole":
            self.config = cartpole_config(**(config or {}))

    def reset(self, state: None | torch.Tensor = None) -> torch.Tensor:
        N = self.num_envs
        if state is not None:
            self.state = state.clone()
        elif self.name in ("discrete_dummy", "continuous_dummy"):
            bounds = (self.config or {}).get("bounds", 100.0)
            self.state = torch.empty(N, 1).uniform_(-bounds, bounds)
        elif self.name == "cartpole":
            self.state = torch.normal(0, 0.01, size=(4, N), dtype=torch.float32)
        elif self.name == "mountain_car":
            p = torch.normal(-0.5, 0.05, size=(1, N), dtype=torch.float32)
            v = torch.normal(0, 0.05, size=(1, N), dtype=torch.float32)
            self.state = torch.vstack((p, v))
        elif self.name == "pendulum":
            th = torch.empty(1, N).uniform_(-torch.pi, torch.pi)
            thd = torch.empty(1, N).uniform_(-1.0, 1.0)
            self.state = torch.vstack((th, thd))
        return self.obs()

    def obs(self) -> torch.Tensor:
        if self.name in ("discrete_dummy", "continuous_dummy"):
            return self.state
        if self.name == "cartpole":
            return cartpole_obs(self.state)
        if self.name == "mountain_car":
            return self.state.T
        return pendulum_obs(self.state)

    def step(self, action: torch.Tensor):
        if self.name == "discrete_dummy":
            self.state, obs, r = dummy_discrete_step(self.state, action)
        elif self.name == "continuous_dummy":
            self.state, obs, r = dummy_continuous_step(self.state, action)
        elif self.name == "cartpole":
            self.state, obs, r = cartpole_step(self.state, action, self.config)
        elif self.name == "mountain_car":
            self.state, obs, r = mountain_car_step(self.state, action, self.config)
        else:
            self.state, obs, r = pendulum_step(self.state, action, self.config)
        return obs, r


# --------------------------------------------------------------------------------------
# Default models (two independent MLPs) — src/rl8/models/_feedforward.py:234-383
# --------------------------------------------------------------------------------------


def init_params(obs_dim: int, action_kind: str, num_actions: int, hidden: int = 256) -> Params:
    """Default-model parameters with the reference's initialisation, drawn from torch's
    global CPU generator in the reference's construction order
    (src/rl8/models/_feedforward.py:252-290 continuous, 325-363 discrete;
    ``nn.Linear`` default Kaiming-uniform; heads U(+-1e-3) with zero bias)."""

    def linear(out_f: int, in_f: int) -> tuple[torch.Tensor, torch.Tensor]:
        lin = torch.nn.Linear(in_f, out_f)
        return lin.weight.detach().clone(), lin.bias.detach().clone()

    def small_head(out_f: int, in_f: int) -> tuple[torch.Tensor, torch.Tensor]:
        w, b = linear(out_f, in_f)
        torch.nn.init.uniform_(w, a=-1e-3, b=1e-3)
        torch.nn.init.zeros_(b)
        return w, b

    p: Params = {}
    trunk = "feature_model" if action_kind == "discrete" else "latent_model"
    p[f"{trunk}.0.0.weight"], p[f"{trunk}.0.0.bias"] = linear(hidden, obs_dim)
    p[f"{trunk}.0.2.weight"], p[f"{trunk}.0.2.bias"] = linear(hidden, hidden)
    if action_kind == "discrete":
        p["feature_model.2.weight"], p["feature_model.2.bias"] = small_head(num_actions, hidden)
    else:
        p["action_mean.weight"], p["action_mean.bias"] = small_head(num_actions, hidden)
        p["action_log_std.weight"], p["action_log_std.bias"] = small_head(num_actions, hidden)
    p["vf_model.0.0.weight"], p["vf_model.0.0.bias"] = linear(hidden, obs_dim)
    p["vf_model.0.2.weight"], p["vf_model.0.2.bias"] = linear(hidden, hidden)
    p["vf_model.2.weight"], p["vf_model.2.bias"] = linear(1, hidden)
    return p


def _trunk(p: Params, prefix: str, obs: torch.Tensor) -> torch.Tensor:
    h = F.relu(F.linear(obs, p[f"{prefix}.0.0.weight"], p[f"{prefix}.0.0.bias"]))
    return F.relu(F.linear(h, p[f"{prefix}.0.2.weight"], p[f"{prefix}.0.2.bias"]))


def model_forward(p: Params, obs: torch.Tensor) -> tuple[dict[str, torch.Tensor], torch.Tensor]:
    """``obs [B, D]`` -> (features, values ``[B, 1]``).

    Discrete (src/rl8/models/_feedforward.py:365-375): ``{"logits": [B, 1, A]}``.
    Continuous (292-302): ``{"mean": [B, 1], "log_std": tanh(.) [B, 1]}``.
    """
    value = F.linear(
        _trunk(p, "vf_model", obs), p["vf_model.2.weight"], p["vf_model.2.bias"]
    )
    if "feature_model.2.weight" in p:
        z = _trunk(p, "feature_model", obs)
        logits = F.linear(z, p["feature_model.2.weight"], p["feature_model.2.bias"])
        return {"logits": logits.reshape(-1, 1, logits.shape[-1])}, value
    z = _trunk(p, "latent_model", obs)
    mean = F.linear(z, p["action_mean.weight"], p["action_mean.bias"])
    raw = F.linear(z, p["action_log_std.weight"], p["action_log_std.bias"])
    return {"mean": mean, "log_std": torch.tanh(raw)}, value


# --------------------------------------------------------------------------------------
# Distributions — src/rl8/distributions.py:98-170 (+ torch.distributions semantics)
# --------------------------------------------------------------------------------------

_LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def categorical_sample(logits: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """``torch.distributions.Categorical(logits).sample()`` == ``multinomial(probs, 1)``
    == ``argmax(probs / q)`` with ``q ~ Exp(1)`` and first-index tie-break (SURVEY.md
    Appendix A.9; verified bit-for-bit against torch 2.11 in the golden generator).
    ``logits [B,1,A]``, ``q [B,1,A]`` -> int64 ``[B,1]``."""
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    probs = F.softmax(norm, dim=-1)
    return torch.argmax(probs / q, dim=-1)


def categorical_mode(logits: torch.Tensor) -> torch.Tensor:
    """``Categorical.mode`` = argmax of probs (src/rl8/distributions.py:112-113)."""
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    return torch.argmax(F.softmax(norm, dim=-1), dim=-1)


def categorical_logp(logits: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """src/rl8/distributions.py:118-119 — ``log_softmax.gather(a).sum(-1, keepdim)``."""
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    return norm.gather(-1, actions.unsqueeze(-1)).squeeze(-1).sum(-1, keepdim=True)


def categorical_entropy(logits: torch.Tensor) -> torch.Tensor:
    """src/rl8/distributions.py:115-116 over ``torch.distributions.Categorical.entropy``."""
    norm = logits - logits.logsumexp(dim=-1, keepdim=True)
    min_real = torch.finfo(norm.dtype).min
    clamped = torch.clamp(norm, min=min_real)
    p_log_p = clamped * F.softmax(norm, dim=-1)
    return (-p_log_p.sum(-1)).sum(-1, keepdim=True)


def normal_sample(mean: torch.Tensor, log_std: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """``torch.normal(loc, scale)`` = ``z * scale`` then ``+ loc`` as two roundings
    (SURVEY.md Appendix A.10); ``scale = exp(log_std)`` (src/rl8/distributions.py:142-144)."""
    return z * torch.exp(log_std) + mean


def normal_logp(mean: torch.Tensor, log_std: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``torch.distributions.Normal.log_prob`` summed over the action dim."""
    scale = torch.exp(log_std)
    var = scale**2
    lp = -((x - mean) ** 2) / (2 * var) - scale.log() - _LOG_SQRT_2PI
    return lp.sum(-1, keepdim=True)


def normal_entropy(log_std: torch.Tensor) -> torch.Tensor:
    scale = torch.exp(log_std)
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(scale)).sum(-1, keepdim=True)


def squashed_logp(mean: torch.Tensor, log_std: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """src/rl8/distributions.py:159-167."""
    eps = torch.finfo(x.dtype).eps
    xc = x.clamp(min=-1 + eps, max=1 - eps)
    inv = 0.5 * (xc.log1p() - (-xc).log1p())
    scale = torch.exp(log_std)
    lp = -((inv - mean) ** 2) / (2 * scale**2) - scale.log() - _LOG_SQRT_2PI
    lp = torch.clamp(lp, min=-100, max=100).sum(-1, keepdim=True)
    lp = lp - torch.sum(torch.log(1 - x**2 + eps), dim=-1, keepdim=True)
    return lp


class _Bound:
    """A distribution bound to one forward pass' features.  Shared sub-expressions
    (normalised logits / ``scale``) are built ONCE, as ``torch.distributions`` does in
    its constructor, so autograd accumulates gradients in the reference's order."""

    def __init__(self, kind: str, feats: dict[str, torch.Tensor]) -> None:
        self.kind = kind
        if kind == "categorical":
            logits = feats["logits"]
            self.norm = logits - logits.logsumexp(dim=-1, keepdim=True)
            self._probs: None | torch.Tensor = None
        else:
            self.mean = feats["mean"]
            self.scale = torch.exp(feats["log_std"])

    @property
    def probs(self) -> torch.Tensor:
        if self._probs is None:
            self._probs = F.softmax(self.norm, dim=-1)
        return self._probs

    def sample(self, noise: torch.Tensor) -> torch.Tensor:
        if self.kind == "categorical":
            return torch.argmax(self.probs / noise, dim=-1)
        x = noise * self.scale + self.mean
        return x.tanh() if self.kind == "squashed_normal" else x

    def mode(self) -> torch.Tensor:
        if self.kind == "categorical":
            return torch.argmax(self.probs, dim=-1)
        return self.mean.tanh() if self.kind == "squashed_normal" else self.mean

    def _normal_log_prob(self, x: torch.Tensor) -> torch.Tensor:
        var = self.scale**2
        return -((x - self.mean) ** 2) / (2 * var) - self.scale.log() - _LOG_SQRT_2PI

    def logp(self, a: torch.Tensor) -> torch.Tensor:
        if self.kind == "categorical":
            return self.norm.gather(-1, a.unsqueeze(-1)).squeeze(-1).sum(-1, keepdim=True)
        if self.kind == "normal":
            return self._normal_log_prob(a).sum(-1, keepdim=True)
        eps = torch.finfo(a.dtype).eps
        ac = a.clamp(min=-1 + eps, max=1 - eps)
        inv = 0.5 * (ac.log1p() - (-ac).log1p())
        lp = torch.clamp(self._normal_log_prob(inv), min=-100, max=100).sum(-1, keepdim=True)
        lp -= torch.sum(torch.log(1 - a**2 + eps), dim=-1, keepdim=True)
        return lp

    def entropy(self) -> torch.Tensor:
        if self.kind == "categorical":
            clamped = torch.clamp(self.norm, min=torch.finfo(self.norm.dtype).min)
            return (-(clamped * self.probs).sum(-1)).sum(-1, keepdim=True)
        if self.kind == "normal":
            ent = 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(self.scale)
            return ent.sum(-1, keepdim=True)
        raise NotImplementedError("SquashedNormal has no entropy (distributions.py:153-157)")


class Dist:
    """Distribution family selector: ``"categorical" | "normal" | "squashed_normal"``."""

    def __init__(self, kind: str) -> None:
        assert kind in ("categorical", "normal", "squashed_normal")
        self.kind = kind

    def bind(self, feats: dict[str, torch.Tensor]) -> _Bound:
        return _Bound(self.kind, feats)


# --------------------------------------------------------------------------------------
# collect() — src/rl8/algorithms/_feedforward.py:301-441
# --------------------------------------------------------------------------------------


def new_buffer(N: int, T: int, obs_dim: int, action_kind: str) -> dict[str, torch.Tensor]:
    """``buffer_spec.zero([N, T+1])`` (src/rl8/algorithms/_feedforward.py:239-256)."""
    act_dtype = torch.int64 if action_kind == "discrete" else torch.float32
    z = lambda: torch.zeros(N, T + 1, 1)  # noqa: E731
    return {
        "obs": torch.zeros(N, T + 1, obs_dim),
        "rewards": z(),
        "actions": torch.zeros(N, T + 1, 1, dtype=act_dtype),
        "logp": z(),
        "values": z(),
        "advantages": z(),
        "returns": z(),
        "reversed_discounted_returns": z(),
    }


def collect(
    p: Params,
    env: OracleEnv,
    buf: dict[str, torch.Tensor],
    dist: Dist,
    noise: None | torch.Tensor,
    *,
    gamma: float = 0.95,
    reset: bool = True,
    reset_state: None | torch.Tensor = None,
    deterministic: bool = False,
    normalize_rewards: bool = True,
    noise_fn: None | Callable[[int], torch.Tensor] = None,
) -> dict[str, float]:
    """One rollout of ``T`` steps into ``buf`` (in place).  ``noise[t]`` is the injected
    draw for step ``t`` (``[T, N, 1, A]`` categorical / ``[T, N, 1]`` normal).

    Returns the CollectStats values (src/rl8/algorithms/_feedforward.py:411-439) plus
    ``"reward_scale"`` (428-436).
    """
    T = buf["obs"].shape[1] - 1
    rdr = buf["reversed_discounted_returns"]
    with torch.no_grad():
        if reset:
            buf["obs"][:, 0] = env.reset(reset_state)
            if normalize_rewards:
                rdr[:, 0] = 0.0
        else:
            buf["obs"][:, 0] = buf["obs"][:, -1]
            if normalize_rewards:
                rdr[:, 0] = rdr[:, -1]
        for t in range(T):
            feats, values = model_forward(p, buf["obs"][:, t])
            d = dist.bind(feats)
            if deterministic:
                actions = d.mode()
            else:
                actions = d.sample(noise_fn(t) if noise_fn else noise[t])
            logp = d.logp(actions)
            obs, rewards = env.step(actions)
            if normalize_rewards:
                rdr[:, t + 1] = gamma * rdr[:, t] + rewards
            buf["actions"][:, t] = actions
            buf["logp"][:, t] = logp
            buf["values"][:, t] = values
            buf["rewards"][:, t] = rewards
            buf["obs"][:, t + 1] = obs
        _, values = model_forward(p, buf["obs"][:, -1])
        buf["values"][:, -1] = values

        rewards = buf["rewards"][:, :-1]
        returns = torch.sum(rewards, dim=1)
        returns_std, returns_mean = torch.std_mean(returns)
        rewards_std, rewards_mean = torch.std_mean(rewards)
        stats = {
            "returns/min": float(torch.min(returns)),
            "returns/max": float(torch.max(returns)),
            "returns/mean": float(returns_mean),
            "returns/std": float(returns_std),
            "rewards/min": float(torch.min(rewards)),
            "rewards/max": float(torch.max(rewards)),
            "rewards/mean": float(rewards_mean),
            "rewards/std": float(rewards_std),
            "reward_scale": float(torch.std(rdr[:, 1:])) if normalize_rewards else 1.0,
        }
    return stats


# --------------------------------------------------------------------------------------
# GAE — src/rl8/nn/functional.py:50-123
# --------------------------------------------------------------------------------------


def gae(
    rewards: torch.Tensor,
    values: torch.Tensor,
    *,
    gamma: float = 0.95,
    gae_lambda: float = 0.95,
    reward_scale: float = 1.0,
    normalize_advantages: bool = True,
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``rewards, values [N, T+1, 1]`` -> ``(scaled_rewards, advantages, returns)``.

    Order of operations follows the reference exactly: rewards are divided by
    ``reward_scale + 1e-8`` (106, and written back by the caller), the reverse scan
    starts from ``A_T = 0`` (107-115), returns are ``A + V`` over all ``T+1`` slots from
    the *un-normalised* advantages (117), and only ``A[:, :-1]`` is normalised with the
    unbiased std (118-122).
    """
    r = rewards / (reward_scale + 1e-8)
    adv = torch.zeros_like(r)
    prev: Any = 0.0
    for t in reversed(range(r.shape[1] - 1)):
        delta = r[:, t] + (gamma * values[:, t + 1] - values[:, t])
        adv[:, t] = prev = delta + (gamma * gae_lambda * prev)
    ret = adv + values
    if normalize_advantages:
        std, mean = torch.std_mean(adv[:, :-1])
        adv[:, :-1] = (adv[:, :-1] - mean) / (std + 1e-8)
    return r, adv, ret


# --------------------------------------------------------------------------------------
# PPO losses — src/rl8/nn/functional.py:259-363, KL: _feedforward.py:552-559
# --------------------------------------------------------------------------------------


def ppo_losses(
    logp_new: torch.Tensor,
    values_new: torch.Tensor,
    entropy: None | torch.Tensor,
    logp_old: torch.Tensor,
    advantages: torch.Tensor,
    returns: torch.Tensor,
    *,
    clip_param: float = 0.2,
    dual_clip_param: None | float = None,
    entropy_coeff: float = 0.0,
    vf_clip_param: float = 5.0,
    vf_coeff: float = 1.0,
) -> dict[str, torch.Tensor]:
    """All inputs ``[M, 1]``.  ``losses["policy"]`` is the surrogate *objective*;
    ``total = vf_coeff*vf - policy - entropy_coeff*entropy`` (349-352)."""
    ratio = torch.exp(logp_new - logp_old)
    vf = torch.mean(
        torch.clamp(F.smooth_l1_loss(values_new, returns, reduction="none"), 0.0, vf_clip_param)
    )
    s1 = advantages * ratio
    s2 = advantages * torch.clamp(ratio, 1 - clip_param, 1 + clip_param)
    if dual_clip_param:
        c1 = torch.min(s1, s2)
        c2 = torch.max(c1, dual_clip_param * advantages)
        pol = torch.where(advantages < 0, c2, c1).mean()
    else:
        pol = torch.min(s1, s2).mean()
    total = vf_coeff * vf - pol
    if entropy_coeff != 0:
        assert entropy is not None
        ent = entropy.mean()
        total = total - entropy_coeff * ent
    else:
        ent = torch.tensor(0.0)
    return {"entropy": ent, "policy": pol, "vf": vf, "total": total}


def approx_kl(logp_new: torch.Tensor, logp_old: torch.Tensor) -> torch.Tensor:
    """src/rl8/algorithms/_feedforward.py:552-559 — ``mean((ratio-1) - log ratio)``."""
    lr = logp_new - logp_old
    return torch.mean((torch.exp(lr) - 1) - lr)


# --------------------------------------------------------------------------------------
# step() — src/rl8/algorithms/_feedforward.py:443-615
# --------------------------------------------------------------------------------------


def minibatch_losses(
    p: Params,
    dist: Dist,
    mb: dict[str, torch.Tensor],
    *,
    clip_param: float,
    dual_clip_param: None | float,
    entropy_coeff: float,
    vf_clip_param: float,
    vf_coeff: float,
) -> tuple[dict[str, torch.Tensor], torch.Tensor]:
    feats, values = model_forward(p, mb["obs"])
    d = dist.bind(feats)
    logp_new = d.logp(mb["actions"])
    ent = d.entropy() if entropy_coeff != 0 else None
    losses = ppo_losses(
        logp_new,
        values,
        ent,
        mb["logp"],
        mb["advantages"],
        mb["returns"],
        clip_param=clip_param,
        dual_clip_param=dual_clip_param,
        entropy_coeff=entropy_coeff,
        vf_clip_param=vf_clip_param,
        vf_coeff=vf_coeff,
    )
    with torch.no_grad():
        kl = approx_kl(d.logp(mb["actions"]), mb["logp"])
    return losses, kl


class _RunningMean:
    """src/rl8/_utils.py:228-256 cumulative average."""

    def __init__(self) -> None:
        self.avg, self.n = 0.0, 0

    def update(self, v: float) -> None:
        self.avg = (v + self.n * self.avg) / (self.n + 1)
        self.n += 1


def step(
    p: Params,
    buf: dict[str, torch.Tensor],
    dist: Dist,
    opt_state: dict[str, Any],
    *,
    reward_scale: float,
    gamma: float = 0.95,
    gae_lambda: float = 0.95,
    normalize_advantages: bool = True,
    sgd_minibatch_size: None | int = None,
    num_sgd_iters: int = 4,
    shuffle: bool = False,
    accumulate_grads: bool = False,
    clip_param: float = 0.2,
    dual_clip_param: None | float = None,
    entropy_coeff: float = 0.0,
    vf_clip_param: float = 5.0,
    vf_coeff: float = 1.0,
    target_kl_div: None | float = None,
    max_grad_norm: float = 5.0,
    lr: float = 1e-3,
    betas: tuple[float, float] = (0.9, 0.999),
    eps: float = 1e-8,
    perms: None | list[torch.Tensor] = None,
    grad_hook: None | Callable[[Params], None] = None,
    optimizer_cls: Any = torch.optim.Adam,
    optimizer_config: None | dict[str, Any] = None,
) -> dict[str, float]:
    """GAE + PPO epochs + clip + Adam, in place on ``p`` (leaf tensors) and ``buf``.

    ``optimizer_cls(params, **optimizer_config)`` replaces the default Adam the way
    ``AlgorithmConfig.optimizer_cls / optimizer_config`` do (_feedforward.py:257-260).

    ``opt_state`` holds a persistent ``torch.optim.Adam`` across calls (created on first
    use) — the reference uses the same library optimizer
    (src/rl8/algorithms/_feedforward.py:257-260, 585-593).  ``perms`` optionally injects
    the per-epoch row permutations (``Batcher``, src/rl8/_utils.py:211-218).
    """
    N, Tp1 = buf["rewards"].shape[:2]
    T = Tp1 - 1
    r, adv, ret = gae(
        buf["rewards"],
        buf["values"],
        gamma=gamma,
        gae_lambda=gae_lambda,
        reward_scale=reward_scale,
        normalize_advantages=normalize_advantages,
    )
    buf["rewards"], buf["advantages"], buf["returns"] = r, adv, ret

    flat = {
        "obs": buf["obs"][:, :-1].flatten(end_dim=1),
        "actions": buf["actions"][:, :-1].reshape(N * T, -1),
        "logp": buf["logp"][:, :-1].reshape(N * T, 1),
        "advantages": adv[:, :-1].reshape(N * T, 1),
        "returns": ret[:, :-1].reshape(N * T, 1),
    }
    M = sgd_minibatch_size or N * T
    num_mb = (N * T) // M
    accum = num_mb if accumulate_grads else 1

    for v in p.values():
        v.requires_grad_(True)
    if "adam" not in opt_state:
        if optimizer_config is not None or optimizer_cls is not torch.optim.Adam:
            opt_state["adam"] = optimizer_cls(list(p.values()), **(optimizer_config or {"lr": lr}))
            opt_state["own_lr"] = True  # the lr of optimizer_config stands (no schedule injected by the caller)
        else:
            opt_state["adam"] = torch.optim.Adam(list(p.values()), lr=lr, betas=betas, eps=eps)
    adam = opt_state["adam"]
    for g in adam.param_groups:
        g["lr"] = g["lr"] if opt_state.get("own_lr") else lr

    keys = ("losses/entropy", "losses/policy", "losses/vf", "losses/total", "monitors/kl_div")
    sums = {k: 0.0 for k in keys}
    means = {k: _RunningMean() for k in keys}
    coeff_means = {"coefficients/entropy": _RunningMean(), "coefficients/vf": _RunningMean()}
    stop = False
    for epoch in range(num_sgd_iters):
        if perms is not None:
            order = perms[epoch]
        elif shuffle:
            order = torch.randperm(N * T)
        else:
            order = torch.arange(N * T)
        for i, idx in enumerate(torch.split(order, M)):
            step_now = (i + 1) % accum == 0
            mb = {k: v[idx] for k, v in flat.items()}
            losses, kl = minibatch_losses(
                p,
                dist,
                mb,
                clip_param=clip_param,
                dual_clip_param=dual_clip_param,
                entropy_coeff=entropy_coeff,
                vf_clip_param=vf_clip_param,
                vf_coeff=vf_coeff,
            )
            losses = {k: v / accum for k, v in losses.items()}
            kl_f = float(kl)
            sums["losses/entropy"] += float(losses["entropy"].detach())
            sums["losses/policy"] += float(losses["policy"].detach())
            sums["losses/vf"] += float(losses["vf"].detach())
            sums["losses/total"] += float(losses["total"].detach())
            sums["monitors/kl_div"] += kl_f / accum
            coeff_means["coefficients/entropy"].update(entropy_coeff)
            coeff_means["coefficients/vf"].update(vf_coeff)
            if step_now:
                for k in keys:
                    means[k].update(sums[k])
                    sums[k] = 0.0
            if target_kl_div is not None and kl_f > 1.5 * target_kl_div:
                stop = True
                break
            losses["total"].backward()
            if step_now:
                if grad_hook is not None:
                    grad_hook({k: v.grad.detach().clone() for k, v in p.items()})
                torch.nn.utils.clip_grad_norm_(list(p.values()), max_grad_norm)
                adam.step()
                adam.zero_grad()
        if stop:
            break
    for v in p.values():
        v.requires_grad_(False)

    # Buffer is re-zeroed with only obs[:, -1] kept (603-610).
    final_obs = buf["obs"][:, -1].clone()
    for k, v in buf.items():
        buf[k] = torch.zeros_like(v)
    buf["obs"][:, -1] = final_obs
    out = {k: m.avg for k, m in means.items()}
    out.update({k: m.avg for k, m in coeff_means.items()})
    return out
