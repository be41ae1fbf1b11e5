"""Recurrent PPO: ``RecurrentAlgorithmConfig(...).build(env_cls)``, ``collect()``, ``step()``.

Drop-in for the reference's recurrent family -- default LSTM models
(src/rl8/models/_recurrent.py:169-341), ``RecurrentPolicy``
(src/rl8/policies/_recurrent.py:20-164) and ``RecurrentAlgorithm``
(src/rl8/algorithms/_recurrent.py:29-756) -- on sm_100a kernels:

* ``collect()`` -> one ``rl8_lstm_collect`` call: per step the LSTM cell, both heads,
  sampling, log-probability, the env transition and the buffer / state-slab writes;
* ``step()`` -> the shared GAE kernels, then per minibatch of ``seq_len`` SEQUENCES
  ``rl8_lstm_ppo_minibatch`` (replay from the stored chunk-start states, PPO losses,
  hand-derived back-propagation through time) and ``rl8_clip_adam``.

Everything that is not LSTM-specific (reset cadence, statistics, GAE, the epoch loop, early
stopping, gradient accumulation, Adam, multi-GPU reductions) is inherited from
:class:`rl8_b200.algorithms.Algorithm`.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Literal, Mapping

import torch
import torch.nn as nn
import torch.optim as optim

from . import _lib, parallel
from .algorithms import Algorithm, _NoopGradScaler, make_flat_optimizer
from .buffer import RolloutBuffer
from .data import DataKeys, Device, RecurrentAlgorithmHparams, RecurrentAlgorithmState
from .distributions import Distribution
from .env import Env, KernelEnv
from .models import Model, _small_head
from .policies import PolicyExport
from .schedulers import EntropyScheduler, LRScheduler, ScheduleKind
from .specs import Categorical, Composite, TensorSpec, Unbounded
from .trainers import Trainer

# --------------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------------


class RecurrentModel(Model):
    """Base of the fused default recurrent models: ONE ``nn.LSTM(D, 256)`` whose latents feed
    the policy head(s) and the value head (src/rl8/models/_recurrent.py:19-137)."""

    #: Spec of the recurrent states (``hidden_states`` / ``cell_states`` ``[num_layers, H]``).
    state_spec: Composite

    def __init__(
        self,
        observation_spec: TensorSpec,
        action_spec: TensorSpec,
        /,
        *,
        hidden_size: int = 256,
        num_layers: int = 1,
        bias: bool = True,
    ) -> None:
        super().__init__(observation_spec, action_spec)
        if hidden_size != 256 or num_layers != 1 or not bias:
            raise NotImplementedError(
                "the fused path implements the reference's default recurrent architecture only:"
                f" hidden_size=256, num_layers=1, bias=True (got {hidden_size}, {num_layers}, {bias})"
            )
        self.hidden = hidden_size
        dev = action_spec.device
        self.state_spec = Composite(
            {
                DataKeys.HIDDEN_STATES: Unbounded((num_layers, hidden_size), device=dev),
                DataKeys.CELL_STATES: Unbounded((num_layers, hidden_size), device=dev),
            }
        )
        self.lstm = nn.LSTM(
            observation_spec.shape[0], hidden_size, num_layers=num_layers, bias=bias, batch_first=True
        )

    @staticmethod
    def default_model_cls(  # type: ignore[override]
        observation_spec: TensorSpec, action_spec: TensorSpec, /
    ) -> type["RecurrentModel"]:
        """src/rl8/models/_recurrent.py:42-73."""
        if not isinstance(observation_spec, Unbounded):
            raise TypeError(f"Observation spec {observation_spec} has no default model support.")
        if len(observation_spec.shape) != 1:
            raise TypeError("Default models support 1D observations only.")
        if isinstance(action_spec, Categorical):
            return DefaultDiscreteRecurrentModel
        if isinstance(action_spec, Unbounded):
            return DefaultContinuousRecurrentModel
        raise TypeError(f"Action spec {action_spec} has no default model support.")

    def init_states(self, n: int, /) -> dict[str, torch.Tensor]:
        """Zero states ``[n, num_layers, H]`` (src/rl8/models/_recurrent.py:105-121)."""
        return self.state_spec.to(self.flat_params.device).zero([n])

    def struct_for(self, flat: torch.Tensor) -> _lib.LstmModel:  # type: ignore[override]
        m = _lib.LstmModel()
        m.D = self.observation_spec.shape[0]
        m.H = self.hidden
        m.P = self.head_width
        base = flat.data_ptr()
        for field, off in self._offsets.items():
            setattr(m, field, base + 4 * off)
        return m

    def _head_segments(self) -> list[tuple[str, list[nn.Parameter]]]:
        raise NotImplementedError

    def _kernel_segments(self) -> list[tuple[str, list[nn.Parameter]]]:
        lstm = self.lstm
        return [
            ("w_ih", [lstm.weight_ih_l0]), ("w_hh", [lstm.weight_hh_l0]),
            ("b_ih", [lstm.bias_ih_l0]), ("b_hh", [lstm.bias_hh_l0]),
            *self._head_segments(),
        ]


class DefaultDiscreteRecurrentModel(RecurrentModel):
    """1-D observations, one discrete action (src/rl8/models/_recurrent.py:259-341)."""

    def __init__(self, observation_spec: Unbounded, action_spec: Categorical, /, **config: Any):
        super().__init__(observation_spec, action_spec, **config)
        n_act = max(1, action_spec.shape[0] if len(action_spec.shape) else 1)
        if n_act != 1:
            raise NotImplementedError("the fused path supports a single discrete action")
        self.head_width = action_spec.space.n
        self.feature_head = _small_head(self.hidden, self.head_width)
        self.vf_head = nn.Linear(self.hidden, 1)

    def _head_segments(self) -> list[tuple[str, list[nn.Parameter]]]:
        return [
            ("pi_w", [self.feature_head.weight]), ("pi_b", [self.feature_head.bias]),
            ("vf_w", [self.vf_head.weight]), ("vf_b", [self.vf_head.bias]),
        ]

    def features_from_head(self, head: torch.Tensor) -> dict[str, torch.Tensor]:
        return {"logits": head.reshape(-1, 1, self.head_width)}


class DefaultContinuousRecurrentModel(RecurrentModel):
    """1-D observations, one continuous action (src/rl8/models/_recurrent.py:169-256)."""

    def __init__(self, observation_spec: Unbounded, action_spec: Unbounded, /, **config: Any):
        super().__init__(observation_spec, action_spec, **config)
        if action_spec.shape[0] != 1:
            raise NotImplementedError("the fused path supports a single continuous action")
        self.head_width = 2
        self.action_mean = _small_head(self.hidden, 1)
        self.action_log_std = _small_head(self.hidden, 1)
        self.vf_model = nn.Linear(self.hidden, 1)

    def _head_segments(self) -> list[tuple[str, list[nn.Parameter]]]:
        return [
            ("pi_w", [self.action_mean.weight, self.action_log_std.weight]),
            ("pi_b", [self.action_mean.bias, self.action_log_std.bias]),
            ("vf_w", [self.vf_model.weight]), ("vf_b", [self.vf_model.bias]),
        ]

    def features_from_head(self, head: torch.Tensor) -> dict[str, torch.Tensor]:
        return {"mean": head[:, 0:1].contiguous(), "log_std": head[:, 1:2].contiguous()}


# --------------------------------------------------------------------------------------------
# policy
# --------------------------------------------------------------------------------------------


class RecurrentPolicy(PolicyExport):
    """Union of a recurrent model and an action distribution
    (src/rl8/policies/_recurrent.py:20-164)."""

    def __init__(
        self,
        observation_spec: TensorSpec,
        action_spec: TensorSpec,
        /,
        *,
        model: None | RecurrentModel = None,
        model_cls: None | type[RecurrentModel] = None,
        model_config: None | dict[str, Any] = None,
        distribution_cls: None | type[Distribution] = None,
        device: Device = "cpu",
    ) -> None:
        self.model_config = model_config or {}
        if model and model_cls:
            raise ValueError(
                "`model` and `model_cls` args are mutually exclusive."
                "Provide one or the other, but not both."
            )
        if torch.device(device).type != "cuda":
            raise RuntimeError("rl8_b200 policies run on CUDA only; there is no CPU path.")
        if model is None:
            model_cls = model_cls or RecurrentModel.default_model_cls(observation_spec, action_spec)
            model = model_cls(observation_spec, action_spec, **self.model_config)
        if not isinstance(model, RecurrentModel):
            raise NotImplementedError(
                "custom model classes are outside the fused hot path; use the default models"
            )
        self.model = model.flatten_(device)
        self.distribution_cls = distribution_cls or Distribution.default_dist_cls(action_spec)
        if not (isinstance(self.distribution_cls, type) and issubclass(self.distribution_cls, Distribution)):
            raise NotImplementedError("custom distributions must subclass rl8_b200 distributions")
        self.device = torch.device(device)
        self.precision = _lib.PREC_FP32
        self._lib = _lib.load()
        self._ws: None | torch.Tensor = None

    @property
    def state_spec(self) -> Composite:
        return self.model.state_spec

    def init_states(self, n: int, /) -> dict[str, torch.Tensor]:
        return self.model.init_states(n)

    def _continuous(self) -> bool:
        return self.distribution_cls.rl8_kind != _lib.DIST_CATEGORICAL

    def step_net(
        self, obs: torch.Tensor, h: torch.Tensor, c: torch.Tensor
    ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """One LSTM step + heads: ``obs [B, D]`` (any strides), ``h, c [B, H]`` contiguous ->
        ``(head [B, P], values [B, 1], h', c')``."""
        _lib.require_cuda(obs, "obs")
        B, H = obs.shape[0], self.model.hidden
        h, c = h.contiguous().float(), c.contiguous().float()
        h2, c2 = torch.empty_like(h), torch.empty_like(c)
        head = torch.empty(B, self.model.head_width, device=obs.device)
        values = torch.empty(B, 1, device=obs.device)
        nbytes = B * 4 * H * 4
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        m = self.model.struct_for(self.model.flat_params)
        rc = self._lib.rl8_lstm_forward(
            m, _lib.ptr(obs), obs.stride(0), obs.stride(1), _lib.ptr(h), _lib.ptr(c), _lib.ptr(h2),
            _lib.ptr(c2), _lib.ptr(head), _lib.ptr(values), B, int(self._continuous()),
            self.precision, _lib.ptr(self._ws), self._ws.numel(), _lib.stream(),
        )
        _lib.check(rc, "rl8_lstm_forward")
        return head, values, h2, c2

    def sample(
        self,
        batch: Mapping[str, torch.Tensor],
        /,
        states: None | Mapping[str, torch.Tensor] = None,
        *,
        deterministic: bool = False,
        inplace: bool = False,
        requires_grad: bool = False,
        return_actions: bool = True,
        return_logp: bool = False,
        return_values: bool = False,
    ) -> tuple[dict[str, Any], dict[str, torch.Tensor]]:
        """Sample the policy on ``batch["obs"] [B, T, D]`` starting from ``states[...][:, 0]``
        (``[B, T, 1, H]``; zeros when ``None``).  Outputs have ``B*T`` rows (row ``b*T + t``),
        the returned states ``[B, 1, H]`` are those after the last step."""
        if requires_grad:
            raise NotImplementedError(
                "autograd through RecurrentPolicy.sample is not part of the fused path; gradients"
                " are produced by RecurrentAlgorithm.step's fused backward"
            )
        obs = batch[DataKeys.OBS]
        if obs.dtype != torch.float32:
            obs = obs.float()
        B, T = obs.shape[:2]
        H = self.model.hidden
        if states is None:
            h = torch.zeros(B, H, device=obs.device)
            c = torch.zeros(B, H, device=obs.device)
        else:
            h = states[DataKeys.HIDDEN_STATES][:, 0].reshape(B, H)
            c = states[DataKeys.CELL_STATES][:, 0].reshape(B, H)
        heads, values = [], []
        for t in range(T):
            head, val, h, c = self.step_net(obs[:, t], h, c)
            heads.append(head)
            values.append(val)
        head = torch.stack(heads, dim=1).reshape(B * T, -1)
        out: dict[str, Any] = dict(batch) if inplace else {}
        features = self.model.features_from_head(head)
        out[DataKeys.FEATURES] = features
        if return_actions:
            dist = self.distribution_cls(features, self.model)
            actions = dist.deterministic_sample() if deterministic else dist.sample()
            out[DataKeys.ACTIONS] = actions
            if return_logp:
                out[DataKeys.LOGP] = dist.logp(actions)
        if return_values:
            self.model._value = torch.stack(values, dim=1).reshape(B * T, 1)
            out[DataKeys.VALUES] = self.model._value
        out_states = {
            DataKeys.HIDDEN_STATES: h.reshape(B, 1, H),
            DataKeys.CELL_STATES: c.reshape(B, 1, H),
        }
        return out, out_states


# --------------------------------------------------------------------------------------------
# algorithm
# --------------------------------------------------------------------------------------------


@dataclass
class RecurrentAlgorithmConfig:
    """Configuration of a recurrent PPO algorithm (same fields and defaults as the reference,
    src/rl8/algorithms/_recurrent.py:29-186)."""

    model: None | RecurrentModel = None
    model_cls: None | type[RecurrentModel] = None
    model_config: None | dict[str, Any] = None
    distribution_cls: None | type[Distribution] = None
    horizon: int = 32
    horizons_per_env_reset: int = 1
    num_envs: int = 8192
    #: Truncated back-propagation-through-time length; must divide ``horizon``.
    seq_len: int = 4
    #: Sequences between recurrent-state re-initialisations.
    seqs_per_state_reset: int = 8
    optimizer_cls: type[optim.Optimizer] = optim.Adam
    optimizer_config: None | dict[str, Any] = None
    accumulate_grads: bool = False
    #: ``True``: the LSTM's hidden-to-hidden GEMMs (forward and both backward contractions) run in
    #: bf16 on tcgen05 with fp32 accumulation; ``False``: everything in fp32 on CUDA cores.
    enable_amp: bool = False
    lr_schedule: None | list[tuple[int, float]] = None
    lr_schedule_kind: ScheduleKind = "step"
    entropy_coeff: float = 0.0
    entropy_coeff_schedule: None | list[tuple[int, float]] = None
    entropy_coeff_schedule_kind: ScheduleKind = "step"
    gae_lambda: float = 0.95
    gamma: float = 0.95
    #: SEQUENCES per minibatch; ``None``: all ``num_envs * (horizon // seq_len)`` of them.
    sgd_minibatch_size: None | int = None
    num_sgd_iters: int = 4
    shuffle_minibatches: bool = True
    clip_param: float = 0.2
    vf_clip_param: float = 5.0
    dual_clip_param: None | float = None
    vf_coeff: float = 1.0
    target_kl_div: None | float = None
    max_grad_norm: float = 5.0
    normalize_advantages: bool = True
    normalize_rewards: bool = True
    device: Device | Literal["auto"] = "auto"

    def build(self, env_cls: Any) -> "RecurrentAlgorithm":
        """Build and validate a :class:`RecurrentAlgorithm`."""
        algo = RecurrentAlgorithm(env_cls, config=self)
        algo.validate()
        return algo


class RecurrentAlgorithm(Algorithm):
    """Recurrent PPO over a tensor-batched environment with the rollout and the update
    (truncated back-propagation through time) on CUDA kernels."""

    _stats_reward_t0 = 1  # statistics over rewards[:, 1:-1] (src/rl8/algorithms/_recurrent.py:449)
    _graph_capable = False  # the TBPTT host loop allocates per minibatch: not captured

    def __init__(self, env_cls: Any, /, config: None | RecurrentAlgorithmConfig = None) -> None:
        config = config or RecurrentAlgorithmConfig()
        if not torch.cuda.is_available():
            raise RuntimeError(
                "rl8_b200 needs a CUDA device (B200, sm_100a): there is no CPU path."
            )
        device = "cuda" if config.device == "auto" else str(config.device)
        if torch.device(device).type != "cuda":
            raise RuntimeError(f"device={device!r}: rl8_b200 runs on CUDA only.")
        if device == "cuda":
            device = f"cuda:{torch.cuda.current_device()}"
        self._lib = _lib.load()
        max_num_envs = getattr(env_cls, "max_num_envs", config.num_envs)
        num_envs = min(config.num_envs, max_num_envs)
        horizon = min(config.horizon, getattr(env_cls, "max_horizon", 1_000_000))
        self.env: Env = env_cls(num_envs, horizon, device=device)
        for name in ("observation_spec", "action_spec"):
            spec = getattr(self.env, name)
            if not isinstance(spec, (Unbounded, Categorical)):
                raise TypeError(f"`{name}` must be an Unbounded or Categorical spec")
        self.policy = RecurrentPolicy(  # type: ignore[assignment]
            self.env.observation_spec,
            self.env.action_spec,
            model=config.model,
            model_cls=config.model_cls,
            model_config=config.model_config,
            distribution_cls=config.distribution_cls,
            device=device,
        )
        self._fused_model = True  # the default recurrent models are always on the kernels
        # enable_amp: the 256 x 1024 LSTM contractions in bf16 on tcgen05 (fp32 accumulate)
        self.policy.precision = _lib.PREC_BF16 if config.enable_amp else _lib.PREC_FP32
        self.buffer_spec = Composite(
            {
                DataKeys.OBS: self.env.observation_spec,
                DataKeys.STATES: self.policy.state_spec,  # type: ignore[dict-item]
                DataKeys.REWARDS: Unbounded(1, device=device),
                DataKeys.ACTIONS: self.env.action_spec,
                DataKeys.LOGP: Unbounded(1, device=device),
                DataKeys.VALUES: Unbounded(1, device=device),
                DataKeys.ADVANTAGES: Unbounded(1, device=device),
                DataKeys.RETURNS: Unbounded(1, device=device),
            }
        )
        if config.normalize_rewards:
            self.buffer_spec.set(DataKeys.REVERSED_DISCOUNTED_RETURNS, Unbounded(1, device=device))
        self.buffer = RolloutBuffer(self.buffer_spec, num_envs, horizon, device)
        optimizer_config = dict(config.optimizer_config or {"lr": 1e-3})
        self._grads = torch.zeros_like(self.policy.model.flat_params)
        self.optimizer, moments = make_flat_optimizer(self.policy.model, self._grads, config.optimizer_cls,
                                                      optimizer_config)
        self._exp_avg, self._exp_avg_sq = moments if moments is not None else (None, None)
        parallel.sync_replicas(self.policy.model)  # replicas start from rank 0's parameters
        self._scale_dev = torch.ones(2, dtype=torch.float32, device=device)  # reward scale of the last collect()
        self._grad_norm = torch.zeros(1, device=device)
        self.lr_scheduler = LRScheduler(
            self.optimizer, schedule=config.lr_schedule, kind=config.lr_schedule_kind
        )
        self.entropy_scheduler = EntropyScheduler(
            config.entropy_coeff,
            schedule=config.entropy_coeff_schedule,
            kind=config.entropy_coeff_schedule_kind,
        )
        sgd_minibatch_size = config.sgd_minibatch_size or num_envs * (horizon // config.seq_len)
        self.hparams = RecurrentAlgorithmHparams(  # type: ignore[assignment]
            accumulate_grads=config.accumulate_grads,
            clip_param=config.clip_param,
            device=device,
            dual_clip_param=config.dual_clip_param,
            enable_amp=config.enable_amp,
            gae_lambda=config.gae_lambda,
            gamma=config.gamma,
            horizon=horizon,
            horizons_per_env_reset=config.horizons_per_env_reset,
            max_grad_norm=config.max_grad_norm,
            normalize_advantages=config.normalize_advantages,
            normalize_rewards=config.normalize_rewards,
            num_envs=num_envs,
            num_sgd_iters=config.num_sgd_iters,
            seq_len=config.seq_len,
            seqs_per_state_reset=config.seqs_per_state_reset,
            sgd_minibatch_size=sgd_minibatch_size,
            shuffle_minibatches=config.shuffle_minibatches,
            target_kl_div=config.target_kl_div,
            vf_clip_param=config.vf_clip_param,
            vf_coeff=config.vf_coeff,
        ).validate()
        self.state = RecurrentAlgorithmState()  # type: ignore[assignment]
        self.grad_scaler = _NoopGradScaler(config.enable_amp)
        self.device = torch.device(device)
        self._fused_env = isinstance(self.env, KernelEnv)
        self._ws: dict[str, torch.Tensor] = {}
        self._stats_acc = torch.zeros(16, dtype=torch.float64, device=device)
        init = [0.0] * 16
        init[6] = init[8] = math.inf
        init[7] = init[9] = -math.inf
        self._stats_init = torch.tensor(init, dtype=torch.float64, device=device)
        self._moments = torch.zeros(3, dtype=torch.float64, device=device)
        max_updates = self.hparams.num_sgd_iters * self.hparams.num_minibatches
        self._loss_sums = torch.zeros(max_updates, 5, dtype=torch.float64, device=device)
        self._on_grads: Any = None
        self.last_launches = {"collect": 0, "step": 0}

    # -- collect ---------------------------------------------------------------------------------
    def _pre_collect(self) -> None:
        """``states[:, 0] = states[:, -1]`` (src/rl8/algorithms/_recurrent.py:380-382)."""
        T = self.hparams.horizon
        for k in (DataKeys.HIDDEN_STATES, DataKeys.CELL_STATES):
            self.buffer.hm[k][0].copy_(self.buffer.hm[k][T])

    def _post_collect(self) -> None:
        self.state.seqs += self.hparams.horizon // self.hparams.seq_len  # :430-431

    def _state_reset_at(self, t: int) -> bool:
        hp = self.hparams
        if t % hp.seq_len:
            return False
        seqs = self.state.seqs + t // hp.seq_len
        if seqs and hp.seqs_per_state_reset < 0:
            return False
        return seqs % hp.seqs_per_state_reset == 0

    def _collect_fused(self, noise: None | torch.Tensor, deterministic: bool) -> None:
        hp, buf = self.hparams, self.buffer
        model = self.policy.model
        m = model.struct_for(model.flat_params)
        prec = self.policy.precision
        nbytes = int(self._lib.rl8_lstm_collect_workspace(m, hp.num_envs, hp.horizon, prec))
        if nbytes < 0:
            _lib.check(nbytes, "rl8_lstm_collect_workspace")
        ws = self._workspace("collect", nbytes)
        rro = _lib.RecurrentRollout()
        rro.ro = self._rollout_struct(noise, deterministic)
        rro.hidden = buf.hm[DataKeys.HIDDEN_STATES].data_ptr()
        rro.cell = buf.hm[DataKeys.CELL_STATES].data_ptr()
        rro.seq_len, rro.seqs_per_state_reset = hp.seq_len, hp.seqs_per_state_reset
        rro.seqs = self.state.seqs
        rc = self._lib.rl8_lstm_collect(m, rro, prec, _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "rl8_lstm_collect")
        self.last_launches["collect"] = 5 * hp.horizon + 3

    def _collect_generic(self, noise: None | torch.Tensor, deterministic: bool) -> None:
        """Rollout with a user-defined (Python / torch) environment: the LSTM step, heads,
        sampling and log-probabilities run on this library's kernels; ``env.step`` is the
        user's code (the reference's Env plug-in point)."""
        hp, buf = self.hparams, self.buffer
        N, T = hp.num_envs, hp.horizon
        obs_hm, act_hm = buf.hm[DataKeys.OBS], buf.hm[DataKeys.ACTIONS]
        hs, cs = buf.hm[DataKeys.HIDDEN_STATES], buf.hm[DataKeys.CELL_STATES]
        rdr_hm = buf.hm.get(DataKeys.REVERSED_DISCOUNTED_RETURNS)
        P = self.policy.model.head_width
        kind = self.policy.distribution_cls.rl8_kind
        launches = 0
        for t in range(T):
            if self._state_reset_at(t):
                hs[t].zero_()
                cs[t].zero_()
            head, values, h2, c2 = self.policy.step_net(obs_hm[t].T, hs[t], cs[t])
            hs[t + 1].copy_(h2)
            cs[t + 1].copy_(c2)
            buf.hm[DataKeys.VALUES][t].copy_(values.view(N))
            nz = None if noise is None else noise[t]
            rc = self._lib.rl8_dist_sample(
                kind, _lib.ptr(head), P, _lib.ptr(nz), int(deterministic), _lib.ptr(act_hm[t]),
                _lib.ptr(buf.hm[DataKeys.LOGP][t]), N, _lib.stream(),
            )
            _lib.check(rc, "rl8_dist_sample")
            out = self.env.step(act_hm[t].view(N, 1))
            rewards = out[DataKeys.REWARDS].reshape(N)
            if rdr_hm is not None:
                torch.add(rewards, rdr_hm[t], alpha=hp.gamma, out=rdr_hm[t + 1])
            buf.hm[DataKeys.REWARDS][t].copy_(rewards)
            obs_hm[t + 1].copy_(out[DataKeys.OBS].reshape(N, -1).T)
            launches += 5
        _, values, _, _ = self.policy.step_net(obs_hm[T].T, hs[T], cs[T])
        buf.hm[DataKeys.VALUES][T].copy_(values.view(N))
        self.last_launches["collect"] = launches + 4

    # -- step ------------------------------------------------------------------------------------
    def _update_units(self) -> tuple[int, int]:
        hp = self.hparams
        return hp.num_envs * (hp.horizon // hp.seq_len), hp.seq_len

    def _batch_struct(self) -> _lib.RecurrentBatch:  # type: ignore[override]
        rb = _lib.RecurrentBatch()
        rb.b = super()._batch_struct()
        rb.hidden = self.buffer.hm[DataKeys.HIDDEN_STATES].data_ptr()
        rb.cell = self.buffer.hm[DataKeys.CELL_STATES].data_ptr()
        rb.seq_len = self.hparams.seq_len
        return rb

    def _minibatch_launcher(self, batch: Any, M: int) -> tuple[Any, int]:
        lib, model, prec = self._lib, self.policy.model, self.policy.precision
        L = self.hparams.seq_len
        m = model.struct_for(model.flat_params)
        g = model.struct_for(self._grads)
        nbytes = int(lib.rl8_lstm_ppo_workspace(m, M, L, prec))
        if nbytes < 0:
            _lib.check(nbytes, "rl8_lstm_ppo_workspace")
        ws = self._workspace("ppo", nbytes)

        def launch(seqs: Any, begin: int, denom: float, ppo: Any, sums_ptr: Any) -> None:
            rc = lib.rl8_lstm_ppo_minibatch(
                m, g, batch, _lib.ptr(seqs), begin, M, denom, ppo, sums_ptr, prec, _lib.ptr(ws),
                ws.numel(), _lib.stream(),
            )
            _lib.check(rc, "rl8_lstm_ppo_minibatch")

        chunks = max(1, -(-M // max(1, 32768 // L)))
        return launch, chunks * (1 + 12 * L)

    def _reset_buffer(self) -> None:
        """Fresh buffer keeping the final observation AND the final recurrent states
        (src/rl8/algorithms/_recurrent.py:636-646)."""
        buf, T = self.buffer, self.hparams.horizon
        keep = {
            k: buf.hm[k][T].clone()
            for k in (DataKeys.OBS, DataKeys.HIDDEN_STATES, DataKeys.CELL_STATES)
        }
        buf.zero_()
        for k, v in keep.items():
            buf.hm[k][T].copy_(v)

    # -- validate ----------------------------------------------------------------------------------
    def validate(self) -> None:
        """Shape checks on one reset / sample / step (src/rl8/algorithms/_recurrent.py:654-756)."""
        N = self.hparams.num_envs
        obs = self.env.reset()
        self.env.observation_spec.assert_is_in(obs)
        try:
            self.buffer[DataKeys.OBS][:, 0, ...] = obs
        except RuntimeError as e:
            raise AssertionError(
                f"The observation from {type(self.env).__name__}.reset doesn't match the"
                " observation spec shape."
            ) from e
        states = self.policy.init_states(N)
        for k, v in states.items():
            self.policy.state_spec[k].assert_is_in(v)
            self.buffer[DataKeys.STATES][k][:, 0, ...] = v
        sample, sample_states = self.policy.sample(
            {DataKeys.OBS: self.buffer[DataKeys.OBS][:, :1]},
            {k: v[:, :1] for k, v in self.buffer[DataKeys.STATES].items()},
            return_actions=True,
            return_logp=True,
            return_values=True,
        )
        actions = sample[DataKeys.ACTIONS]
        assert actions.ndim >= 2, "Actions must be at least 2D and have shape ``[N, ...]``."
        self.env.action_spec.assert_is_in(actions)
        try:
            self.buffer[DataKeys.ACTIONS][:, 0, ...] = actions
        except RuntimeError as e:
            raise AssertionError(
                "The action sampled from the policy doesn't match the action spec."
            ) from e
        assert sample[DataKeys.LOGP].shape == torch.Size([N, 1]), (
            "Action log probabilities must be 2D and have shape ``[N, 1]``."
        )
        assert sample[DataKeys.VALUES].shape == torch.Size([N, 1]), (
            "Expected value estimates must be 2D and have shape ``[N, 1]``."
        )
        for k, v in sample_states.items():
            self.policy.state_spec[k].assert_is_in(v)
            self.buffer[DataKeys.STATES][k][:, 1, ...] = v
        out = self.env.step(actions)
        obs = out[DataKeys.OBS]
        self.env.observation_spec.assert_is_in(obs)
        try:
            self.buffer[DataKeys.OBS][:, 1, ...] = obs
        except RuntimeError as e:
            raise AssertionError(
                f"The observation from {type(self.env).__name__}.step doesn't match the"
                " observation spec shape."
            ) from e
        assert out[DataKeys.REWARDS].shape == torch.Size([N, 1]), (
            "Rewards must be 2D and have shape ``[N, 1]``."
        )


class RecurrentTrainer(Trainer):
    """Training loop over a :class:`RecurrentAlgorithm` (src/rl8/trainers/_recurrent.py)."""


__all__ = [
    "DefaultContinuousRecurrentModel",
    "DefaultDiscreteRecurrentModel",
    "RecurrentAlgorithm",
    "RecurrentAlgorithmConfig",
    "RecurrentModel",
    "RecurrentPolicy",
    "RecurrentTrainer",
]
