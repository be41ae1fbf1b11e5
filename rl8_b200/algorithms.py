"""Feedforward PPO: ``AlgorithmConfig(...).build(env_cls)``, ``collect()``, ``step()``.

Drop-in for src/rl8/algorithms/_feedforward.py (``AlgorithmConfig`` 29-179, ``Algorithm``
182-697) with the hot path on sm_100a kernels:

* ``collect()`` -> one ``rl8_collect`` call (T fused policy-forward / sample / env-step /
  buffer-write steps + a batched value pass) and one ``rl8_collect_stats`` reduction; ONE
  device->host readback per call (the reference does nine).
* ``step()`` -> ``rl8_gae_scan`` / ``rl8_gae_normalize``, then per minibatch
  ``rl8_ppo_minibatch`` (forward + clipped losses + hand-derived backward) and
  ``rl8_clip_adam``; loss statistics stay on the device until the end of the call.

Multi-GPU: environments shard across ranks -- every rank builds its own ``Algorithm`` with
its share of ``num_envs`` -- and, when ``torch.distributed`` is initialised, gradients are
summed with one NCCL all-reduce per optimizer step and the global statistics (advantage
moments, reward scale, loss sums) with one small all-reduce per phase.
"""

from __future__ import annotations

import ctypes
import math
import os
import time
from dataclasses import asdict, dataclass
from typing import Any, Literal

import torch
import torch.distributed as dist
import torch.optim as optim

from . import _lib, parallel
from .buffer import RolloutBuffer
from .data import AlgorithmHparams, AlgorithmState, CollectStats, DataKeys, Device, MemoryStats, StepStats
from .distributions import Distribution
from .env import Env, KernelEnv
from .models import Model
from .policies import Policy
from .schedulers import EntropyScheduler, LRScheduler, ScheduleKind
from .specs import Categorical, Composite, Unbounded


@dataclass
class AlgorithmConfig:
    """Configuration of a feedforward PPO algorithm (same fields and defaults as the
    reference, src/rl8/algorithms/_feedforward.py:33-173)."""

    model: None | Model = None
    model_cls: None | type[Model] = None
    model_config: None | dict[str, Any] = None
    distribution_cls: None | type[Distribution] = None
    #: Transitions per env per ``collect`` (buffer is ``[num_envs, horizon + 1]``).
    horizon: int = 32
    #: ``collect`` calls between env resets (negative: reset once, ever).
    horizons_per_env_reset: int = 1
    #: Parallel environments ON THIS RANK.
    num_envs: int = 8192
    optimizer_cls: type[optim.Optimizer] = optim.Adam
    optimizer_config: None | dict[str, Any] = None
    accumulate_grads: bool = False
    #: Mixed precision: ``False`` -> fp32 CUDA-core GEMMs (bit-comparable with the
    #: reference's fp32 path), ``True`` -> bf16 tcgen05 GEMMs with fp32 accumulation.
    enable_amp: bool = False
    lr_schedule: None | list[tuple[int, float]] = None
    lr_schedule_kind: ScheduleKind = "step"
    entropy_coeff: float = 0.0
    entropy_coeff_schedule: None | list[tuple[int, float]] = None
    entropy_coeff_schedule_kind: ScheduleKind = "step"
    gae_lambda: float = 0.95
    gamma: float = 0.95
    #: ``None``: the whole buffer is one batch.
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

    def build(self, env_cls: Any) -> "Algorithm":
        """Build and validate an :class:`Algorithm`."""
        algo = Algorithm(env_cls, config=self)
        algo.validate()
        return algo


class _NoopGradScaler:
    """bf16 needs no loss scaling; kept so ``algo.grad_scaler`` exists like the reference's."""

    def __init__(self, enabled: bool) -> None:
        self._enabled = enabled

    def is_enabled(self) -> bool:
        return self._enabled

    def get_scale(self) -> float:
        return 1.0


class FusedAdam(optim.Adam):
    """``torch.optim.Adam``'s param groups and ``state_dict`` over the flat moment buffers that ``rl8_clip_adam``
    updates (src/rl8/algorithms/_feedforward.py:257-260): LR schedules mutate ``param_groups`` as in the reference, and
    ``state_dict()`` / ``load_state_dict()`` carry the moments and the step count, so a save / resume keeps Adam's
    moments and bias correction.  ``step()`` is never called: the update is the fused kernel."""

    #: optimizer updates applied so far (the 1-based ``step`` of Adam's bias correction)
    update_count = 0

    def bind_flat(self, model: Any, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor) -> None:
        self._flat = (model, exp_avg, exp_avg_sq)
        self._publish()

    def _publish(self) -> None:
        model, m, v = self._flat
        vm, vv = model.named_flat_views(m), model.named_flat_views(v)
        for n, p in model.named_parameters():
            self.state[p] = {"step": torch.tensor(float(self.update_count)), "exp_avg": vm[n], "exp_avg_sq": vv[n]}

    def state_dict(self) -> dict[str, Any]:
        for st in self.state.values():
            st["step"] = torch.tensor(float(self.update_count))
        return super().state_dict()

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        super().load_state_dict(state_dict)  # replaces the per-parameter state with copies
        model, m, v = self._flat
        vm, vv = model.named_flat_views(m), model.named_flat_views(v)
        steps = 0
        for n, p in model.named_parameters():
            st = self.state.get(p)
            if st:
                vm[n].copy_(st["exp_avg"])
                vv[n].copy_(st["exp_avg_sq"])
                steps = int(st["step"])
            else:
                vm[n].zero_()
                vv[n].zero_()
        self.update_count = steps
        self._publish()


def make_flat_optimizer(model: Any, grads: torch.Tensor, optimizer_cls: Any,
                        optimizer_config: dict[str, Any]) -> tuple[optim.Optimizer, None | tuple[torch.Tensor, ...]]:
    """The optimizer of a default (flat-parameter) model.  Plain Adam -- the reference default -- is the fused
    ``rl8_clip_adam`` behind a :class:`FusedAdam`; returns ``(optimizer, (exp_avg, exp_avg_sq))``.  Any other
    ``optimizer_cls`` / Adam option (src/rl8/algorithms/_feedforward.py:257-260) is the caller's torch optimizer
    stepping on the model's parameters with ``.grad`` bound to views of the flat gradient buffer the update kernels
    fill (``rl8_clip_grads`` does the clipping): returns ``(optimizer, None)``."""
    plain = {k: v for k, v in optimizer_config.items() if k not in ("lr", "betas", "eps") and v}
    if optimizer_cls is optim.Adam and not plain:
        opt = FusedAdam(model.parameters(), **optimizer_config)
        m, v = torch.zeros_like(grads), torch.zeros_like(grads)
        opt.bind_flat(model, m, v)
        return opt, (m, v)
    opt = optimizer_cls(model.parameters(), **optimizer_config)
    opt.update_count = 0
    views = model.named_flat_views(grads)
    for n, p in model.named_parameters():
        p.grad = views[n]
    return opt, None


_mem_cache: dict[int, tuple[int, int, int]] = {}  # device -> (allocator bytes reserved, free, total)


def memory_stats() -> MemoryStats:
    """Device memory like the reference's ``torch.cuda.mem_get_info()``
    (src/rl8/_utils.py:102-115).  ``cudaMemGetInfo`` takes a driver lock and sporadically
    stalls for milliseconds, and ``Trainer.step`` calls this every iteration, so the driver
    is only asked again when this process's caching allocator has grown or shrunk since the
    last query (the only way this process changes the device's free memory)."""
    dev = torch.cuda.current_device()
    reserved = torch.cuda.memory_reserved(dev)
    hit = _mem_cache.get(dev)
    if hit is None or hit[0] != reserved:
        free, total = torch.cuda.mem_get_info(dev)
        hit = _mem_cache[dev] = (reserved, free, total)
    _, free, total = hit
    return {"memory/free": free, "memory/total": total, "memory/percent": 100 * (total - free) / total}


_world = parallel.world_size


def _index_views(views: Any, idx: Any) -> Any:
    """Rows ``idx`` of every tensor leaf of a (nested) views mapping."""
    if isinstance(views, torch.Tensor):
        return views[idx]
    return {k: _index_views(v, idx) for k, v in views.items()}


class LazyStats(dict):
    """``CollectStats`` whose device-computed entries are read back on first access.

    Host-known entries (``env/steps``, ``env/resets``, ``profiling/collect_ms``) are plain items; touching any other
    key, iterating, printing or unpacking (``{**stats}``) waits -- once -- for the asynchronous device->host copy
    enqueued by ``collect()`` and fills the rest in."""

    _EAGER = ("env/resets", "env/steps", "profiling/collect_ms")

    def __init__(self, read_back: Any) -> None:
        super().__init__()
        self._read_back = read_back

    def set_eager(self, key: str, value: Any) -> None:
        dict.__setitem__(self, key, value)

    def materialize(self) -> None:
        if self._read_back is not None:
            read_back, self._read_back = self._read_back, None
            for k, v in read_back().items():
                if not dict.__contains__(self, k):
                    dict.__setitem__(self, k, v)

    def __getitem__(self, key: str) -> Any:
        if not dict.__contains__(self, key):
            self.materialize()
        return dict.__getitem__(self, key)

    def get(self, key: str, default: Any = None) -> Any:
        if not dict.__contains__(self, key):
            self.materialize()
        return dict.get(self, key, default)

    def __contains__(self, key: object) -> bool:
        if dict.__contains__(self, key):
            return True
        self.materialize()
        return dict.__contains__(self, key)

    def _full(name: str):  # noqa: ANN202, N805
        def method(self, *a: Any, **kw: Any) -> Any:
            self.materialize()
            return getattr(dict, name)(self, *a, **kw)

        method.__name__ = name
        return method

    for _n in ("keys", "values", "items", "__iter__", "__len__", "__repr__", "__eq__", "__ne__", "copy", "__or__",
               "__ror__", "__reduce_ex__", "__reduce__"):
        locals()[_n] = _full(_n)
    del _n, _full
    __hash__ = None  # type: ignore[assignment]


class _RunningMean:
    def __init__(self) -> None:
        self.avg, self.n = 0.0, 0

    def update(self, v: float) -> None:
        self.avg = (v + self.n * self.avg) / (self.n + 1)
        self.n += 1


class Algorithm:
    """PPO over a tensor-batched environment with the rollout and the update on CUDA kernels."""

    def __init__(self, env_cls: Any, /, config: None | AlgorithmConfig = None) -> None:
        config = config or AlgorithmConfig()
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
        self.policy = Policy(
            self.env.observation_spec,
            self.env.action_spec,
            model=config.model,
            model_cls=config.model_cls,
            model_config=config.model_config,
            distribution_cls=config.distribution_cls,
            device=device,
        )
        self.policy.precision = _lib.precision_for(config.enable_amp)
        self.buffer_spec = Composite(
            {
                DataKeys.OBS: self.env.observation_spec,
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
        self._fused_model = self.policy.fused
        optimizer_config = dict(config.optimizer_config or {"lr": 1e-3})
        if self._fused_model:
            # Plain Adam: param groups in a torch optimizer object (lr schedules mutate them), the update itself in
            # rl8_clip_adam on the flat buffers.  Other optimizer classes / options: torch steps on the flat views.
            self._grads = torch.zeros_like(self.policy.model.flat_params)
            self.optimizer, moments = make_flat_optimizer(self.policy.model, self._grads, config.optimizer_cls,
                                                          optimizer_config)
            self._exp_avg, self._exp_avg_sq = moments if moments is not None else (None, None)
        else:
            # user-defined torch model: its parameters belong to torch, any optimizer class works
            self.optimizer = config.optimizer_cls(self.policy.model.parameters(), **optimizer_config)
            self.optimizer.update_count = 0
        # Multi-GPU: replicas start from rank 0's parameters whatever each rank's RNG state was
        # (seed env resets / sampling noise per rank; the model is made identical here).
        parallel.sync_replicas(self.policy.model)
        self._grad_norm = torch.zeros(1, device=device)
        self.lr_scheduler = LRScheduler(
            self.optimizer, schedule=config.lr_schedule, kind=config.lr_schedule_kind
        )
        self.entropy_scheduler = EntropyScheduler(
            config.entropy_coeff,
            schedule=config.entropy_coeff_schedule,
            kind=config.entropy_coeff_schedule_kind,
        )
        sgd_minibatch_size = config.sgd_minibatch_size or num_envs * horizon
        self.hparams = AlgorithmHparams(
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
            sgd_minibatch_size=sgd_minibatch_size,
            shuffle_minibatches=config.shuffle_minibatches,
            target_kl_div=config.target_kl_div,
            vf_clip_param=config.vf_clip_param,
            vf_coeff=config.vf_coeff,
        ).validate()
        self.state = AlgorithmState()
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
        #: [reward scale, f32(scale + 1e-8)] of the last collect(), device-resident for step()
        self._scale_dev = torch.ones(2, dtype=torch.float32, device=device)
        max_updates = self.hparams.num_sgd_iters * self.hparams.num_minibatches
        self._loss_sums = torch.zeros(max_updates, 5, dtype=torch.float64, device=device)
        #: test / debugging hook: called with the named (unclipped) gradients of every update
        self._on_grads: Any = None
        #: number of kernels of this library launched by the last collect() / step()
        self.last_launches = {"collect": 0, "step": 0}
        # CUDA-graph replay of GAE + the update epochs (see _graph_key): the per-step scalars live on the device
        self._update_graphs: dict[tuple, tuple[torch.cuda.CUDAGraph, int, list[bool], int]] = {}
        self._graph_eligible_calls = 0
        self._lr_dev = torch.zeros(1, dtype=torch.float64, device=device)
        self._steps_dev = torch.zeros(1, dtype=torch.int64, device=device)
        self._lr_mirror: None | float = None
        self._steps_mirror: None | int = None
        self._adam_scratch = torch.zeros(16, dtype=torch.float32, device=device)

    # ------------------------------------------------------------------------------------
    @property
    def _opt_steps(self) -> int:
        """Optimizer updates applied so far (kept on the optimizer object: it travels with its state_dict)."""
        return self.optimizer.update_count  # type: ignore[attr-defined]

    @_opt_steps.setter
    def _opt_steps(self, value: int) -> None:
        self.optimizer.update_count = value  # type: ignore[attr-defined]

    @property
    def horizons_per_env_reset(self) -> int:
        return self.hparams.horizons_per_env_reset

    def memory_stats(self) -> MemoryStats:
        return memory_stats()

    @property
    def params(self) -> dict[str, Any]:
        return {
            "env_cls": type(self.env).__name__,
            "model_cls": type(self.policy.model).__name__,
            "distribution_cls": self.policy.distribution_cls.__name__,
            "optimizer_cls": type(self.optimizer).__name__,
            "entropy_coeff": self.entropy_scheduler.coeff,
            **asdict(self.hparams),
        }

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------
    # collect
    # ------------------------------------------------------------------------------------
    def collect(
        self, *, env_config: None | dict[str, Any] = None, deterministic: bool = False
    ) -> CollectStats:
        """Roll the policy out for ``horizon`` steps into the buffer
        (src/rl8/algorithms/_feedforward.py:301-441)."""
        start = time.perf_counter_ns()
        hp, buf = self.hparams, self.buffer
        N, T = hp.num_envs, hp.horizon
        obs_hm = buf.hm[DataKeys.OBS]
        rdr_hm = buf.hm.get(DataKeys.REVERSED_DISCOUNTED_RETURNS)
        env_was_reset = False
        carry = (self.state.horizons and hp.horizons_per_env_reset < 0) or (
            self.state.horizons % hp.horizons_per_env_reset
        )
        if carry:
            obs_hm[0].copy_(obs_hm[T])
            if rdr_hm is not None:
                rdr_hm[0].copy_(rdr_hm[T])
        else:
            obs0 = self.env.reset(config=env_config)
            obs_hm[0].copy_(obs0.reshape(N, -1).T)
            env_was_reset = True
            if rdr_hm is not None:
                rdr_hm[0].zero_()
        self._pre_collect()

        P = self._head_width()
        dist_cls = self.policy.distribution_cls
        noise = None
        if not deterministic:
            noise = dist_cls.draw_noise(T, N, P, self.device).contiguous()
            assert noise.dtype == torch.float32 and noise.is_cuda
        if not self._fused_model:
            self._collect_generic_model(noise, deterministic)
        elif self._fused_env:
            self._collect_fused(noise, deterministic)
        else:
            self._collect_generic(noise, deterministic)

        # statistics + reward scale: one reduction kernel (+ one collective); NOTHING is read back here -- the reward
        # scale stays on the device for step(), the statistics are copied to pinned memory asynchronously and the
        # returned mapping waits for them on first access (Trainer.step touches them after step() is enqueued)
        acc = self._stats_acc
        acc.copy_(self._stats_init)
        t0 = self._stats_reward_t0
        rc = self._lib.rl8_collect_stats_from(
            _lib.ptr(buf.hm[DataKeys.REWARDS]), _lib.ptr(rdr_hm), N, T, t0, _lib.ptr(acc),
            _lib.stream(),
        )
        _lib.check(rc, "rl8_collect_stats_from")
        world = _world()
        parallel.reduce_collect_acc_(acc)
        n_r, n_R = float(N * T * world), float(N * world)
        rc = self._lib.rl8_reward_scale(_lib.ptr(acc), n_r, int(hp.normalize_rewards), _lib.ptr(self._scale_dev),
                                        _lib.stream())
        _lib.check(rc, "rl8_reward_scale")
        self.last_launches["collect"] += 2
        host = torch.empty(18, dtype=torch.float64, pin_memory=True)  # torch's caching host allocator: no cudaHostAlloc
        host[:16].copy_(acc, non_blocking=True)
        host[16:18].copy_(self._scale_dev.double(), non_blocking=True)
        ready = torch.cuda.Event()
        ready.record()
        state = self.state
        n_reward = float(N * (T - t0) * world)

        def read_back() -> dict[str, float]:
            ready.synchronize()  # the one device->host wait of collect(), deferred to the first reader
            a = host.tolist()
            mean_std = parallel.mean_std
            r_mean, r_std = mean_std(a[0], a[1], n_reward)
            R_mean, R_std = mean_std(a[2], a[3], n_R)
            if state._pending_stats is pending:  # still the latest collect(): its scale is the state's
                state._reward_scale = a[16]
                state._pending_stats = None
            return {
                "returns/min": a[8], "returns/max": a[9], "returns/mean": R_mean, "returns/std": R_std,
                "rewards/min": a[6], "rewards/max": a[7], "rewards/mean": r_mean, "rewards/std": r_std,
            }

        stats = LazyStats(read_back)
        pending = stats.materialize
        state._pending_stats = pending
        state._scale_on_device = True
        self.state.horizons += 1
        self.state.buffered = True
        self._post_collect()
        # global counts (all ranks), like every other statistic of the call
        stats.set_eager("env/resets", hp.num_envs * world * int(env_was_reset))
        stats.set_eager("env/steps", hp.num_envs * world * hp.horizon)
        stats.set_eager("profiling/collect_ms", (time.perf_counter_ns() - start) / 1e6)
        return stats  # type: ignore[return-value]

    #: first reward slot of the collect statistics (the recurrent algorithm uses 1)
    _stats_reward_t0 = 0

    def _pre_collect(self) -> None:
        """Hook between the env reset / carry-over and the rollout."""

    def _post_collect(self) -> None:
        """Hook after the rollout (counters)."""

    def _rollout_struct(self, noise: None | torch.Tensor, deterministic: bool) -> _lib.Rollout:
        hp, buf, env = self.hparams, self.buffer, self.env
        assert isinstance(env, KernelEnv)
        ro = _lib.Rollout()
        ro.env_kind = env.rl8_kind
        ro.dist_kind = self.policy.distribution_cls.rl8_kind
        ro.T, ro.N = hp.horizon, hp.num_envs
        ro.deterministic = int(deterministic)
        ro.gamma = hp.gamma
        ro.normalize_rewards = int(hp.normalize_rewards)
        ro.env_cfg = env.rl8_cfg()
        ro.env_state = env.state.data_ptr()
        ro.obs = buf.hm[DataKeys.OBS].data_ptr()
        ro.actions = buf.hm[DataKeys.ACTIONS].data_ptr()
        ro.logp = buf.hm[DataKeys.LOGP].data_ptr()
        ro.values = buf.hm[DataKeys.VALUES].data_ptr()
        ro.rewards = buf.hm[DataKeys.REWARDS].data_ptr()
        rdr = buf.hm.get(DataKeys.REVERSED_DISCOUNTED_RETURNS)
        ro.rdr = rdr.data_ptr() if rdr is not None else None
        ro.noise = noise.data_ptr() if noise is not None else None
        return ro

    def _collect_fused(self, noise: None | torch.Tensor, deterministic: bool) -> None:
        hp = self.hparams
        model = self.policy.model
        m = model.struct_for(model.flat_params)
        nbytes = int(self._lib.rl8_collect_workspace(m, hp.num_envs, hp.horizon, self.policy.precision))
        if nbytes < 0:
            _lib.check(nbytes, "rl8_collect_workspace")
        ws = self._workspace("collect", nbytes)
        ro = self._rollout_struct(noise, deterministic)
        rc = self._lib.rl8_collect(m, ro, self.policy.precision, _lib.ptr(ws), ws.numel(), _lib.stream())
        _lib.check(rc, "rl8_collect")
        T = hp.horizon
        self.last_launches["collect"] = {
            _lib.PREC_FP32: 4 * T + 3 * (T + 1),  # layer 1, SGEMM, head, tail per step; 3 per value slab
            _lib.PREC_BF16: 4,                    # 2 x W2 packing, rollout kernel, value pass
            # W2 piece images (one launch), max |obs| of slab 0 (two launches), (split forward + tail) per step, value pass
            _lib.PREC_FP32_TC: 1 + 2 + 2 * T + 1,
        }[self.policy.precision]

    def _head_width(self) -> int:
        """Policy-head outputs the sampling kernels see: A logits, or 2 = {mean, log_std}."""
        if self._fused_model:
            return self.policy.model.head_width
        spec = self.env.action_spec
        return int(spec.space.n) if isinstance(spec, Categorical) else 2

    def _collect_generic_model(self, noise: None | torch.Tensor, deterministic: bool) -> None:
        """Rollout with a user-defined torch model (rl8_b200.models.GenericModel): the model's forward
        runs through torch every step; sampling, log-probabilities, the env step (bundled envs) and the
        buffer writes stay on this library's kernels."""
        hp, buf = self.hparams, self.buffer
        N, T = hp.num_envs, hp.horizon
        obs_hm, act_hm = buf.hm[DataKeys.OBS], buf.hm[DataKeys.ACTIONS]
        rdr_hm = buf.hm.get(DataKeys.REVERSED_DISCOUNTED_RETURNS)
        model, kind = self.policy.model, self.policy.distribution_cls.rl8_kind
        obs_em = buf[DataKeys.OBS]  # [N, T+1, D] view
        launches = 0
        model.eval()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=hp.enable_amp):
            for t in range(T + 1):
                views = model.apply_view_requirements({DataKeys.OBS: obs_em[:, : t + 1]}, kind="last")
                features = model(views)
                buf.hm[DataKeys.VALUES][t].copy_(model.value_function().reshape(N))
                if t == T:
                    break
                packed = self.policy.distribution_cls(features, model)._packed()
                nz = None if noise is None else noise[t]
                rc = self._lib.rl8_dist_sample(
                    kind, _lib.ptr(packed), packed.shape[1], _lib.ptr(nz), int(deterministic),
                    _lib.ptr(act_hm[t]), _lib.ptr(buf.hm[DataKeys.LOGP][t]), N, _lib.stream(),
                )
                _lib.check(rc, "rl8_dist_sample")
                out = self.env.step(act_hm[t].view(N, 1))
                rewards = out[DataKeys.REWARDS].reshape(N)
                if rdr_hm is not None:
                    torch.add(rewards, rdr_hm[t], alpha=hp.gamma, out=rdr_hm[t + 1])
                buf.hm[DataKeys.REWARDS][t].copy_(rewards)
                obs_hm[t + 1].copy_(out[DataKeys.OBS].reshape(N, -1).T)
                launches += 2  # rl8_dist_sample + the env's step kernel
        self.last_launches["collect"] = launches

    def _collect_generic(self, noise: None | torch.Tensor, deterministic: bool) -> None:
        """Rollout with a user-defined (Python / torch) environment: the policy forward,
        sampling and log-probabilities still run on this library's kernels; ``env.step`` is
        the user's code (the reference's Env plug-in point)."""
        hp, buf = self.hparams, self.buffer
        N, T = hp.num_envs, hp.horizon
        obs_hm, act_hm = buf.hm[DataKeys.OBS], buf.hm[DataKeys.ACTIONS]
        rdr_hm = buf.hm.get(DataKeys.REVERSED_DISCOUNTED_RETURNS)
        P = self.policy.model.head_width
        kind = self.policy.distribution_cls.rl8_kind
        launches = 0
        for t in range(T):
            head = self.policy.forward_net(0, obs_hm[t].T)
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
            launches += 4
        for t in range(T + 1):
            self.policy.forward_net(1, obs_hm[t].T, out=buf.hm[DataKeys.VALUES][t].view(N, 1))
            launches += 3
        self.last_launches["collect"] = launches

    # ------------------------------------------------------------------------------------
    # step
    # ------------------------------------------------------------------------------------
    def step(self) -> StepStats:
        """GAE, then ``num_sgd_iters`` PPO epochs over the buffer
        (src/rl8/algorithms/_feedforward.py:443-615)."""
        if not self.state.buffered:
            raise RuntimeError(
                f"{type(self).__name__} is not buffered. Call `collect` once prior to `step`."
            )
        start = time.perf_counter_ns()
        hp = self.hparams
        if not self._fused_model:
            return self._step_generic_model(start, self._enqueue_gae())

        entropy_coeff = self.entropy_scheduler.coeff
        if entropy_coeff != 0 and self.policy.distribution_cls.rl8_kind == _lib.DIST_SQUASHED_NORMAL:
            self.policy.distribution_cls({}, self.policy.model).entropy()  # raises like the reference
        key = self._graph_key(entropy_coeff)
        if key is None:
            launches, k, applied = self._enqueue_update(entropy_coeff, dev_scalars=False)
        else:
            launches, k, applied = self._replay_update(key, entropy_coeff)
        accum = hp.num_minibatches if hp.accumulate_grads else 1
        return self._finish_step(start, launches, k, applied, accum, entropy_coeff)

    # -- CUDA-graph replay of the update ---------------------------------------------------------------------------
    #: the feedforward update is a fixed launch sequence; the recurrent one keeps its own host loop
    _graph_capable = True

    def _graph_key(self, entropy_coeff: float) -> None | tuple:
        """Key of the CUDA graph that can stand for this call's GAE + update epochs, or None when the sequence
        depends on the host: early stopping reads a KL per minibatch, shuffled minibatches draw permutations, a
        caller-assigned reward scale / gradient hook / non-fused optimizer are host values, NCCL calls sit between
        the kernels when world > 1.  Everything that changes between calls of an eligible configuration lives in
        device memory: the reward scale, the learning rate and the optimizer step count (rl8_clip_adam_dev)."""
        hp = self.hparams
        if (not self._graph_capable or os.environ.get("RL8_CUDA_GRAPH", "1") == "0" or _world() > 1
                or self._exp_avg is None or hp.target_kl_div is not None or self._on_grads is not None
                or (hp.shuffle_minibatches and hp.num_minibatches > 1) or not self.state._scale_on_device):
            return None
        pg = self.optimizer.param_groups[0]
        key = (self.policy.precision, float(entropy_coeff), tuple(pg.get("betas", (0.9, 0.999))),
               float(pg.get("eps", 1e-8)), hp.clip_param, hp.dual_clip_param, hp.vf_clip_param, hp.vf_coeff,
               hp.max_grad_norm, hp.gamma, hp.gae_lambda, hp.normalize_advantages, hp.num_sgd_iters,
               hp.sgd_minibatch_size, hp.accumulate_grads, self.buffer.hm[DataKeys.OBS].data_ptr(),
               self._ws["ppo"].data_ptr() if "ppo" in self._ws else 0,  # a re-sized workspace invalidates the capture
               self.policy.model.flat_params.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if key not in self._update_graphs and len(self._update_graphs) >= 4:
            return None  # e.g. an entropy schedule that changes every step: not worth a capture per value
        return key

    def _replay_update(self, key: tuple, entropy_coeff: float) -> tuple[int, int, list[bool]]:
        pg = self.optimizer.param_groups[0]
        lr = float(pg["lr"])
        if self._lr_mirror != lr:
            self._lr_dev.fill_(lr)
            self._lr_mirror = lr
        if self._steps_mirror != self._opt_steps:  # first use, or a loaded optimizer state
            self._steps_dev.fill_(self._opt_steps)
            self._steps_mirror = self._opt_steps
        entry = self._update_graphs.get(key)
        if entry is None:
            self._graph_eligible_calls += 1
            if self._graph_eligible_calls < 2:
                # the first eligible call runs eagerly: it sizes the workspaces and loads the kernels
                res = self._enqueue_update(entropy_coeff, dev_scalars=True)
            else:
                graph = torch.cuda.CUDAGraph()
                torch.cuda.current_stream().synchronize()
                with torch.cuda.graph(graph, stream=self._capture_stream()):
                    res = self._enqueue_update(entropy_coeff, dev_scalars=True)
                self._update_graphs[key] = (graph, *res)
                graph.replay()
        else:
            entry[0].replay()
            res = entry[1:]
        launches, k, applied = res
        n_opt = sum(applied)
        self._opt_steps += n_opt
        self._steps_mirror = self._opt_steps
        return launches, k, list(applied)

    def _capture_stream(self) -> torch.cuda.Stream:
        cs = getattr(self, "_capture_stream_obj", None)
        if cs is None:
            cs = self._capture_stream_obj = torch.cuda.Stream()
        return cs

    def _enqueue_gae(self) -> int:
        """GAE over the buffer (the reward scale of the last collect() is read from device memory: no host round
        trip).  Returns the number of kernels launched."""
        hp, buf, lib = self.hparams, self.buffer, self._lib
        N, T = hp.num_envs, hp.horizon
        st = _lib.stream()
        launches = 0
        self._moments.zero_()
        if self.state._scale_on_device:
            rc = lib.rl8_gae_scan_dev(
                _lib.ptr(buf.hm[DataKeys.REWARDS]), _lib.ptr(buf.hm[DataKeys.VALUES]),
                _lib.ptr(buf.hm[DataKeys.ADVANTAGES]), _lib.ptr(buf.hm[DataKeys.RETURNS]),
                N, T, 1, N, hp.gamma, hp.gae_lambda, _lib.ptr(self._scale_dev),
                0,  # the buffer (rewards included) is re-zeroed at the end of step(): no scaled-reward write-back
                _lib.ptr(self._moments), st,
            )
        else:  # a reward scale assigned by the caller
            rc = lib.rl8_gae_scan(
                _lib.ptr(buf.hm[DataKeys.REWARDS]), _lib.ptr(buf.hm[DataKeys.VALUES]),
                _lib.ptr(buf.hm[DataKeys.ADVANTAGES]), _lib.ptr(buf.hm[DataKeys.RETURNS]),
                N, T, 1, N, hp.gamma, hp.gae_lambda, self.state.reward_scale,
                _lib.ptr(self._moments), st,
            )
        _lib.check(rc, "rl8_gae_scan")
        launches += 1
        if hp.normalize_advantages:
            if _world() > 1:
                dist.all_reduce(self._moments)
            rc = lib.rl8_gae_normalize(
                _lib.ptr(buf.hm[DataKeys.ADVANTAGES]), N, T, 1, N, _lib.ptr(self._moments), st
            )
            _lib.check(rc, "rl8_gae_normalize")
            launches += 1
        return launches

    def _enqueue_update(self, entropy_coeff: float, dev_scalars: bool) -> tuple[int, int, list[bool]]:
        """GAE + the PPO epochs on the current stream.  ``dev_scalars``: the learning rate and the optimizer step
        count are read from (and the count advanced in) device memory, so the launch sequence can be captured in a
        CUDA graph; the caller then owns the host-side count.  Returns ``(kernel launches, minibatches processed,
        step-boundary flag per minibatch)``."""
        hp, lib = self.hparams, self._lib
        world = _world()
        st = _lib.stream()
        launches = self._enqueue_gae()

        model = self.policy.model
        M = hp.sgd_minibatch_size
        batch = self._batch_struct()
        launch_minibatch, mb_launches = self._minibatch_launcher(batch, M)
        units, rows_per_unit = self._update_units()

        accum = hp.num_minibatches if hp.accumulate_grads else 1
        ppo = _lib.PpoHparams(
            hp.clip_param, hp.dual_clip_param or 0.0, entropy_coeff, hp.vf_clip_param,
            hp.vf_coeff, 1.0 / accum,
        )
        pg = self.optimizer.param_groups[0]
        sums = self._loss_sums
        sums.zero_()
        self._grads.zero_()
        k = 0  # minibatches processed
        applied: list[bool] = []  # step boundary flags per minibatch
        stop_early = False
        for _ in range(hp.num_sgd_iters):
            # A single minibatch is the whole buffer: its loss is a sum over all rows, so the
            # reference's permutation (src/rl8/_utils.py:211-218) only reorders that sum.
            shuffle = hp.shuffle_minibatches and hp.num_minibatches > 1
            perm = torch.randperm(units, device=self.device) if shuffle else None
            for i in range(hp.num_minibatches):
                step_this_batch = (i + 1) % accum == 0
                rows = None if perm is None else perm[i * M : (i + 1) * M]
                launch_minibatch(
                    rows, i * M, float(M * rows_per_unit * world), ppo,
                    ctypes.c_void_p(sums.data_ptr() + 40 * k),
                )
                launches += mb_launches
                applied.append(step_this_batch)
                k += 1
                if hp.target_kl_div is not None:
                    # Early stopping needs this minibatch's KL on the host (the reference
                    # syncs on every minibatch regardless); the triggering minibatch is not
                    # applied (:576-585).
                    row = sums[k - 1].clone()
                    if world > 1:
                        dist.all_reduce(row)
                    kl = float(row[3] / row[4])
                    if kl > 1.5 * hp.target_kl_div:
                        stop_early = True
                        self._grads.zero_()
                        break
                if step_this_batch:
                    if world > 1:
                        dist.all_reduce(self._grads)  # grads already carry 1 / (M * world)
                    if self._on_grads is not None:
                        self._on_grads(model.named_flat_views(self._grads))
                    betas = pg.get("betas", (0.9, 0.999))
                    if dev_scalars:
                        rc = lib.rl8_clip_adam_dev(
                            _lib.ptr(model.flat_params), _lib.ptr(self._grads), _lib.ptr(self._exp_avg),
                            _lib.ptr(self._exp_avg_sq), self._grads.numel(), hp.max_grad_norm,
                            _lib.ptr(self._lr_dev), betas[0], betas[1], pg.get("eps", 1e-8),
                            _lib.ptr(self._steps_dev), _lib.ptr(self._adam_scratch), st,
                        )
                        _lib.check(rc, "rl8_clip_adam_dev")
                        launches += 1
                    elif self._exp_avg is not None:
                        self._opt_steps += 1
                        rc = lib.rl8_clip_adam(
                            _lib.ptr(model.flat_params), _lib.ptr(self._grads), _lib.ptr(self._exp_avg),
                            _lib.ptr(self._exp_avg_sq), self._grads.numel(), hp.max_grad_norm,
                            pg["lr"], betas[0], betas[1], pg.get("eps", 1e-8), self._opt_steps,
                            _lib.ptr(self._grad_norm), st,
                        )
                        _lib.check(rc, "rl8_clip_adam")
                    else:  # optimizer_cls of the caller: clip here, torch steps on the flat views (.grad = self._grads)
                        self._opt_steps += 1
                        rc = lib.rl8_clip_grads(_lib.ptr(self._grads), self._grads.numel(), hp.max_grad_norm,
                                                _lib.ptr(self._grad_norm), st)
                        _lib.check(rc, "rl8_clip_grads")
                        self.optimizer.step()
                    launches += 2
                    self._grads.zero_()
            if stop_early:
                break
        return launches, k, applied

    def _finish_step(self, start: int, launches: int, k: int, applied: list[bool], accum: int,
                     entropy_coeff: float) -> StepStats:
        """Statistics of the update (one readback), schedulers, buffer reset."""
        hp = self.hparams
        world = _world()
        sums = self._loss_sums
        # -- statistics: one readback ----------------------------------------------------------
        used = sums[:k]
        if world > 1 and k:
            used = used.clone()
            dist.all_reduce(used)
        rows_host = used.tolist()
        keys = ("losses/entropy", "losses/policy", "losses/vf", "losses/total", "monitors/kl_div")
        means = {key: _RunningMean() for key in keys}
        coeff_e, coeff_v = _RunningMean(), _RunningMean()
        run = dict.fromkeys(keys, 0.0)
        for (s_ent, s_pol, s_vf, s_kl, cnt), reduce in zip(rows_host, applied):
            ent, pol, vf, kl = s_ent / cnt, s_pol / cnt, s_vf / cnt, s_kl / cnt
            total = hp.vf_coeff * vf - pol - (entropy_coeff * ent if entropy_coeff != 0 else 0.0)
            run["losses/entropy"] += ent / accum
            run["losses/policy"] += pol / accum
            run["losses/vf"] += vf / accum
            run["losses/total"] += total / accum
            run["monitors/kl_div"] += kl / accum
            coeff_e.update(entropy_coeff)
            coeff_v.update(hp.vf_coeff)
            if reduce:
                for key in keys:
                    means[key].update(run[key])
                    run[key] = 0.0

        # schedules are defined in env transitions of the whole job (src/rl8/schedulers.py:121-232): global env count
        self.lr_scheduler.step(hp.num_envs * world * self.state.horizons)
        self.entropy_scheduler.step(hp.num_envs * world * self.state.horizons)

        self._reset_buffer()
        self.state.buffered = False
        self.last_launches["step"] = launches

        stats: StepStats = {
            "coefficients/entropy": coeff_e.avg,
            "coefficients/vf": coeff_v.avg,
            **{key: means[key].avg for key in keys},  # type: ignore[typeddict-item]
        }
        torch.cuda.current_stream().synchronize()
        stats["profiling/step_ms"] = (time.perf_counter_ns() - start) / 1e6
        return stats

    def _step_generic_model(self, start: int, launches: int) -> StepStats:
        """The PPO epochs for a user-defined torch model (src/rl8/algorithms/_feedforward.py:469-600):
        per minibatch ``model(views)`` through torch, ``rl8_ppo_losses_direct`` for the clipped losses and
        their gradients w.r.t. the model outputs, ``torch.autograd.backward`` through the model, then
        ``clip_grad_norm_`` and the user's optimizer."""
        hp, buf, lib = self.hparams, self.buffer, self._lib
        N, T = hp.num_envs, hp.horizon
        world = _world()
        model, dist_cls = self.policy.model, self.policy.distribution_cls
        kind = dist_cls.rl8_kind
        if model.drop_size > 0:
            # A "rolling_window" requirement yields N * (T - drop_size) rows while actions, log-probabilities,
            # advantages and returns keep N * T: the reference assigns such views into its [N * T] buffer and
            # fails on the batch-size mismatch (src/rl8/algorithms/_feedforward.py:474-482); pairing the shorter
            # views with row indices over N * T would silently train on misaligned rows.
            raise RuntimeError(
                f"view requirements that drop {model.drop_size} leading step(s) per environment (method="
                "'rolling_window', shift > 0) cannot be batched with the [num_envs * horizon] transition rows;"
                " use method='padded_rolling_window' (the reference's feedforward algorithm has the same limit)"
            )
        views = model.apply_view_requirements(
            {k: buf[k][:, :-1] for k in model.view_requirements}, kind="all"
        )

        def column(key: str) -> torch.Tensor:  # [N, T+1, 1] view -> contiguous [N * T], row = n * T + t
            return buf[key][:, :-1].reshape(N * T).contiguous()

        actions, logp_old = column(DataKeys.ACTIONS), column(DataKeys.LOGP)
        advantages, returns = column(DataKeys.ADVANTAGES), column(DataKeys.RETURNS)
        M = hp.sgd_minibatch_size
        accum = hp.num_minibatches if hp.accumulate_grads else 1
        entropy_coeff = self.entropy_scheduler.coeff
        if entropy_coeff != 0 and kind == _lib.DIST_SQUASHED_NORMAL:
            dist_cls({}, model).entropy()  # raises like the reference
        ppo = _lib.PpoHparams(
            hp.clip_param, hp.dual_clip_param or 0.0, entropy_coeff, hp.vf_clip_param, hp.vf_coeff, 1.0 / accum,
        )
        sums = self._loss_sums
        sums.zero_()
        params = [p for p in model.parameters() if p.requires_grad]
        self.optimizer.zero_grad(set_to_none=True)
        model.train()
        k = 0
        applied: list[bool] = []
        stop_early = False
        for _ in range(hp.num_sgd_iters):
            perm = torch.randperm(N * T, device=self.device) if hp.shuffle_minibatches else None
            for i in range(hp.num_minibatches):
                step_this_batch = (i + 1) % accum == 0
                idx = perm[i * M : (i + 1) * M] if perm is not None else slice(i * M, (i + 1) * M)
                mb_views = _index_views(views, idx)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=hp.enable_amp):
                    features = model(mb_views)
                    values = model.value_function()
                packed = dist_cls(features, model)._packed()  # [M, P], differentiable
                vflat = values.float().reshape(-1)
                P = packed.shape[1]
                d_feat, d_val = torch.empty_like(packed), torch.empty_like(vflat)
                acts = actions[idx].contiguous()
                rc = lib.rl8_ppo_losses_direct(
                    kind, _lib.ptr(packed.detach()), P, _lib.ptr(vflat.detach().contiguous()), _lib.ptr(acts),
                    _lib.ptr(logp_old[idx].contiguous()), _lib.ptr(advantages[idx].contiguous()),
                    _lib.ptr(returns[idx].contiguous()), M, float(M * world), ppo,
                    ctypes.c_void_p(sums.data_ptr() + 40 * k), _lib.ptr(d_feat), _lib.ptr(d_val), _lib.stream(),
                )
                _lib.check(rc, "rl8_ppo_losses_direct")
                launches += 1
                applied.append(step_this_batch)
                k += 1
                if hp.target_kl_div is not None:
                    row = sums[k - 1].clone()
                    if world > 1:
                        dist.all_reduce(row)
                    if float(row[3] / row[4]) > 1.5 * hp.target_kl_div:
                        stop_early = True
                        self.optimizer.zero_grad(set_to_none=True)
                        break
                torch.autograd.backward([packed, vflat], [d_feat, d_val])
                if step_this_batch:
                    if world > 1:
                        for p in params:
                            if p.grad is not None:
                                dist.all_reduce(p.grad)  # already carries 1 / (M * world)
                    if self._on_grads is not None:
                        self._on_grads({n: p.grad for n, p in model.named_parameters() if p.grad is not None})
                    torch.nn.utils.clip_grad_norm_(params, hp.max_grad_norm)
                    self.optimizer.step()
                    self.optimizer.zero_grad(set_to_none=True)
                    self._opt_steps += 1
            if stop_early:
                break
        return self._finish_step(start, launches, k, applied, accum, entropy_coeff)

    # -- pieces of step() the recurrent algorithm overrides --------------------------------------
    def _batch_struct(self) -> _lib.Batch:
        hp, buf = self.hparams, self.buffer
        batch = _lib.Batch()
        batch.dist_kind = self.policy.distribution_cls.rl8_kind
        batch.T, batch.N = hp.horizon, hp.num_envs
        batch.obs = buf.hm[DataKeys.OBS].data_ptr()
        batch.actions = buf.hm[DataKeys.ACTIONS].data_ptr()
        batch.logp = buf.hm[DataKeys.LOGP].data_ptr()
        batch.advantages = buf.hm[DataKeys.ADVANTAGES].data_ptr()
        batch.returns = buf.hm[DataKeys.RETURNS].data_ptr()
        return batch

    def _update_units(self) -> tuple[int, int]:
        """(number of minibatch units in the buffer, transitions per unit)."""
        return self.hparams.num_envs * self.hparams.horizon, 1

    def _minibatch_launcher(self, batch: Any, M: int) -> tuple[Any, int]:
        """Returns ``(launch(rows, begin, denominator, ppo, sums_ptr), kernel launches per call)``."""
        lib, model, prec = self._lib, self.policy.model, self.policy.precision
        m = model.struct_for(model.flat_params)
        g = model.struct_for(self._grads)
        nbytes = int(lib.rl8_ppo_workspace(m, M, prec))
        if nbytes < 0:
            _lib.check(nbytes, "rl8_ppo_workspace")
        ws = self._workspace("ppo", nbytes)

        def launch(rows: Any, begin: int, denom: float, ppo: Any, sums_ptr: Any) -> None:
            rc = lib.rl8_ppo_minibatch(
                m, g, batch, _lib.ptr(rows), begin, M, denom, ppo, sums_ptr, prec, _lib.ptr(ws),
                ws.numel(), _lib.stream(),
            )
            _lib.check(rc, "rl8_ppo_minibatch")

        # bf16: 2 x W2 packing + (activation kernel + weight-gradient kernel) per 2^21-row chunk
        per_call = {
            _lib.PREC_BF16: 2 + 2 * max(1, -(-M // (1 << 21))),     # 2 x W2 packing + 2 kernels per 2^21-row chunk
            # max |obs| + the four W2 piece images (one launch) + 3 kernels per chunk
            _lib.PREC_FP32_TC: 2 + 3 * max(1, -(-M // (1 << 21))),
            _lib.PREC_FP32: 34 * max(1, -(-M // 65536)),            # the CUDA-core chain per 65 536-row chunk
        }[prec]
        return launch, per_call

    def _reset_buffer(self) -> None:
        """Fresh (zeroed) buffer that only keeps the final observation (:603-610)."""
        buf, T = self.buffer, self.hparams.horizon
        final_obs = buf.hm[DataKeys.OBS][T].clone()
        buf.zero_()
        buf.hm[DataKeys.OBS][T].copy_(final_obs)

    # ------------------------------------------------------------------------------------
    def validate(self) -> None:
        """Shape checks on one reset / sample / step (src/rl8/algorithms/_feedforward.py:617-697)."""
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
        sample = self.policy.sample(
            {DataKeys.OBS: self.buffer[DataKeys.OBS][:, :1]},
            kind="last",
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
