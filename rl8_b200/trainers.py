"""High-level training loop (src/rl8/trainers/_base.py:16-201).

``Trainer(algo).step()`` = ``memory_stats -> collect -> step``; ``run()`` repeats it until a
stop condition fires, optionally interleaving ``eval()``.  The reference logs every stats
dict to MLflow from the training thread; MLflow is not a dependency here -- pass a
``log_fn(stats, step)`` callable to receive the same dictionaries.
"""

from __future__ import annotations

from collections import defaultdict
from typing import Any, Callable

from .algorithms import Algorithm
from .conditions import Condition
from .data import TrainerState

LogFn = Callable[[dict[str, Any], int], None]


def reduce_stats(x: dict[str, list[float]], /) -> dict[str, float]:
    """Reduce per-horizon stats by the operation their key names (min, max, mean, std as the
    root mean square, anything else summed; src/rl8/_utils.py:128-144)."""
    y: dict[str, float] = {}
    for k, v in x.items():
        op = k.split("/")[-1]
        if op == "min":
            y[k] = min(v)
        elif op == "max":
            y[k] = max(v)
        elif op == "mean":
            y[k] = sum(v) / len(v)
        elif op == "std":
            y[k] = (sum(s * s for s in v) / len(v)) ** 0.5
        else:
            y[k] = sum(v)
    return y


class Trainer:
    """Owns an :class:`Algorithm` and the running counters."""

    def __init__(self, algorithm: Algorithm, /, *, log_fn: None | LogFn = None) -> None:
        self.algorithm = algorithm
        self.state: TrainerState = {"algorithm/collects": 0, "algorithm/steps": 0, "env/steps": 0}
        self.log_fn = log_fn

    def _log(self, stats: dict[str, Any]) -> None:
        if self.log_fn is not None:
            self.log_fn(stats, self.state["env/steps"])

    def eval(
        self, *, env_config: None | dict[str, Any] = None, deterministic: bool = True
    ) -> dict[str, Any]:
        """Collect ``horizons_per_env_reset`` horizons without learning; shares the training
        buffer, so it may only run on a reset boundary (src/rl8/trainers/_base.py:43-102)."""
        hper = self.algorithm.horizons_per_env_reset
        if env_config and hper < 0 and self.state["algorithm/collects"]:
            raise ValueError(
                "An eval environment config was provided even though the environment is not"
                " expected to use the config because `horizons_per_env_reset` is < 0. Either do"
                " not provide an eval environment config or set `horizons_per_env_reset` > 0."
            )
        if hper > 0 and self.state["algorithm/collects"] % hper:
            raise RuntimeError(
                "Trainer.eval can only be called every `horizons_per_env_reset`: training and"
                " evaluation share one buffer."
            )
        stats: dict[str, list[float]] = defaultdict(list)
        for _ in range(max(1, hper)):
            collected = self.algorithm.collect(env_config=env_config, deterministic=deterministic)
            for k, v in collected.items():
                stats[k].append(v)  # type: ignore[arg-type]
            self.state["algorithm/collects"] += 1
        eval_stats = {f"eval/{k}": v for k, v in reduce_stats(stats).items()}
        self._log(eval_stats)
        return eval_stats

    def run(
        self,
        *,
        env_config: None | dict[str, Any] = None,
        eval_env_config: None | dict[str, Any] = None,
        steps_per_eval: None | int = None,
        stop_conditions: None | list[Condition] = None,
    ) -> dict[str, Any]:
        """Train until one stop condition is true (forever without conditions)."""
        hper = self.algorithm.horizons_per_env_reset
        if steps_per_eval and hper < 0 and eval_env_config:
            raise ValueError(
                "An eval environment config was provided even though `horizons_per_env_reset`"
                " is < 0 (the environment is reset once, at the beginning of training)."
            )
        if steps_per_eval and hper > 0 and steps_per_eval % hper:
            raise ValueError(
                "Trainer.eval can only be called every `horizons_per_env_reset`; set"
                " `steps_per_eval` to a multiple of it."
            )
        eval_env_config = eval_env_config or env_config
        stop_conditions = stop_conditions or []
        train_stats = self.step(env_config=env_config)
        while not any([condition(train_stats) for condition in stop_conditions]):
            if steps_per_eval and not (self.state["algorithm/steps"] % steps_per_eval):
                self.eval(env_config=eval_env_config)
            train_stats = self.step(env_config=env_config)
        return train_stats

    def step(self, *, env_config: None | dict[str, Any] = None) -> dict[str, Any]:
        """One ``collect`` + one ``step`` of the algorithm."""
        memory_stats = self.algorithm.memory_stats()
        collect_stats = self.algorithm.collect(env_config=env_config)
        step_stats = self.algorithm.step()
        train_stats: dict[str, Any] = {**memory_stats, **collect_stats, **step_stats}
        self.state["algorithm/collects"] += 1
        self.state["algorithm/steps"] += 1
        self.state["env/steps"] += collect_stats["env/steps"]
        train_stats.update(self.state)
        self._log(train_stats)
        return train_stats
