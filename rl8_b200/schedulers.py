"""Host-side value schedules for the learning rate and the entropy coefficient.

Same behaviour as src/rl8/schedulers.py:121-232: schedules are lists of
``(env transition count, value)`` pairs starting at count 0; ``"step"`` holds a value until
the next count, ``"interp"`` interpolates linearly.  They only produce the scalars handed
to the kernels (``lr`` of the fused Adam, ``entropy_coeff`` of the loss), so they stay in
Python.
"""

from __future__ import annotations

from typing import Any, Literal

import numpy as np

ScheduleKind = Literal["interp", "step"]


class _Constant:
    def __init__(self, value: float) -> None:
        self.value = value

    def at(self, count: int) -> float:
        return self.value


class _Table:
    def __init__(self, schedule: list[tuple[int, float]], kind: str) -> None:
        if schedule[0][0]:
            raise ValueError("a schedule's first count (`schedule[0][0]`) must be 0")
        if kind not in ("interp", "step"):
            raise ValueError("schedulers only support kinds `interp` and `step`")
        self.counts = [int(c) for c, _ in schedule]
        self.values = [float(v) for _, v in schedule]
        self.kind = kind

    def at(self, count: int) -> float:
        if self.kind == "interp":
            return float(np.interp(count, self.counts, self.values))
        value = 0.0
        for c, v in zip(self.counts, self.values):
            if count >= c:
                value = v
        return value


class EntropyScheduler:
    """Entropy coefficient as a function of env transitions."""

    def __init__(
        self,
        coeff: float,
        /,
        *,
        schedule: None | list[tuple[int, float]] = None,
        kind: ScheduleKind = "step",
    ) -> None:
        self.scheduler: Any = _Constant(coeff) if schedule is None else _Table(schedule, kind)
        self.coeff = self.step(0)

    def step(self, count: int, /) -> float:
        self.coeff = self.scheduler.at(count)
        return self.coeff


class LRScheduler:
    """Learning rate as a function of env transitions; without a schedule the optimizer's own
    learning rate is never touched (src/rl8/schedulers.py:227-232)."""

    def __init__(
        self,
        optimizer: Any,
        /,
        *,
        schedule: None | list[tuple[int, float]] = None,
        kind: ScheduleKind = "step",
    ) -> None:
        self.optimizer = optimizer
        self.scheduler: Any = _Constant(0.0) if schedule is None else _Table(schedule, kind)
        self.coeff = self.step(0)

    def step(self, count: int, /) -> float:
        self.coeff = self.scheduler.at(count)
        if isinstance(self.scheduler, _Table):
            for pg in self.optimizer.param_groups:
                pg["lr"] = self.coeff
        return self.coeff
