"""Stop conditions for ``Trainer.run`` (host-side; behaviour of src/rl8/conditions.py).

A condition is any callable ``(train_stats) -> bool``; training stops as soon as one of the
conditions passed to ``Trainer.run`` returns ``True``.
"""

from __future__ import annotations

from typing import Any, Callable, Mapping

Condition = Callable[[Mapping[str, Any]], bool]


class And:
    """True when every wrapped condition is true."""

    def __init__(self, conditions: list[Condition], /) -> None:
        self.conditions = conditions

    def __call__(self, train_stats: Mapping[str, Any], /) -> bool:
        # every condition is evaluated (they may keep state), like the reference's list form
        return all([c(train_stats) for c in self.conditions])


class HitsLowerBound:
    """True once ``train_stats[key] <= lower_bound``."""

    def __init__(self, key: str, lower_bound: float, /) -> None:
        self.key, self.lower_bound = key, lower_bound

    def __call__(self, train_stats: Mapping[str, Any], /) -> bool:
        return bool(train_stats[self.key] <= self.lower_bound)


class HitsUpperBound:
    """True once ``train_stats[key] >= upper_bound``."""

    def __init__(self, key: str, upper_bound: float, /) -> None:
        self.key, self.upper_bound = key, upper_bound

    def __call__(self, train_stats: Mapping[str, Any], /) -> bool:
        return bool(train_stats[self.key] >= self.upper_bound)


class _Streak:
    def __init__(self, key: str, patience: int) -> None:
        self.key, self.patience, self.losses = key, patience, 0

    def _count(self, bad: bool) -> bool:
        self.losses = self.losses + 1 if bad else 0
        return self.losses >= self.patience


class Plateaus(_Streak):
    """True after ``patience`` consecutive calls whose value moved by at most ``rtol``
    (relative to the previous value)."""

    def __init__(self, key: str, /, *, patience: int = 5, rtol: float = 1e-3) -> None:
        super().__init__(key, patience)
        self.rtol = rtol
        self.old_value = 0.0

    def __call__(self, train_stats: Mapping[str, Any], /) -> bool:
        new = train_stats[self.key]
        bad = abs(new - self.old_value) <= self.rtol * abs(self.old_value)
        self.old_value = new
        return self._count(bad)


class StopsDecreasing(_Streak):
    """True after ``patience`` consecutive calls without a new minimum."""

    def __init__(self, key: str, /, *, patience: int = 5) -> None:
        super().__init__(key, patience)
        self.min_ = float("inf")

    def __call__(self, train_stats: Mapping[str, Any], /) -> bool:
        new = train_stats[self.key]
        bad = new >= self.min_
        if not bad:
            self.min_ = new
        return self._count(bad)


class StopsIncreasing(_Streak):
    """True after ``patience`` consecutive calls without a new maximum."""

    def __init__(self, key: str, /, *, patience: int = 5) -> None:
        super().__init__(key, patience)
        self.max_ = float("-inf")

    def __call__(self, train_stats: Mapping[str, Any], /) -> bool:
        new = train_stats[self.key]
        bad = new <= self.max_
        if not bad:
            self.max_ = new
        return self._count(bad)
