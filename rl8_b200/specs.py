"""Minimal tensor specs: the slice of ``torchrl.data`` the hot path touches.

The reference describes observation / action / buffer layouts with torchrl specs
(``Unbounded``, ``Categorical``, ``Composite``; src/rl8/env.py:6, src/rl8/algorithms/
_feedforward.py:239-256).  torchrl is not a dependency here, so these duck-typed records
carry shape / dtype / device and the handful of methods the algorithm layer calls.
"""

from __future__ import annotations

from typing import Any, Iterator, Sequence

import torch


def _size(shape: int | Sequence[int] | torch.Size) -> torch.Size:
    return torch.Size([shape]) if isinstance(shape, int) else torch.Size(shape)


class TensorSpec:
    """Shape, dtype and device of one element (without batch dimensions)."""

    def __init__(self, shape: Any = 1, *, device: Any = "cpu", dtype: torch.dtype = torch.float32):
        self.shape = _size(shape)
        self.device = torch.device(device)
        self.dtype = dtype

    def to(self, device: Any) -> "TensorSpec":
        out = self.__class__.__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        out.device = torch.device(device)
        return out

    def zero(self, batch: Sequence[int] = ()) -> torch.Tensor:
        return torch.zeros(*batch, *self.shape, dtype=self.dtype, device=self.device)

    def is_in(self, value: torch.Tensor) -> bool:
        n = len(self.shape)
        return value.dtype == self.dtype and (n == 0 or value.shape[-n:] == self.shape)

    def assert_is_in(self, value: torch.Tensor) -> None:
        if not self.is_in(value):
            raise AssertionError(
                f"value of shape {tuple(value.shape)} / dtype {value.dtype} is not in {self!r}"
            )

    def __repr__(self) -> str:
        return f"{type(self).__name__}(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device})"


class Unbounded(TensorSpec):
    """Real-valued element of a fixed shape."""

    def __init__(self, shape: Any = 1, *, device: Any = "cpu", dtype: torch.dtype = torch.float32):
        super().__init__(shape, device=device, dtype=dtype)

    def rand(self, batch: Sequence[int] = ()) -> torch.Tensor:
        return torch.randn(*batch, *self.shape, dtype=self.dtype, device=self.device)


class _Space:
    def __init__(self, n: int) -> None:
        self.n = n


class Categorical(TensorSpec):
    """Integer element in ``[0, n)`` (``torchrl.data.Categorical``; int64 by default)."""

    def __init__(
        self, n: int, shape: Any = (), *, device: Any = "cpu", dtype: torch.dtype = torch.int64
    ):
        super().__init__(shape, device=device, dtype=dtype)
        self.n = n
        self.space = _Space(n)

    def rand(self, batch: Sequence[int] = ()) -> torch.Tensor:
        return torch.randint(0, self.n, (*batch, *self.shape), dtype=self.dtype, device=self.device)

    def is_in(self, value: torch.Tensor) -> bool:
        return super().is_in(value) and bool(((value >= 0) & (value < self.n)).all())


class Composite:
    """Named collection of specs."""

    def __init__(self, specs: None | dict[str, TensorSpec] = None) -> None:
        self._specs: dict[str, TensorSpec] = dict(specs or {})

    def set(self, key: str, spec: TensorSpec) -> None:
        self._specs[key] = spec

    def __getitem__(self, key: str) -> TensorSpec:
        return self._specs[key]

    def __contains__(self, key: str) -> bool:
        return key in self._specs

    def __iter__(self) -> Iterator[str]:
        return iter(self._specs)

    def keys(self) -> Any:
        return self._specs.keys()

    def items(self) -> Any:
        return self._specs.items()

    def to(self, device: Any) -> "Composite":
        return Composite({k: v.to(device) for k, v in self._specs.items()})

    def zero(self, batch: Sequence[int] = ()) -> dict[str, torch.Tensor]:
        return {k: v.zero(batch) for k, v in self._specs.items()}
