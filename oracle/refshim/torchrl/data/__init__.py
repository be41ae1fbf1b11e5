"""Container-only stand-in for ``torchrl.data`` tensor specs (TEST INFRASTRUCTURE).

Specs here are shape/dtype/device records with ``zero()``; they do no arithmetic.
See ``oracle/refshim/README.md``.
"""

from __future__ import annotations

from typing import Any

import torch
from tensordict import TensorDict


def _shape(shape: Any) -> torch.Size:
    if shape is None:
        return torch.Size([])
    if isinstance(shape, int):
        return torch.Size([shape])
    return torch.Size(list(shape))


class TensorSpec:
    shape: torch.Size
    dtype: torch.dtype
    device: Any

    def __init__(self, shape=None, *, device=None, dtype=torch.float32) -> None:
        self.shape = _shape(shape)
        self.device = device
        self.dtype = dtype

    @property
    def ndim(self) -> int:
        return len(self.shape)

    def to(self, device):
        self.device = device
        return self

    def zero(self, batch_shape=None) -> torch.Tensor:
        return torch.zeros(*_shape(batch_shape), *self.shape, dtype=self.dtype, device=self.device)

    def rand(self, batch_shape=None) -> torch.Tensor:
        return torch.randn(*_shape(batch_shape), *self.shape, device=self.device).to(self.dtype)

    def encode(self, x: Any) -> torch.Tensor:
        return torch.as_tensor(x, dtype=self.dtype, device=self.device)

    def assert_is_in(self, x: torch.Tensor) -> None:
        assert tuple(x.shape[-self.ndim :]) == tuple(self.shape) or self.ndim == 0, (
            f"{tuple(x.shape)} is not in spec of shape {tuple(self.shape)}"
        )

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(shape={tuple(self.shape)}, dtype={self.dtype})"


class Unbounded(TensorSpec):
    pass


class _Space:
    def __init__(self, n: int) -> None:
        self.n = n


class Categorical(TensorSpec):
    def __init__(self, n: int, shape=None, *, device=None, dtype=torch.int64) -> None:
        super().__init__(shape, device=device, dtype=dtype)
        self.space = _Space(n)

    def rand(self, batch_shape=None) -> torch.Tensor:
        return torch.randint(
            0, self.space.n, (*_shape(batch_shape), *self.shape), device=self.device
        ).to(self.dtype)


class Composite(TensorSpec):
    def __init__(self, specs=None, *, device=None, **kwargs: Any) -> None:
        super().__init__(None, device=device)
        self._specs: dict[str, TensorSpec] = {}
        for k, v in {**(specs or {}), **kwargs}.items():
            self.set(k, v)

    def set(self, key: str, spec: TensorSpec) -> None:
        self._specs[key] = spec

    def keys(self):
        return self._specs.keys()

    def items(self):
        return self._specs.items()

    def __iter__(self):
        return iter(self._specs)

    def __getitem__(self, key: str) -> TensorSpec:
        return self._specs[key]

    def __contains__(self, key: str) -> bool:
        return key in self._specs

    def to(self, device):
        self.device = device
        for v in self._specs.values():
            v.to(device)
        return self

    def zero(self, batch_shape=None) -> TensorDict:
        return TensorDict(
            {k: v.zero(batch_shape) for k, v in self._specs.items()},
            batch_size=_shape(batch_shape),
            device=self.device,
        )

    def rand(self, batch_shape=None) -> TensorDict:
        return TensorDict(
            {k: v.rand(batch_shape) for k, v in self._specs.items()},
            batch_size=_shape(batch_shape),
            device=self.device,
        )

    def encode(self, x: Any) -> TensorDict:
        return TensorDict({k: self._specs[k].encode(v) for k, v in x.items()}, batch_size=[])

    def assert_is_in(self, x: Any) -> None:
        for k, v in self._specs.items():
            v.assert_is_in(x[k])
