"""Container-only stand-in for the ``tensordict`` package (TEST INFRASTRUCTURE).

Only the surface the upstream rl8 reference touches on its rollout/update path is
provided: a nested mapping of tensors that share leading batch dimensions, with
tensor-style indexing applied to every leaf. It performs no arithmetic of its own.
See ``oracle/refshim/README.md``.
"""

from __future__ import annotations

from typing import Any, Callable, Iterator

import torch


def _as_size(batch_size: Any) -> torch.Size:
    if batch_size is None:
        return torch.Size([])
    if isinstance(batch_size, int):
        return torch.Size([batch_size])
    return torch.Size(list(batch_size))


class TensorDict:
    def __init__(self, source=None, batch_size=None, device=None) -> None:
        self._d: dict[str, Any] = {}
        self.batch_size = _as_size(batch_size)
        self.device = device
        for k, v in (source or {}).items():
            self[k] = v

    # -- mapping protocol -------------------------------------------------
    def keys(self):
        return self._d.keys()

    def items(self):
        return self._d.items()

    def values(self):
        return self._d.values()

    def __contains__(self, key: str) -> bool:
        return key in self._d

    def __iter__(self) -> Iterator[Any]:
        raise TypeError("iteration over a TensorDict stand-in is not supported")

    def __len__(self) -> int:
        return self.batch_size[0] if len(self.batch_size) else 0

    def __delitem__(self, key: str) -> None:
        del self._d[key]

    # -- indexing ---------------------------------------------------------
    @staticmethod
    def _is_key(idx: Any) -> bool:
        if isinstance(idx, str):
            return True
        return (
            isinstance(idx, tuple) and len(idx) > 0 and all(isinstance(i, str) for i in idx)
        )

    def __getitem__(self, idx: Any) -> Any:
        if self._is_key(idx):
            if isinstance(idx, str):
                return self._d[idx]
            out = self
            for k in idx:
                out = out[k]
            return out
        probe = torch.empty(self.batch_size, device="meta")[idx]
        out = TensorDict({}, batch_size=probe.shape, device=self.device)
        for k, v in self._d.items():
            out._d[k] = v[idx]
        return out

    def __setitem__(self, idx: Any, value: Any) -> None:
        if self._is_key(idx):
            if isinstance(idx, tuple):
                if len(idx) == 1:
                    idx = idx[0]
                else:
                    self._d[idx[0]][idx[1:]] = value
                    return
            if isinstance(value, dict):
                value = TensorDict(value, batch_size=self.batch_size, device=self.device)
            self._d[idx] = value
            return
        for k, v in self._d.items():
            v[idx] = value[k] if isinstance(value, TensorDict) else value

    # -- shape ------------------------------------------------------------
    @property
    def shape(self) -> torch.Size:
        return self.batch_size

    def size(self, dim: None | int = None):
        return self.batch_size if dim is None else self.batch_size[dim]

    def numel(self) -> int:
        return self.batch_size.numel()

    def reshape(self, *shape: Any) -> "TensorDict":
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        nb = len(self.batch_size)
        new_bs = torch.empty(self.batch_size, device="meta").reshape(*shape).shape
        out = TensorDict({}, batch_size=new_bs, device=self.device)
        for k, v in self._d.items():
            if isinstance(v, TensorDict):
                out._d[k] = v.reshape(*new_bs)
            else:
                out._d[k] = v.reshape(*new_bs, *v.shape[nb:])
        return out

    def apply(self, fn: Callable[[Any], Any], batch_size=None) -> "TensorDict":
        out = TensorDict(
            {},
            batch_size=self.batch_size if batch_size is None else batch_size,
            device=self.device,
        )
        for k, v in self._d.items():
            out._d[k] = v.apply(fn, batch_size=batch_size) if isinstance(v, TensorDict) else fn(v)
        return out

    def select(self, *keys: str) -> "TensorDict":
        return TensorDict(
            {k: self._d[k] for k in keys}, batch_size=self.batch_size, device=self.device
        )

    def to(self, device) -> "TensorDict":
        return self.apply(lambda x: x.to(device)) if device is not None else self

    def clone(self) -> "TensorDict":
        return self.apply(lambda x: x.clone())

    def __eq__(self, other: Any):  # type: ignore[override]
        return self.apply(lambda x: x) if other is self else TensorDict(
            {k: (v == other[k]) for k, v in self._d.items()},
            batch_size=self.batch_size,
            device=self.device,
        )

    def all(self) -> bool:
        return all(bool(v.all()) for v in self._d.values())

    def __repr__(self) -> str:
        body = ", ".join(
            f"{k}: {tuple(v.shape) if hasattr(v, 'shape') else v}" for k, v in self._d.items()
        )
        return f"TensorDict({{{body}}}, batch_size={tuple(self.batch_size)})"
