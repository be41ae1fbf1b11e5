"""View requirements: sliding windows over the time axis of batch items
(drop-in for src/rl8/views.py).

Same names, arguments and results as the reference -- ``rolling_window``,
``pad_last_sequence``, ``pad_whole_sequence``, ``RollingWindow``, ``PaddedRollingWindow``,
``ViewRequirement`` -- with every materialising view on ONE kernel, ``rl8_view_windows``
(include/rl8_b200.h): the reference's ``unfold + permute + reshape`` gather becomes a
shared-memory-staged copy that reads the item once, coalesced for both the horizon-major
rollout buffer (``buffer["obs"]`` is a strided view of ``[T+1][D][N]``) and ordinary env-major
tensors, and writes each sequence's windows as one contiguous run.

Items are CUDA tensors ``[B, T, ...]`` or (nested) mappings of them (the reference accepts
tensordicts); mappings map leaf-wise, like ``TensorDict.apply``.  Padded views return
``{"inputs": ..., "padding_mask": ...}`` per tensor.  Views that are pure slicing in the
reference (``rolling_window`` itself, ``RollingWindow.apply_last``, shift-0 requirements) stay
zero-copy torch views here too.  There is no CPU path: tensors must live on the GPU.
"""

from __future__ import annotations

from typing import Any, Callable, Literal, Mapping, Protocol

import torch

from . import _lib
from .data import DataKeys

ViewKind = Literal["last", "all"]
ViewMethod = Literal["rolling_window", "padded_rolling_window"]

Item = Any  # torch.Tensor | Mapping[str, Item]


def _map(fn: Callable[[torch.Tensor], Any], x: Item) -> Any:
    """Apply ``fn`` to every tensor leaf (``TensorDict.apply`` of the reference)."""
    if isinstance(x, torch.Tensor):
        return fn(x)
    if isinstance(x, Mapping) or hasattr(x, "items"):
        return {k: _map(fn, v) for k, v in x.items()}
    raise TypeError(f"view requirements apply to tensors or mappings of tensors, not {type(x).__name__}")


def _windows(
    x: torch.Tensor, size: int, t_first: int, count: int, *, want_mask: bool
) -> tuple[torch.Tensor, None | torch.Tensor]:
    """``out[b, w, s] = x[b, t_first + w + s]`` (zero / masked where negative): ``([B, count,
    size, *F], [B, count, size] bool | None)`` through ``rl8_view_windows``."""
    _lib.require_cuda(x, "x")
    if x.dim() < 2:
        raise ValueError("view requirements need items of shape [B, T, ...]")
    lib = _lib.load()
    B, T = x.shape[:2]
    feat = tuple(x.shape[2:])
    F = 1
    for d in feat:
        F *= int(d)
    elem = x.element_size()
    src = x
    if elem not in (4, 8):
        raise NotImplementedError(f"rl8_view_windows moves 4- and 8-byte elements, not {x.dtype}")
    # one feature stride: flatten trailing dims (a view for every layout the buffer produces)
    src = x.reshape(B, T, F) if x.dim() != 3 else x
    out = torch.empty((B, count, size) + feat, dtype=x.dtype, device=x.device)
    mask = torch.empty((B, count, size), dtype=torch.bool, device=x.device) if want_mask else None
    rc = lib.rl8_view_windows(
        _lib.ptr(src), elem, B, T, F, src.stride(0), src.stride(1), src.stride(2) if F else 1, size,
        t_first, count, _lib.ptr(out), _lib.ptr(mask), _lib.stream(),
    )
    _lib.check(rc, "rl8_view_windows")
    return out, mask


def pad_last_sequence(x: torch.Tensor, size: int, /) -> dict[str, torch.Tensor]:
    """Last ``size`` steps of ``x [B, T, ...]``, zero-padded in front when ``T < size``
    (src/rl8/views.py:57-88): ``{"inputs": [B, size, ...], "padding_mask": [B, size]}``."""
    T = x.shape[1]
    out, mask = _windows(x, size, T - size, 1, want_mask=True)
    assert mask is not None
    return {DataKeys.INPUTS: out[:, 0], DataKeys.PADDING_MASK: mask[:, 0]}


def pad_whole_sequence(x: torch.Tensor, size: int, /) -> dict[str, torch.Tensor]:
    """``size - 1`` zero steps in front of every sequence (src/rl8/views.py:91-118):
    ``{"inputs": [B, T + size - 1, ...], "padding_mask": [B, T + size - 1]}``."""
    T = x.shape[1]
    pad = RollingWindow.drop_size(size)
    out, mask = _windows(x, 1, -pad, T + pad, want_mask=True)
    assert mask is not None
    return {DataKeys.INPUTS: out[:, :, 0], DataKeys.PADDING_MASK: mask[:, :, 0]}


def rolling_window(x: torch.Tensor, size: int, /, *, step: int = 1) -> torch.Tensor:
    """``[B, (T - size) / step + 1, size, ...]`` rolling windows of ``x [B, T, ...]`` as a
    zero-copy strided view (src/rl8/views.py:121-150)."""
    dims = list(range(x.dim()))
    dims.insert(2, -1)
    return x.unfold(1, size, step).permute(*dims)


class View(Protocol):
    """Protocol of a view method (src/rl8/views.py:14-54)."""

    @staticmethod
    def apply_all(x: Item, size: int, /) -> Item: ...

    @staticmethod
    def apply_last(x: Item, size: int, /) -> Item: ...

    @staticmethod
    def drop_size(size: int, /) -> int: ...


class RollingWindow:
    """Rolling windows without padding: the first ``size - 1`` samples of every sequence
    are dropped (src/rl8/views.py:153-231)."""

    @staticmethod
    def apply_all(x: Item, size: int, /) -> Item:
        """``[B * (T - size + 1), size, ...]``."""

        def one(t: torch.Tensor) -> torch.Tensor:
            T = t.shape[1]
            if T < size:
                raise RuntimeError(
                    f"maximum size for tensor at dimension 1 is {T} but size is {size}"
                )
            out, _ = _windows(t, size, 0, T - size + 1, want_mask=False)
            return out.reshape((-1, size) + tuple(t.shape[2:]))

        return _map(one, x)

    @staticmethod
    def apply_last(x: Item, size: int, /) -> Item:
        """``x[:, -size:]`` (a view)."""
        return _map(lambda t: t[:, -size:, ...], x)

    @staticmethod
    def drop_size(size: int, /) -> int:
        return size - 1


class PaddedRollingWindow:
    """Rolling windows with zero padding and a padding mask, so no sample is dropped
    (src/rl8/views.py:234-310)."""

    @staticmethod
    def apply_all(x: Item, size: int, /) -> Item:
        """``{"inputs": [B * T, size, ...], "padding_mask": [B * T, size]}``."""

        def one(t: torch.Tensor) -> dict[str, torch.Tensor]:
            T = t.shape[1]
            out, mask = _windows(t, size, -(size - 1), T, want_mask=True)
            assert mask is not None
            return {
                DataKeys.INPUTS: out.reshape((-1, size) + tuple(t.shape[2:])),
                DataKeys.PADDING_MASK: mask.reshape(-1, size),
            }

        return _map(one, x)

    @staticmethod
    def apply_last(x: Item, size: int, /) -> Item:
        """``{"inputs": [B, size, ...], "padding_mask": [B, size]}``."""
        return _map(lambda t: pad_last_sequence(t, size), x)

    @staticmethod
    def drop_size(size: int, /) -> int:
        return size - size


class ViewRequirement:
    """Batch preprocessing that gives a model the last ``shift + 1`` samples of an item
    (src/rl8/views.py:313-453).

    Args:
        shift: additional previous samples along the time axis to include.
        method: ``"rolling_window"`` (drops the first ``shift`` samples of every sequence) or
            ``"padded_rolling_window"`` (zero-pads and returns a padding mask).
    """

    method: type[View]
    shift: int

    def __init__(self, *, shift: int = 0, method: ViewMethod = "padded_rolling_window") -> None:
        self.shift = shift
        if shift < 0:
            raise ValueError(f"{self.__class__.__name__} `shift` must be non-negative.")
        match method:
            case "rolling_window":
                self.method = RollingWindow
            case "padded_rolling_window":
                self.method = PaddedRollingWindow

    def apply_all(self, key: str | tuple[str, ...], batch: Mapping[str, Any], /) -> Item:
        """All time steps of ``batch[key] [B, T, ...]``: ``[B_NEW, shift + 1, ...]`` with
        ``B_NEW <= B * T``, or ``[B * T, ...]`` when ``shift == 0``."""
        item = _get(batch, key)
        if not self.shift:
            return _map(lambda t: t.flatten(end_dim=1), item)
        return self.method.apply_all(item, self.shift + 1)

    def apply_last(self, key: str | tuple[str, ...], batch: Mapping[str, Any], /) -> Item:
        """The most recent samples of ``batch[key]``: ``[B, shift + 1, ...]``, or ``[B, ...]``
        when ``shift == 0``."""
        item = _get(batch, key)
        if not self.shift:
            return _map(lambda t: t[:, -1, ...], item)
        return self.method.apply_last(item, self.shift + 1)

    @property
    def drop_size(self) -> int:
        """Samples dropped at the start of every sequence by the method."""
        return self.method.drop_size(self.shift + 1)


def _get(batch: Mapping[str, Any], key: str | tuple[str, ...]) -> Item:
    if isinstance(key, tuple):
        item: Any = batch
        for k in key:
            item = item[k]
        return item
    return batch[key]
