"""CPU restatement of the reference's view requirements (src/rl8/views.py).  TEST INFRASTRUCTURE:
imported only by ``tests/`` (and the golden generator); never by the product path.

Parity: **pinned** -- ``tests/golden/generate_views_golden.py`` runs the unmodified upstream
``rl8.views`` functions (behind ``oracle/refshim``) on the input shapes of the reference's own
``tests/test_views.py`` and on seeded random tensors, asserts this restatement agrees bit for bit,
and stores inputs + reference outputs in ``tests/golden/views.npz``.

Everything here is index arithmetic on ``[B, T, *F]`` arrays (plain numpy, no torch):

* ``rolling_window``        src/rl8/views.py:121-150  (``unfold`` + ``permute``)
* ``pad_last_sequence``     :57-88
* ``pad_whole_sequence``    :91-118
* ``RollingWindow``         :153-231
* ``PaddedRollingWindow``   :234-310
* ``ViewRequirement``       :313-453
"""

from __future__ import annotations

import numpy as np


def windows(x: np.ndarray, size: int, t_first: int, count: int) -> tuple[np.ndarray, np.ndarray]:
    """The one primitive every view reduces to: ``out[b, w, s] = x[b, t_first + w + s]`` for
    ``w < count``, ``s < size``, zero (and ``mask`` true) where that time index is negative.

    Returns ``(out [B, count, size, *F], mask [B, count, size] bool)``.
    """
    B, T = x.shape[:2]
    out = np.zeros((B, count, size) + x.shape[2:], dtype=x.dtype)
    mask = np.zeros((B, count, size), dtype=bool)
    for w in range(count):
        for s in range(size):
            t = t_first + w + s
            if t < 0:
                mask[:, w, s] = True
            else:
                assert t < T
                out[:, w, s] = x[:, t]
    return out, mask


def rolling_window(x: np.ndarray, size: int, step: int = 1) -> np.ndarray:
    """``[B, (T - size) / step + 1, size, *F]`` (src/rl8/views.py:121-150)."""
    T = x.shape[1]
    count = (T - size) // step + 1
    return np.stack([x[:, w * step : w * step + size] for w in range(count)], axis=1)


def pad_last_sequence(x: np.ndarray, size: int) -> dict[str, np.ndarray]:
    """Last ``size`` steps, zero-padded in front when ``T < size`` (:57-88)."""
    T = x.shape[1]
    out, mask = windows(x, size, T - size, 1)
    return {"inputs": out[:, 0], "padding_mask": mask[:, 0]}


def pad_whole_sequence(x: np.ndarray, size: int) -> dict[str, np.ndarray]:
    """``size - 1`` zero steps in front of every sequence (:91-118)."""
    B, T = x.shape[:2]
    pad = size - 1
    padding = np.zeros((B, pad) + x.shape[2:], dtype=x.dtype)
    mask = np.zeros((B, T + pad), dtype=bool)
    mask[:, :pad] = True
    return {"inputs": np.concatenate([padding, x], axis=1), "padding_mask": mask}


def rolling_window_apply_all(x: np.ndarray, size: int) -> np.ndarray:
    """``[B * (T - size + 1), size, *F]`` (:158-193)."""
    T = x.shape[1]
    out, _ = windows(x, size, 0, T - size + 1)
    return out.reshape((-1, size) + x.shape[2:])


def rolling_window_apply_last(x: np.ndarray, size: int) -> np.ndarray:
    """``x[:, -size:]`` (:195-221) -- shorter than ``size`` when ``T < size``."""
    return x[:, -size:]


def padded_rolling_window_apply_all(x: np.ndarray, size: int) -> dict[str, np.ndarray]:
    """``[B * T, size, *F]`` inputs + ``[B * T, size]`` padding mask (:245-281)."""
    T = x.shape[1]
    out, mask = windows(x, size, -(size - 1), T)
    return {
        "inputs": out.reshape((-1, size) + x.shape[2:]),
        "padding_mask": mask.reshape(-1, size),
    }


def padded_rolling_window_apply_last(x: np.ndarray, size: int) -> dict[str, np.ndarray]:
    return pad_last_sequence(x, size)


def view_apply_all(x: np.ndarray, shift: int, method: str):  # noqa: ANN201
    """``ViewRequirement.apply_all`` on a tensor item (:365-407)."""
    if not shift:
        return x.reshape((-1,) + x.shape[2:])
    if method == "rolling_window":
        return rolling_window_apply_all(x, shift + 1)
    return padded_rolling_window_apply_all(x, shift + 1)


def view_apply_last(x: np.ndarray, shift: int, method: str):  # noqa: ANN201
    """``ViewRequirement.apply_last`` on a tensor item (:409-445)."""
    if not shift:
        return x[:, -1]
    if method == "rolling_window":
        return rolling_window_apply_last(x, shift + 1)
    return padded_rolling_window_apply_last(x, shift + 1)


def drop_size(shift: int, method: str) -> int:
    """Samples lost at the start of every sequence (:223-231, 303-310, 447-453)."""
    return shift if method == "rolling_window" else 0
