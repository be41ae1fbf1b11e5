"""Golden vectors of the view requirements from the UNMODIFIED upstream ``rl8.views``.
Container-only tool (the reference does not exist on the GPU box)::

    PYTHONPATH=oracle/refshim:/root/reference/src python tests/golden/generate_views_golden.py

Cases: the input shapes of the reference's own tests (tests/test_views.py:15-431 -- ``arange``
tensors of shape [4,1], [2,2,2], [2,4,1,1,1], [2,4], ...) plus seeded random tensors with wider
features, int64 items and the rollout-buffer shape [N, T+1, D].  For every case and every view
function the script asserts ``oracle/views_oracle.py`` reproduces the reference bit for bit and
stores the input and the reference outputs in ``tests/golden/views.npz``.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from rl8.views import (  # noqa: E402  (upstream)
    PaddedRollingWindow,
    RollingWindow,
    ViewRequirement,
    pad_last_sequence,
    pad_whole_sequence,
    rolling_window,
)
from tensordict import TensorDict  # noqa: E402  (refshim)

from oracle import views_oracle as V  # noqa: E402


def arange(*shape: int) -> torch.Tensor:
    return torch.arange(int(np.prod(shape))).reshape(*shape).float()


def cases() -> dict[str, tuple[torch.Tensor, int]]:
    g = torch.Generator().manual_seed(11)
    out = {
        # shapes used by the reference's tests
        "ref_b4_t1": (arange(4, 1), 2),
        "ref_b2_t2_f2": (arange(2, 2, 2), 2),
        "ref_b2_t4_f111": (arange(2, 4, 1, 1, 1), 2),
        "ref_b2_t4": (arange(2, 4), 2),
        "ref_b2_t4_f2_size3": (arange(2, 4, 2), 3),
        "ref_b1_t3_size4": (arange(1, 3, 1), 4),
        # seeded random
        "rand_b7_t9_f5_size4": (torch.randn(7, 9, 5, generator=g), 4),
        "rand_b33_t33_f5_size3": (torch.randn(33, 33, 5, generator=g), 3),
        "rand_b5_t6_f3x2_size6": (torch.randn(5, 6, 3, 2, generator=g), 6),
        "rand_b3_t2_f4_size5": (torch.randn(3, 2, 4, generator=g), 5),
        "int_b6_t8_f1_size2": (torch.randint(0, 1000, (6, 8, 1), generator=g), 2),
        "size1_b4_t5_f2": (torch.randn(4, 5, 2, generator=g), 1),
    }
    return out


def np_(x):  # noqa: ANN001, ANN201
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


def same(ref, got, what: str) -> None:  # noqa: ANN001
    ref = np_(ref)
    assert ref.shape == got.shape, (what, ref.shape, got.shape)
    assert ref.dtype == got.dtype, (what, ref.dtype, got.dtype)
    assert np.array_equal(ref, got), what


def main() -> None:
    store: dict[str, np.ndarray] = {}
    for name, (x, size) in cases().items():
        xn = x.numpy()
        B, T = x.shape[:2]
        store[f"{name}/x"] = xn
        store[f"{name}/size"] = np.array(size)

        r = pad_last_sequence(x, size)
        o = V.pad_last_sequence(xn, size)
        for k in ("inputs", "padding_mask"):
            same(r[k], o[k], f"{name} pad_last {k}")
            store[f"{name}/pad_last/{k}"] = np_(r[k])

        r = pad_whole_sequence(x, size)
        o = V.pad_whole_sequence(xn, size)
        for k in ("inputs", "padding_mask"):
            same(r[k], o[k], f"{name} pad_whole {k}")
            store[f"{name}/pad_whole/{k}"] = np_(r[k])

        r = PaddedRollingWindow.apply_all(x, size)
        o = V.padded_rolling_window_apply_all(xn, size)
        for k in ("inputs", "padding_mask"):
            same(r[k], o[k], f"{name} padded_all {k}")
            store[f"{name}/padded_all/{k}"] = np_(r[k])

        r = PaddedRollingWindow.apply_last(x, size)
        o = V.padded_rolling_window_apply_last(xn, size)
        for k in ("inputs", "padding_mask"):
            same(r[k], o[k], f"{name} padded_last {k}")
            store[f"{name}/padded_last/{k}"] = np_(r[k])

        r = RollingWindow.apply_last(x, size)
        same(r, V.rolling_window_apply_last(xn, size), f"{name} rolling_last")
        store[f"{name}/rolling_last"] = np_(r)

        if T >= size:
            r = rolling_window(x, size)
            same(r, V.rolling_window(xn, size), f"{name} rolling_window")
            store[f"{name}/rolling_window"] = np_(r)
            r = RollingWindow.apply_all(x, size)
            same(r, V.rolling_window_apply_all(xn, size), f"{name} rolling_all")
            store[f"{name}/rolling_all"] = np_(r)

        # ViewRequirement on a batch (TensorDict) item, both methods, shift = size - 1
        batch = TensorDict({"obs": x}, batch_size=[B, T])
        for method in ("rolling_window", "padded_rolling_window"):
            vr = ViewRequirement(shift=size - 1, method=method)
            assert vr.drop_size == V.drop_size(size - 1, method)
            if method == "rolling_window" and T < size:
                continue
            r = vr.apply_all("obs", batch)
            o = V.view_apply_all(xn, size - 1, method)
            if isinstance(o, dict):
                for k in ("inputs", "padding_mask"):
                    same(r[k], o[k], f"{name} vr_all {method} {k}")
            else:
                same(r, o, f"{name} vr_all {method}")
            r = vr.apply_last("obs", batch)
            o = V.view_apply_last(xn, size - 1, method)
            if isinstance(o, dict):
                for k in ("inputs", "padding_mask"):
                    same(r[k], o[k], f"{name} vr_last {method} {k}")
            else:
                same(r, o, f"{name} vr_last {method}")

    path = os.path.join(HERE, "views.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {len(store)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
