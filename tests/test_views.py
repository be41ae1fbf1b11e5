"""View requirements: ``rl8_b200.views`` / ``rl8_view_windows`` against the
golden vectors of the unmodified reference (tests/golden/views.npz, made by
tests/golden/generate_views_golden.py from src/rl8/views.py) and against the numpy oracle.

CPU tests pin the oracle to the golden vectors; ``gpu`` tests run the kernel (bit-exact: a view is
a copy) on the golden cases -- including the input shapes of the reference's own
tests/test_views.py -- on both memory layouts of the rollout buffer, and at BASELINE.json sizes
through size-independent properties.
"""

from __future__ import annotations

import os

import numpy as np
import pytest
import torch

from oracle import views_oracle as V

from .conftest import GOLDEN_DIR

Z = np.load(os.path.join(GOLDEN_DIR, "views.npz"), allow_pickle=False)
CASES = sorted({k.split("/")[0] for k in Z.files})
DEV = "cuda"


def case(name: str) -> tuple[np.ndarray, int]:
    return np.array(Z[f"{name}/x"]), int(Z[f"{name}/size"])


def has(name: str, key: str) -> bool:
    return f"{name}/{key}" in Z.files


# ------------------------------------------------------------------------------------------------
# CPU: the oracle reproduces the reference
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name: str) -> None:
    x, size = case(name)
    pairs = {
        "pad_last": V.pad_last_sequence(x, size),
        "pad_whole": V.pad_whole_sequence(x, size),
        "padded_all": V.padded_rolling_window_apply_all(x, size),
        "padded_last": V.padded_rolling_window_apply_last(x, size),
    }
    for fn, got in pairs.items():
        for k in ("inputs", "padding_mask"):
            ref = Z[f"{name}/{fn}/{k}"]
            assert got[k].dtype == ref.dtype and np.array_equal(got[k], ref), (fn, k)
    assert np.array_equal(V.rolling_window_apply_last(x, size), Z[f"{name}/rolling_last"])
    if has(name, "rolling_all"):
        assert np.array_equal(V.rolling_window(x, size), Z[f"{name}/rolling_window"])
        assert np.array_equal(V.rolling_window_apply_all(x, size), Z[f"{name}/rolling_all"])


def test_oracle_known_answers_of_the_reference_tests() -> None:
    """The hand-written expectations of upstream tests/test_views.py:15-60 (arange inputs)."""
    x = np.arange(4, dtype=np.float32).reshape(4, 1)
    out = V.pad_last_sequence(x, 2)
    assert out["inputs"].tolist() == [[0, 0], [0, 1], [0, 2], [0, 3]]
    assert out["padding_mask"].tolist() == [[True, False]] * 4
    x = np.arange(8, dtype=np.float32).reshape(2, 2, 2)
    out = V.pad_last_sequence(x, 2)
    assert np.array_equal(out["inputs"], x) and not out["padding_mask"].any()
    x = np.arange(8, dtype=np.float32).reshape(2, 4)
    assert V.rolling_window_apply_all(x, 2).tolist() == [[0, 1], [1, 2], [2, 3], [4, 5], [5, 6], [6, 7]]
    out = V.padded_rolling_window_apply_all(x, 2)
    assert out["inputs"].tolist() == [[0, 0], [0, 1], [1, 2], [2, 3], [0, 4], [4, 5], [5, 6], [6, 7]]
    assert out["padding_mask"][:, 0].tolist() == [True, False, False, False] * 2
    assert V.drop_size(3, "rolling_window") == 3 and V.drop_size(3, "padded_rolling_window") == 0


def test_view_requirement_host_logic() -> None:
    """Argument validation and drop sizes need no device (src/rl8/views.py:351-363, 447-453)."""
    from rl8_b200.views import PaddedRollingWindow, RollingWindow, ViewRequirement

    with pytest.raises(ValueError):
        ViewRequirement(shift=-1)
    assert ViewRequirement(shift=3, method="rolling_window").drop_size == 3
    assert ViewRequirement(shift=3).drop_size == 0
    assert ViewRequirement(shift=3).method is PaddedRollingWindow
    assert ViewRequirement(shift=1, method="rolling_window").method is RollingWindow
    with pytest.raises(RuntimeError, match="no CPU path"):
        PaddedRollingWindow.apply_all(torch.zeros(2, 3, 1), 2)


# ------------------------------------------------------------------------------------------------
# GPU: the kernel
# ------------------------------------------------------------------------------------------------


def _eq(got: torch.Tensor, ref: np.ndarray, what: str) -> None:
    g = got.cpu().numpy()
    assert g.shape == ref.shape, (what, g.shape, ref.shape)
    assert g.dtype == ref.dtype, (what, g.dtype, ref.dtype)
    assert np.array_equal(g, ref), what


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(name: str) -> None:
    from rl8_b200 import views as W

    xn, size = case(name)
    x = torch.from_numpy(xn).to(DEV)
    for fn, got in {
        "pad_last": W.pad_last_sequence(x, size),
        "pad_whole": W.pad_whole_sequence(x, size),
        "padded_all": W.PaddedRollingWindow.apply_all(x, size),
        "padded_last": W.PaddedRollingWindow.apply_last(x, size),
    }.items():
        for k in ("inputs", "padding_mask"):
            _eq(got[k], Z[f"{name}/{fn}/{k}"], f"{fn}/{k}")
    _eq(W.RollingWindow.apply_last(x, size), Z[f"{name}/rolling_last"], "rolling_last")
    if has(name, "rolling_all"):
        _eq(W.rolling_window(x, size), Z[f"{name}/rolling_window"], "rolling_window")
        _eq(W.RollingWindow.apply_all(x, size), Z[f"{name}/rolling_all"], "rolling_all")
    else:
        with pytest.raises(RuntimeError):
            W.RollingWindow.apply_all(x, size)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["rolling_window", "padded_rolling_window"])
@pytest.mark.parametrize("shift", [0, 1, 3])
def test_view_requirement_on_batches(method: str, shift: int) -> None:
    """``ViewRequirement.apply_all / apply_last`` on a (nested) batch, tuple keys included
    (upstream tests/test_views.py:434-505 use shift 0; shifted views follow :365-445)."""
    from rl8_b200.views import ViewRequirement

    g = torch.Generator().manual_seed(3)
    obs = torch.randn(6, 9, 4, generator=g)
    prices = torch.randint(0, 50, (6, 9, 2), generator=g)
    batch = {"obs": obs.to(DEV), "nested": {"prices": prices.to(DEV)}}
    vr = ViewRequirement(shift=shift, method=method)
    for key, x in (("obs", obs), (("nested", "prices"), prices)):
        for kind, ora in (("all", V.view_apply_all), ("last", V.view_apply_last)):
            got = getattr(vr, f"apply_{kind}")(key, batch)
            ref = ora(x.numpy(), shift, method)
            if isinstance(ref, dict):
                for k in ("inputs", "padding_mask"):
                    _eq(got[k], ref[k], f"{key}/{kind}/{k}")
            else:
                _eq(got, ref, f"{key}/{kind}")
    # a mapping item maps leaf-wise (TensorDict.apply in the reference)
    got = vr.apply_all("nested", batch)["prices"]
    ref = V.view_apply_all(prices.numpy(), shift, method)
    if isinstance(ref, dict):
        _eq(got["inputs"], ref["inputs"], "nested inputs")
    else:
        _eq(got, ref, "nested")


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,D,size", [(64, 8, 5, 4), (1000, 32, 5, 3), (37, 5, 3, 7), (5, 64, 1, 64)])
def test_views_on_the_rollout_buffer_layout(N: int, T: int, D: int, size: int) -> None:
    """``buffer["obs"]`` is an [N, T+1, D] strided view of horizon-major [T+1][D][N] memory and
    ``buffer["actions"]`` an int64 one: the kernel reads both in place (no contiguous copy)."""
    from rl8_b200 import views as W
    from rl8_b200.buffer import RolloutBuffer
    from rl8_b200.specs import Categorical, Composite, Unbounded

    spec = Composite({
        "obs": Unbounded(D, device=DEV), "rewards": Unbounded(1, device=DEV),
        "actions": Categorical(3, shape=(1,), device=DEV), "logp": Unbounded(1, device=DEV),
        "values": Unbounded(1, device=DEV), "advantages": Unbounded(1, device=DEV),
        "returns": Unbounded(1, device=DEV),
    })
    buf = RolloutBuffer(spec, N, T, DEV)
    g = torch.Generator().manual_seed(N)
    obs = torch.randn(N, T + 1, D, generator=g)
    act = torch.randint(0, 3, (N, T + 1, 1), generator=g)
    buf["obs"].copy_(obs.to(DEV))
    buf["actions"].copy_(act.to(DEV))
    assert buf["obs"].stride(0) == 1  # horizon-major underneath
    for x_dev, x in ((buf["obs"], obs), (buf["actions"], act)):
        got = W.PaddedRollingWindow.apply_all(x_dev, size)
        ref = V.padded_rolling_window_apply_all(x.numpy(), size)
        _eq(got["inputs"], ref["inputs"], "padded inputs")
        _eq(got["padding_mask"], ref["padding_mask"], "padded mask")
        if T + 1 >= size:
            _eq(W.RollingWindow.apply_all(x_dev, size), V.rolling_window_apply_all(x.numpy(), size), "rolling")
        last = W.PaddedRollingWindow.apply_last(x_dev, size)
        ref = V.pad_last_sequence(x.numpy(), size)
        _eq(last["inputs"], ref["inputs"], "last inputs")
        _eq(last["padding_mask"], ref["padding_mask"], "last mask")


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["horizon_major", "env_major"])
def test_views_at_baseline_size_properties(layout: str) -> None:
    """CartPole N=65536, T=32 obs (BASELINE configs[1]) with shift 3: every window element is the
    item element it names (checked on a random sample against direct indexing), padding is exactly
    the first ``size - 1 - t`` slots, and the checksum equals the multiplicity-weighted item sum."""
    from rl8_b200 import views as W

    N, T1, D, size = 65536, 33, 5, 4
    g = torch.Generator(device=DEV).manual_seed(7)
    if layout == "horizon_major":
        x = torch.randn(T1, D, N, generator=g, device=DEV).permute(2, 0, 1)
    else:
        x = torch.randn(N, T1, D, generator=g, device=DEV)
    out = W.PaddedRollingWindow.apply_all(x, size)
    inputs, mask = out["inputs"], out["padding_mask"]
    assert inputs.shape == (N * T1, size, D) and mask.shape == (N * T1, size)
    # padding mask: window (b, t) slot s is padding iff t + s - (size - 1) < 0
    t = torch.arange(T1, device=DEV).view(1, T1, 1)
    s = torch.arange(size, device=DEV).view(1, 1, size)
    want_mask = (t + s - (size - 1) < 0).expand(N, T1, size).reshape(N * T1, size)
    assert torch.equal(mask, want_mask)
    assert float(inputs[mask].abs().sum()) == 0.0
    # random sample against direct indexing
    idx = torch.randint(0, N * T1, (200_000,), device=DEV)
    ss = torch.randint(0, size, (200_000,), device=DEV)
    b, tt = idx // T1, idx % T1
    src_t = tt + ss - (size - 1)
    ok = src_t >= 0
    got = inputs[idx, ss]
    ref = torch.where(ok.unsqueeze(1), x[b, src_t.clamp(min=0)], torch.zeros((), device=DEV))
    assert torch.equal(got, ref)
    # checksum: item step t appears in min(size, T1 - t) windows
    mult = torch.minimum(torch.full((T1,), size, device=DEV), T1 - torch.arange(T1, device=DEV)).double()
    want = (x.double().sum(dim=(0, 2)) * mult).sum()
    assert float(inputs.double().sum()) == pytest.approx(float(want), rel=1e-9, abs=1e-6)
    # unpadded rolling windows are the padded ones without the first size - 1 windows of every env
    roll = W.RollingWindow.apply_all(x, size)
    assert torch.equal(roll.view(N, T1 - size + 1, size, D), inputs.view(N, T1, size, D)[:, size - 1:])
