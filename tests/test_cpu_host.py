"""CPU: the C-ABI library loads and exports everything include/rl8_b200.h declares, and the
host-side logic (hyper-parameter validation, schedules, stop conditions, specs) behaves like
the reference's.  No compute calls: there is no GPU here and the product has no CPU path."""

from __future__ import annotations

import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol() -> None:
    from rl8_b200 import _lib

    header = open(os.path.join(ROOT, "include", "rl8_b200.h")).read()
    declared = set(re.findall(r"\b(rl8_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"librl8_b200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.rl8_abi_version() == _lib.ABI_VERSION
    assert lib.rl8_last_error() == b""


def test_argument_errors_without_touching_the_gpu() -> None:
    from rl8_b200 import _lib

    lib = _lib.load()
    assert lib.rl8_env_step(2, None, None, None, None, 1, 1, None, 16, None) == -1
    assert lib.rl8_gae_scan(None, None, None, None, 4, 4, 1, 4, 0.9, 0.9, 1.0, None, None) == -1
    assert lib.rl8_clip_adam(None, None, None, None, 0, 5.0, 1e-3, 0.9, 0.999, 1e-8, 1, None, None) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-1, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(-3, "x")


def test_product_refuses_to_run_without_cuda() -> None:
    from rl8_b200 import AlgorithmConfig
    from rl8_b200.env import CartPole

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        AlgorithmConfig(num_envs=8, horizon=4).build(CartPole)
    with pytest.raises(RuntimeError):
        CartPole(8, 4, device="cpu")


def test_product_never_imports_the_oracle() -> None:
    pkg = os.path.join(ROOT, "rl8_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def _hp(**over):
    from rl8_b200.data import AlgorithmHparams

    base = dict(
        accumulate_grads=False, clip_param=0.2, device="cuda", dual_clip_param=None, enable_amp=False,
        gae_lambda=0.95, gamma=0.95, horizon=32, horizons_per_env_reset=1, max_grad_norm=5.0,
        normalize_advantages=True, normalize_rewards=True, num_envs=64, num_sgd_iters=4,
        sgd_minibatch_size=2048, shuffle_minibatches=True, target_kl_div=None, vf_clip_param=5.0,
        vf_coeff=1.0,
    )
    base.update(over)
    return AlgorithmHparams(**base)


@pytest.mark.parametrize(
    "bad",
    [dict(clip_param=0.0), dict(clip_param=1.0), dict(dual_clip_param=1.0), dict(gae_lambda=0.0),
     dict(gamma=1.5), dict(horizon=0), dict(horizons_per_env_reset=0), dict(max_grad_norm=0.0),
     dict(num_sgd_iters=0), dict(sgd_minibatch_size=0), dict(vf_clip_param=0.0), dict(vf_coeff=0.0),
     dict(target_kl_div=-1.0),
     dict(target_kl_div=0.1, accumulate_grads=True, sgd_minibatch_size=1024),
     dict(accumulate_grads=True), dict(device="cpu", enable_amp=True)],
)
def test_hparam_validation_rules(bad: dict) -> None:
    """src/rl8/data.py:196-252."""
    with pytest.raises(ValueError):
        _hp(**bad)


def test_early_stop_is_accepted_with_amp() -> None:
    """The one validation rule that differs from src/rl8/data.py: the reference rejects target_kl_div with enable_amp
    because of its GradScaler; enable_amp here is bf16 operands with fp32 accumulation and has no loss scaling."""
    assert _hp(target_kl_div=0.1, enable_amp=True).target_kl_div == 0.1


def test_hparam_properties() -> None:
    hp = _hp(sgd_minibatch_size=512).validate()
    assert hp.num_minibatches == 4 and hp.device_type == "cuda"
    with pytest.raises(ValueError):
        _hp(sgd_minibatch_size=100).validate()
    assert _hp(horizons_per_env_reset=-1).horizons_per_env_reset == -1


def test_schedulers() -> None:
    """src/rl8/schedulers.py: step holds, interp interpolates, no schedule leaves lr alone."""
    from rl8_b200.schedulers import EntropyScheduler, LRScheduler

    e = EntropyScheduler(0.3)
    assert e.coeff == 0.3 and e.step(10**9) == 0.3
    e = EntropyScheduler(0.0, schedule=[(0, 1.0), (100, 0.5), (200, 0.0)], kind="step")
    assert [e.step(c) for c in (0, 99, 100, 150, 200, 10**6)] == [1.0, 1.0, 0.5, 0.5, 0.0, 0.0]
    e = EntropyScheduler(0.0, schedule=[(0, 1.0), (100, 0.0)], kind="interp")
    assert e.step(50) == pytest.approx(0.5) and e.step(1000) == 0.0
    with pytest.raises(ValueError):
        EntropyScheduler(0.0, schedule=[(5, 1.0)])
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=3e-4)
    LRScheduler(opt).step(1000)
    assert opt.param_groups[0]["lr"] == 3e-4
    s = LRScheduler(opt, schedule=[(0, 1e-3), (10, 1e-4)])
    assert opt.param_groups[0]["lr"] == 1e-3
    s.step(10)
    assert opt.param_groups[0]["lr"] == 1e-4


def test_conditions() -> None:
    from rl8_b200.conditions import And, HitsLowerBound, HitsUpperBound, Plateaus, StopsDecreasing, StopsIncreasing

    assert HitsUpperBound("x", 3)({"x": 3}) and not HitsUpperBound("x", 3)({"x": 2.9})
    assert HitsLowerBound("x", 3)({"x": 3}) and not HitsLowerBound("x", 3)({"x": 3.1})
    assert And([HitsUpperBound("x", 1), HitsLowerBound("y", 0)])({"x": 2, "y": -1})
    assert not And([HitsUpperBound("x", 1), HitsLowerBound("y", 0)])({"x": 2, "y": 1})
    pl = Plateaus("x", patience=2, rtol=0.1)
    assert [pl({"x": v}) for v in (1.0, 1.05, 1.06, 2.0, 2.0, 2.0)] == [False, False, True, False, False, True]
    sd = StopsDecreasing("x", patience=2)
    assert [sd({"x": v}) for v in (3, 2, 2.5, 2.1, 1.0)] == [False, False, False, True, False]
    si = StopsIncreasing("x", patience=1)
    assert [si({"x": v}) for v in (1, 2, 2)] == [False, False, True]


def test_specs_and_reduce_stats() -> None:
    from rl8_b200.specs import Categorical, Composite, Unbounded
    from rl8_b200.trainers import reduce_stats

    u, c = Unbounded(5), Categorical(3, shape=torch.Size([1]))
    assert u.zero([4, 2]).shape == (4, 2, 5) and c.zero([4]).dtype == torch.int64
    assert c.space.n == 3 and c.is_in(torch.tensor([[2]])) and not c.is_in(torch.tensor([[3]]))
    assert not u.is_in(torch.zeros(4, 3))
    comp = Composite({"obs": u})
    comp.set("actions", c)
    assert list(comp) == ["obs", "actions"] and comp.to("cpu")["obs"].shape == u.shape
    r = reduce_stats({"a/min": [1, 2], "a/max": [1, 2], "a/mean": [1, 3], "a/std": [3, 4], "env/steps": [5, 5]})
    assert r == {"a/min": 1, "a/max": 2, "a/mean": 2, "a/std": (12.5) ** 0.5, "env/steps": 10}


def test_bench_reference_arm_contract() -> None:
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints ONE JSON line with the
    contract's keys; a tiny workload keeps this in seconds."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(
        [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
         "--num-envs-per-gpu", "256", "--horizon", "8"],
        capture_output=True, text=True, timeout=300, cwd=root,
    )
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "transitions/s"
    assert d["metric"].startswith("env transitions/sec")
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    # the staged unmodified reference (tools/stage_ref.py -> git-ignored oracle/_ref/) when present, else the CPU port
    staged = os.path.isdir(os.path.join(root, "oracle", "_ref", "src", "rl8"))
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["num_envs_per_gpu"] == 256 and d["config"]["horizon"] == 8  # the size that actually ran
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("CartPole")


def test_generic_model_host_logic() -> None:
    """GenericModel (the reference's Model plug-in point, src/rl8/models/_feedforward.py:19-231): default view
    requirement, drop size, ambiguous drop sizes rejected; no device needed."""
    from rl8_b200.models import GenericModel
    from rl8_b200.specs import Categorical, Unbounded
    from rl8_b200.views import ViewRequirement

    class M(GenericModel):
        def forward(self, batch):  # noqa: ANN001, ANN201
            return {}

    m = M(Unbounded(3), Categorical(2, shape=(1,)), width=7)
    assert m.config == {"width": 7}
    assert list(m.view_requirements) == ["obs"] and m.view_requirements["obs"].shift == 0
    assert m.drop_size == 0
    m.validate_view_requirements()
    m.view_requirements["obs"] = ViewRequirement(shift=2, method="rolling_window")
    assert m.drop_size == 2
    m.view_requirements["other"] = ViewRequirement(shift=1, method="padded_rolling_window")
    with pytest.raises(RuntimeError, match="ambiguous"):
        m.validate_view_requirements()
    # shift-0 views are pure torch indexing: they work on CPU tensors too
    m.view_requirements = {"obs": ViewRequirement(shift=0)}
    x = torch.arange(24.0).reshape(2, 4, 3)
    assert torch.equal(m.apply_view_requirements({"obs": x}, kind="last")["obs"], x[:, -1])
    assert torch.equal(m.apply_view_requirements({"obs": x}, kind="all")["obs"], x.reshape(8, 3))
