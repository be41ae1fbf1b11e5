"""Generate golden vectors from the UNMODIFIED upstream reference.  Container-only tool.

Run in the build container (the reference does not exist on the GPU box)::

    PYTHONPATH=oracle/refshim:/root/reference/src:/root/reference TORCHDYNAMO_DISABLE=1 \
        python tests/golden/generate_golden.py

For every case this script

1. builds the reference ``Algorithm`` through its public API
   (``AlgorithmConfig(...).build(env_cls)``),
2. injects the env's initial state and the per-step sampling noise through the
   reference's own plug-in points (an ``Env`` subclass and a ``distribution_cls``
   subclass; SURVEY.md §8c "RNG injection"),
3. runs ``collect()`` and ``step()`` (the reference's code, torch fp32, eager CPU),
4. replays the same inputs through ``oracle/ppo_oracle.py`` and asserts agreement, and
5. stores inputs + reference outputs in ``tests/golden/<case>.npz``.

It also regenerates the known-answer vectors listed in SURVEY.md §8c
(``kat.npz``) from the reference's functions.
"""

from __future__ import annotations

import os
import sys
from typing import Any

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

from examples.cartpole.env import CartPole  # noqa: E402  (upstream)
from examples.cartpole.env import step as ref_cartpole_step  # noqa: E402
from examples.mountain_car.env import MountainCar  # noqa: E402
from examples.mountain_car.env import step as ref_mountain_car_step  # noqa: E402
from examples.pendulum.env import Pendulum  # noqa: E402
from examples.pendulum.env import step as ref_pendulum_step  # noqa: E402
from rl8 import AlgorithmConfig  # noqa: E402
from rl8.data import DataKeys  # noqa: E402
from rl8.distributions import Categorical, Normal, SquashedNormal  # noqa: E402
from rl8.env import ContinuousDummyEnv, DiscreteDummyEnv  # noqa: E402
from rl8.nn.functional import generalized_advantage_estimate, ppo_losses  # noqa: E402
from tensordict import TensorDict  # noqa: E402  (refshim)

from oracle import ppo_oracle as O  # noqa: E402

SUBSAMPLE = 16  # stride used when storing gradients / final params


# ---------------------------------------------------------------------------------
# Injection through the reference's plug-in points
# ---------------------------------------------------------------------------------


class _NoiseQueue:
    items: list[torch.Tensor] = []


class InjCategorical(Categorical):
    def sample(self) -> torch.Tensor:
        if not _NoiseQueue.items:  # build() -> validate(): not part of the vectors
            return super().sample()
        q = _NoiseQueue.items.pop(0)
        return torch.argmax(self.dist.probs / q, dim=-1)


class InjNormal(Normal):
    def sample(self) -> torch.Tensor:
        if not _NoiseQueue.items:
            return super().sample()
        z = _NoiseQueue.items.pop(0)
        return z * self.dist.scale + self.dist.loc


class InjSquashedNormal(SquashedNormal):
    def sample(self) -> torch.Tensor:
        if not _NoiseQueue.items:
            return super().sample()
        z = _NoiseQueue.items.pop(0)
        return (z * self.dist.scale + self.dist.loc).tanh()


def injected_env(env_cls: type, states: list[torch.Tensor]) -> type:
    """Subclass whose ``reset`` installs a pre-drawn state (after calling the parent's
    ``reset`` so configs are handled by the reference's own code)."""

    class Injected(env_cls):  # type: ignore[misc, valid-type]
        def reset(self, *, config: None | dict[str, Any] = None) -> torch.Tensor:
            obs = super().reset(config=config)
            if not states:
                return obs
            self.state = states.pop(0).clone()
            if env_cls in (DiscreteDummyEnv, ContinuousDummyEnv):
                return self.state
            if env_cls is CartPole:
                x, xd, th, thd = self.state
                return torch.vstack((x, xd, torch.cos(th), torch.sin(th), thd)).T
            if env_cls is MountainCar:
                return self.state.T
            th, thd = self.state
            return torch.vstack((torch.cos(th), torch.sin(th), thd)).T

    Injected.__name__ = env_cls.__name__
    return Injected


ENV_INFO = {
    # oracle name: (reference class, state shape fn, reset sampler)
    "discrete_dummy": DiscreteDummyEnv,
    "continuous_dummy": ContinuousDummyEnv,
    "cartpole": CartPole,
    "mountain_car": MountainCar,
    "pendulum": Pendulum,
}

DIST_INFO = {
    "categorical": InjCategorical,
    "normal": InjNormal,
    "squashed_normal": InjSquashedNormal,
}


class KinkTooClose(RuntimeError):
    pass


def td_to_np(buf: TensorDict) -> dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy().copy() for k, v in buf.items()}


def assert_same(name: str, a: torch.Tensor, b: torch.Tensor, exact: bool = True) -> None:
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if exact:
        if not torch.equal(a, b):
            diff = (a.double() - b.double()).abs().max()
            raise AssertionError(f"oracle != reference for {name}: max abs diff {diff}")
    else:
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=lambda m: f"{name}: {m}")


def min_preactivation(params: dict[str, torch.Tensor], obs: torch.Tensor) -> float:
    """Smallest |pre-activation| of any ReLU of both networks over ``obs [rows, D]`` (fp64).  A ReLU whose input is
    within an implementation's rounding error of zero makes the first-step gradient discontinuous in that error:
    two correct fp32 implementations (this CPU run, the same code on a GPU, the kernels) can land on different
    sides.  Cases recorded for elementwise gradient comparison keep a margin (``min_kink_margin``)."""
    worst = float("inf")
    for pre in ("feature_model", "latent_model", "vf_model"):
        if f"{pre}.0.0.weight" not in params:
            continue
        x = obs.double()
        z1 = x @ params[f"{pre}.0.0.weight"].double().T + params[f"{pre}.0.0.bias"].double()
        z2 = torch.relu(z1) @ params[f"{pre}.0.2.weight"].double().T + params[f"{pre}.0.2.bias"].double()
        worst = min(worst, float(z1.abs().min()), float(z2.abs().min()))
    return worst


def _run_case(
    name: str,
    env_name: str,
    dist_name: str,
    *,
    N: int,
    T: int,
    seed: int,
    rounds: int = 1,
    horizons_per_env_reset: int = 1,
    min_kink_margin: float = 0.0,
    **algo_kwargs: Any,
) -> None:
    print(f"== {name} (seed {seed})")
    torch.manual_seed(seed)
    ref_env_cls = ENV_INFO[env_name]
    states: list[torch.Tensor] = []
    env_cls = injected_env(ref_env_cls, states)
    algo_kwargs = dict(algo_kwargs)
    algo_kwargs.setdefault("shuffle_minibatches", False)
    algo = AlgorithmConfig(
        num_envs=N,
        horizon=T,
        horizons_per_env_reset=horizons_per_env_reset,
        distribution_cls=DIST_INFO[dist_name],
        device="cpu",
        **algo_kwargs,
    ).build(env_cls)  # build() -> validate() consumed one (non-injected) sample; fine.

    # Parameters (inputs).
    params0 = {k: v.detach().clone() for k, v in algo.policy.model.state_dict().items()}

    # Oracle twin.
    o_env = O.OracleEnv(env_name, N)
    o_p = {k: v.clone() for k, v in params0.items()}
    o_dist = O.Dist(dist_name)
    o_buf = O.new_buffer(N, T, o_env.obs_dim, o_env.action_kind)
    o_opt: dict[str, Any] = {}

    A = o_env.num_actions
    out: dict[str, np.ndarray] = {f"param0/{k}": v.numpy() for k, v in params0.items()}
    meta = dict(
        env=env_name,
        dist=dist_name,
        N=N,
        T=T,
        rounds=rounds,
        horizons_per_env_reset=horizons_per_env_reset,
        subsample=SUBSAMPLE,
        **{k: v for k, v in algo_kwargs.items()},
    )

    gen = torch.Generator().manual_seed(seed + 1000)
    for rnd in range(rounds):
        will_reset = (algo.state.horizons % horizons_per_env_reset) == 0
        # Draw the injected inputs.
        if will_reset:
            draw = ref_env_cls(N, T, device="cpu").reset()  # reference's own reset dist
            del draw
            tmp_env = ref_env_cls(N, T, device="cpu")
            tmp_env.reset()
            state0 = tmp_env.state.detach().clone()
            states.append(state0.clone())
            out[f"r{rnd}/state0"] = state0.numpy()
        if dist_name == "categorical":
            noise = torch.empty(T, N, 1, A).exponential_(1, generator=gen)
        else:
            noise = torch.randn(T, N, 1, generator=gen)
        out[f"r{rnd}/noise"] = noise.numpy()
        _NoiseQueue.items = [noise[t] for t in range(T)]

        # ---- reference collect -------------------------------------------------
        cstats = algo.collect()
        assert not _NoiseQueue.items
        ref_buf = td_to_np(algo.buffer)
        for k, v in ref_buf.items():
            out[f"r{rnd}/collect/{k}"] = v
        for k, v in cstats.items():
            if not k.startswith("profiling"):
                out[f"r{rnd}/collect_stats/{k}"] = np.float64(v)
        out[f"r{rnd}/reward_scale"] = np.float64(algo.state.reward_scale)

        # ---- oracle collect ----------------------------------------------------
        o_stats = O.collect(
            o_p,
            o_env,
            o_buf,
            o_dist,
            noise,
            gamma=algo.hparams.gamma,
            reset=will_reset,
            reset_state=state0 if will_reset else None,
            normalize_rewards=algo.hparams.normalize_rewards,
        )
        margin = min_preactivation(params0 if rnd == 0 else {k: v.detach() for k, v in algo.policy.model.state_dict().items()},
                                   algo.buffer[DataKeys.OBS][:, :-1].reshape(N * T, -1))
        print(f"   smallest |ReLU input| over the buffer: {margin:.3e}")
        if margin < min_kink_margin:
            raise KinkTooClose(f"{name}: a ReLU input is {margin:.2e} from zero (< {min_kink_margin:g}); use another seed")
        meta[f"min_relu_input_r{rnd}"] = margin
        for k in ("obs", "rewards", "actions", "logp", "values", "reversed_discounted_returns"):
            if k in ref_buf:  # no reversed discounted returns without reward normalisation (:252-256)
                assert_same(f"{name}/r{rnd}/collect/{k}", o_buf[k], algo.buffer[k])
        for k, v in o_stats.items():
            ref_v = algo.state.reward_scale if k == "reward_scale" else cstats[k]
            assert v == ref_v, (k, v, ref_v)

        # ---- reference GAE on a copy (stage vector) ------------------------------
        gae_in = TensorDict(
            {
                DataKeys.REWARDS: algo.buffer[DataKeys.REWARDS].clone(),
                DataKeys.VALUES: algo.buffer[DataKeys.VALUES].clone(),
            },
            batch_size=[N, T + 1],
        )
        gae_out = generalized_advantage_estimate(
            gae_in,
            gae_lambda=algo.hparams.gae_lambda,
            gamma=algo.hparams.gamma,
            inplace=True,
            normalize_advantages=algo.hparams.normalize_advantages,
            return_returns=True,
            reward_scale=algo.state.reward_scale,
        )
        out[f"r{rnd}/gae/rewards_scaled"] = gae_out[DataKeys.REWARDS].numpy().copy()
        out[f"r{rnd}/gae/advantages"] = gae_out[DataKeys.ADVANTAGES].numpy().copy()
        out[f"r{rnd}/gae/returns"] = gae_out[DataKeys.RETURNS].numpy().copy()
        o_r, o_adv, o_ret = O.gae(
            o_buf["rewards"],
            o_buf["values"],
            gamma=algo.hparams.gamma,
            gae_lambda=algo.hparams.gae_lambda,
            reward_scale=algo.state.reward_scale,
            normalize_advantages=algo.hparams.normalize_advantages,
        )
        assert_same(f"{name}/gae/r", o_r, gae_out[DataKeys.REWARDS])
        assert_same(f"{name}/gae/adv", o_adv, gae_out[DataKeys.ADVANTAGES])
        assert_same(f"{name}/gae/ret", o_ret, gae_out[DataKeys.RETURNS])

        # ---- reference step; record the first optimizer step's gradients -----------
        grads_first: dict[str, torch.Tensor] = {}
        orig_clip = torch.nn.utils.clip_grad_norm_

        def recording_clip(parameters, max_norm, *a, **kw):
            parameters = list(parameters)
            if not grads_first:
                for (k, _), prm in zip(algo.policy.model.named_parameters(), parameters):
                    grads_first[k] = prm.grad.detach().clone()
            return orig_clip(parameters, max_norm, *a, **kw)

        torch.nn.utils.clip_grad_norm_ = recording_clip
        try:
            sstats = algo.step()
        finally:
            torch.nn.utils.clip_grad_norm_ = orig_clip
        for k, v in sstats.items():
            if not k.startswith("profiling"):
                out[f"r{rnd}/step_stats/{k}"] = np.float64(v)
        params1 = {k: v.detach().clone() for k, v in algo.policy.model.state_dict().items()}
        for k, v in grads_first.items():
            out[f"r{rnd}/grad_first/{k}"] = v.flatten()[::SUBSAMPLE].numpy().copy()
            out[f"r{rnd}/grad_first_norm/{k}"] = np.float64(v.double().norm())
        for k, v in params1.items():
            out[f"r{rnd}/param1/{k}"] = v.flatten()[::SUBSAMPLE].numpy().copy()
            out[f"r{rnd}/param1_norm/{k}"] = np.float64(v.double().norm())

        # ---- oracle step ---------------------------------------------------------
        o_grads: dict[str, torch.Tensor] = {}

        def hook(g: dict[str, torch.Tensor]) -> None:
            if not o_grads:
                o_grads.update(g)

        o_sstats = O.step(
            o_p,
            o_buf,
            o_dist,
            o_opt,
            reward_scale=algo.state.reward_scale,
            gamma=algo.hparams.gamma,
            gae_lambda=algo.hparams.gae_lambda,
            normalize_advantages=algo.hparams.normalize_advantages,
            sgd_minibatch_size=algo.hparams.sgd_minibatch_size,
            num_sgd_iters=algo.hparams.num_sgd_iters,
            shuffle=False,
            accumulate_grads=algo.hparams.accumulate_grads,
            clip_param=algo.hparams.clip_param,
            dual_clip_param=algo.hparams.dual_clip_param,
            entropy_coeff=algo.entropy_scheduler.coeff,
            vf_clip_param=algo.hparams.vf_clip_param,
            vf_coeff=algo.hparams.vf_coeff,
            target_kl_div=algo.hparams.target_kl_div,
            max_grad_norm=algo.hparams.max_grad_norm,
            grad_hook=hook,
        )
        for k in grads_first:
            assert_same(f"{name}/r{rnd}/grad/{k}", o_grads[k], grads_first[k])
        for k in params1:
            assert_same(f"{name}/r{rnd}/param1/{k}", o_p[k], params1[k])
        for k, v in o_sstats.items():
            assert v == sstats[k], (k, v, sstats[k])
        print(f"   round {rnd}: collect+gae+step oracle == reference (bit-exact)")
        print("   ", {k: round(float(v), 6) for k, v in sstats.items() if "profiling" not in k})

    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


# ---------------------------------------------------------------------------------
# Known-answer vectors from the reference's own functions
# ---------------------------------------------------------------------------------


def kat() -> None:
    print("== kat")
    out: dict[str, np.ndarray] = {}

    # Env steps on hand-picked states (post-step state, obs, reward).
    cp_in = torch.tensor(
        [[0.01, -0.5, 2.0], [-0.02, 0.3, -1.0], [0.03, -0.2, 3.0], [0.04, 0.1, -2.5]]
    )
    cp_a = torch.tensor([[0], [1], [2]])
    from examples.cartpole.env import CartPoleConfig
    from dataclasses import asdict

    st, obs, r = ref_cartpole_step(*cp_in, cp_a, **asdict(CartPoleConfig()))
    out["cartpole/state_in"], out["cartpole/action"] = cp_in.numpy(), cp_a.numpy()
    out["cartpole/state_out"], out["cartpole/obs"], out["cartpole/reward"] = (
        st.numpy(),
        obs.contiguous().numpy(),
        r.numpy(),
    )
    o_st, o_obs, o_r = O.cartpole_step(cp_in.clone(), cp_a)
    assert_same("kat/cartpole/state", o_st, st)
    assert_same("kat/cartpole/obs", o_obs, obs)
    assert_same("kat/cartpole/reward", o_r, r)

    pd_in = torch.tensor([[0.5, -3.0, 3.5], [0.1, -7.9, 7.99]])
    pd_a = torch.tensor([[0.3], [-5.0], [2.0]])
    st, obs, r = ref_pendulum_step(pd_in[0].clone(), pd_in[1].clone(), pd_a)
    out["pendulum/state_in"], out["pendulum/action"] = pd_in.numpy(), pd_a.numpy()
    out["pendulum/state_out"], out["pendulum/obs"], out["pendulum/reward"] = (
        st.numpy(),
        obs.contiguous().numpy(),
        r.reshape(-1, 1).numpy(),
    )
    o_st, o_obs, o_r = O.pendulum_step(pd_in.clone(), pd_a)
    assert_same("kat/pendulum/state", o_st, st)
    assert_same("kat/pendulum/obs", o_obs, obs)
    assert_same("kat/pendulum/reward", o_r, r.reshape(-1, 1))

    mc_in = torch.tensor([[-0.5, -1.2, 0.55], [0.0, -0.01, 0.069]])
    mc_a = torch.tensor([[2], [0], [2]])
    work = mc_in.clone()
    st, obs, r = ref_mountain_car_step(work[0], work[1], mc_a)
    out["mountain_car/state_in"], out["mountain_car/action"] = mc_in.numpy(), mc_a.numpy()
    out["mountain_car/state_out"], out["mountain_car/reward"] = (
        st.numpy(),
        r.reshape(-1, 1).numpy(),
    )
    o_st, _, o_r = O.mountain_car_step(mc_in.clone(), mc_a)
    assert_same("kat/mountain_car/state", o_st, st)
    assert_same("kat/mountain_car/reward", o_r, r.reshape(-1, 1))

    # GAE.
    rew = torch.tensor([[1.0, 2.0, 3.0, 0.0], [0.5, -1.0, 2.0, 0.0]]).unsqueeze(-1)
    val = torch.tensor([[0.1, 0.2, 0.3, 0.4], [1.0, -1.0, 0.5, 2.0]]).unsqueeze(-1)
    for norm in (False, True):
        td = TensorDict(
            {DataKeys.REWARDS: rew.clone(), DataKeys.VALUES: val.clone()}, batch_size=[2, 4]
        )
        g = generalized_advantage_estimate(
            td, gae_lambda=0.95, gamma=0.95, inplace=False, normalize_advantages=norm,
            return_returns=True, reward_scale=2.0,
        )
        tag = "norm" if norm else "raw"
        out[f"gae/{tag}/advantages"] = g[DataKeys.ADVANTAGES].numpy().copy()
        out[f"gae/{tag}/returns"] = g[DataKeys.RETURNS].numpy().copy()
        _, o_adv, o_ret = O.gae(
            rew, val, gamma=0.95, gae_lambda=0.95, reward_scale=2.0, normalize_advantages=norm
        )
        assert_same(f"kat/gae/{tag}/adv", o_adv, g[DataKeys.ADVANTAGES])
        assert_same(f"kat/gae/{tag}/ret", o_ret, g[DataKeys.RETURNS])
    out["gae/rewards"], out["gae/values"] = rew.numpy(), val.numpy()

    # PPO loss.
    logits = torch.tensor([[0.1, 0.2, -0.3], [1.0, -1.0, 0.0], [0.0, 0.0, 0.0], [2.0, 1.0, 0.0]])
    logits = logits.reshape(4, 1, 3)
    acts = torch.tensor([[0], [1], [2], [0]])
    logp_old = torch.tensor([[-1.0], [-2.0], [-1.2], [-0.3]])
    adv = torch.tensor([[1.0], [-0.5], [0.2], [-2.0]])
    ret = torch.tensor([[0.5], [3.0], [-0.2], [10.0]])
    vals = torch.tensor([[0.0], [0.5], [0.1], [1.0]])
    dist = Categorical(TensorDict({"logits": logits}, batch_size=[4]), None)
    for dual in (None, 5.0):
        losses = ppo_losses(
            TensorDict(
                {
                    DataKeys.ACTIONS: acts,
                    DataKeys.LOGP: logp_old,
                    DataKeys.ADVANTAGES: adv,
                    DataKeys.RETURNS: ret,
                },
                batch_size=[4],
            ),
            TensorDict({DataKeys.VALUES: vals}, batch_size=[4]),
            dist,
            clip_param=0.2,
            dual_clip_param=dual,
            entropy_coeff=0.01,
            vf_clip_param=5.0,
            vf_coeff=1.0,
        )
        tag = "dual" if dual else "nodual"
        o_l = O.ppo_losses(
            O.categorical_logp(logits, acts), vals, O.categorical_entropy(logits), logp_old, adv,
            ret, clip_param=0.2, dual_clip_param=dual, entropy_coeff=0.01, vf_clip_param=5.0,
        )
        for k in ("entropy", "policy", "vf", "total"):
            out[f"ppo/{tag}/{k}"] = np.float32(losses[k])
            assert_same(f"kat/ppo/{tag}/{k}", o_l[k], losses[k].reshape(()))
    out["ppo/logits"], out["ppo/actions"], out["ppo/logp_old"] = (
        logits.numpy(), acts.numpy(), logp_old.numpy(),
    )
    out["ppo/advantages"], out["ppo/returns"], out["ppo/values"] = (
        adv.numpy(), ret.numpy(), vals.numpy(),
    )
    out["ppo/logp_new"] = dist.logp(acts).numpy()

    # Distributions on a spread of inputs.
    g = torch.Generator().manual_seed(7)
    mean = torch.randn(256, 1, generator=g)
    log_std = torch.tanh(torch.randn(256, 1, generator=g))
    z = torch.randn(256, 1, generator=g)
    feats = TensorDict({"mean": mean, "log_std": log_std}, batch_size=[256])
    nd, sd = Normal(feats, None), SquashedNormal(feats, None)
    x = z * nd.dist.scale + nd.dist.loc
    xs = x.tanh()
    xs[0, 0], xs[1, 0] = 0.99999994, -1.0  # clamp edge cases
    out["dist/mean"], out["dist/log_std"], out["dist/z"] = mean.numpy(), log_std.numpy(), z.numpy()
    out["dist/normal_sample"] = x.numpy()
    out["dist/normal_logp"] = nd.logp(x).numpy()
    out["dist/normal_entropy"] = nd.entropy().numpy()
    out["dist/squashed_x"] = xs.numpy()
    out["dist/squashed_logp"] = sd.logp(xs).numpy()
    assert_same("kat/normal_sample", O.normal_sample(mean, log_std, z), x)
    assert_same("kat/normal_logp", O.normal_logp(mean, log_std, x), nd.logp(x))
    assert_same("kat/normal_entropy", O.normal_entropy(log_std), nd.entropy())
    assert_same("kat/squashed_logp", O.squashed_logp(mean, log_std, xs), sd.logp(xs))
    # torch.normal(loc, scale) == z*scale + loc with the same generator stream.
    g1, g2 = torch.Generator().manual_seed(11), torch.Generator().manual_seed(11)
    t_norm = torch.normal(mean, torch.exp(log_std), generator=g1)
    z2 = torch.randn(256, 1, generator=g2)
    print("   torch.normal == z*s+m bit-exact:", torch.equal(t_norm, z2 * torch.exp(log_std) + mean))

    lg = torch.randn(512, 1, 3, generator=g) * 2
    q = torch.empty(512, 1, 3).exponential_(1, generator=g)
    cd = Categorical(TensorDict({"logits": lg}, batch_size=[512]), None)
    a = torch.argmax(cd.dist.probs / q, dim=-1)
    out["dist/logits"], out["dist/q"], out["dist/cat_sample"] = lg.numpy(), q.numpy(), a.numpy()
    out["dist/cat_logp"] = cd.logp(a).numpy()
    out["dist/cat_entropy"] = cd.entropy().numpy()
    out["dist/cat_mode"] = cd.deterministic_sample().numpy()
    assert_same("kat/cat_sample", O.categorical_sample(lg, q), a)
    assert_same("kat/cat_logp", O.categorical_logp(lg, a), cd.logp(a))
    assert_same("kat/cat_entropy", O.categorical_entropy(lg), cd.entropy())
    assert_same("kat/cat_mode", O.categorical_mode(lg), cd.deterministic_sample())
    # torch's own Categorical.sample() == argmax(probs / q) with the same generator stream.
    torch.manual_seed(123)
    s_ref = cd.sample()
    torch.manual_seed(123)
    q_same = torch.empty(512, 3).exponential_(1)
    s_inj = torch.argmax(cd.dist.probs.reshape(512, 3) / q_same, dim=-1).reshape(512, 1)
    assert torch.equal(s_ref, s_inj), "multinomial != argmax(p/q)"
    print("   Categorical.sample == argmax(probs/q) bit-exact: True")

    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)


def main() -> None:
    torch.set_num_threads(1)
    only = set(sys.argv[1:])  # case names to (re)generate; none = everything

    def run_case(name: str, *a: Any, **kw: Any) -> None:
        if not only or name in only:
            _run_case(name, *a, **kw)

    if not only or "kat" in only:
        kat()
    run_case("ff_discrete_dummy", "discrete_dummy", "categorical", N=64, T=8, seed=1,
             entropy_coeff=0.01, num_sgd_iters=2)
    run_case("ff_continuous_dummy_normal", "continuous_dummy", "normal", N=64, T=8, seed=2,
             entropy_coeff=0.01, num_sgd_iters=2, sgd_minibatch_size=128)
    run_case("ff_continuous_dummy_squashed", "continuous_dummy", "squashed_normal", N=64, T=8,
             seed=3, num_sgd_iters=2, sgd_minibatch_size=128, accumulate_grads=True,
             dual_clip_param=5.0)
    run_case("ff_cartpole", "cartpole", "categorical", N=64, T=16, seed=4, rounds=2,
             horizons_per_env_reset=2)
    run_case("ff_mountain_car", "mountain_car", "categorical", N=64, T=16, seed=5,
             entropy_coeff=0.005, num_sgd_iters=3, sgd_minibatch_size=256, target_kl_div=0.5)
    run_case("ff_pendulum_squashed", "pendulum", "squashed_normal", N=64, T=16, seed=6,
             num_sgd_iters=2)
    run_case("ff_pendulum_normal", "pendulum", "normal", N=64, T=16, seed=7,
             num_sgd_iters=2, entropy_coeff=0.01, normalize_advantages=False)
    # round 2: rows that cross the 128-row CTA halves / 256-row pair tiles of the tensor-core kernels (ragged
    # minibatches of 2.5 tiles), the reward-normalisation switch, and an early stop that really triggers
    # (the first minibatch has KL = 0, the second one exceeds 1.5 * target after one Adam step)
    run_case("ff_cartpole_n320", "cartpole", "categorical", N=320, T=8, seed=8, num_sgd_iters=2,
             sgd_minibatch_size=640, entropy_coeff=0.01)
    run_case("ff_pendulum_n300_raw_rewards", "pendulum", "normal", N=300, T=8, seed=9, num_sgd_iters=2,
             normalize_rewards=False)
    for seed in range(10, 40):  # first seed whose ReLU inputs all keep 1e-4 from zero (|obs| is up to 100 here)
        try:
            run_case("ff_discrete_dummy_early_stop", "discrete_dummy", "categorical", N=64, T=8, seed=seed,
                     num_sgd_iters=4, sgd_minibatch_size=256, target_kl_div=1e-7, min_kink_margin=1e-4)
            break
        except KinkTooClose as e:
            print("   ", e)


if __name__ == "__main__":
    main()
