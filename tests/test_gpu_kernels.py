"""GPU: stage-wise parity of every kernel against the CPU oracle (through the C-ABI).

Integer / index results (discrete actions, reset cadence) must be bit-exact; floating point
must agree within RTOL = 1e-5 (north_star tolerance) plus a small absolute floor for values
that cancel to ~0.  Where the kernel follows the reference's op order exactly (envs, GAE
horizon-major, Adam) the observed mismatch is reported as a count of non-identical floats.
"""

from __future__ import annotations

import math

import pytest
import torch

from oracle import ppo_oracle as O

from .conftest import Golden

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6
DEV = "cuda"


def close(a: torch.Tensor, b: torch.Tensor, rtol: float = RTOL, atol: float = ATOL) -> None:
    torch.testing.assert_close(a.cpu(), b.cpu(), rtol=rtol, atol=atol)


def _lib():
    from rl8_b200 import _lib as L

    return L, L.load()


# ---------------------------------------------------------------------------------------
# Environments
# ---------------------------------------------------------------------------------------

ENVS = {
    "discrete_dummy": ("DiscreteDummyEnv", 1, 1, True),
    "continuous_dummy": ("ContinuousDummyEnv", 1, 1, False),
    "cartpole": ("CartPole", 4, 5, True),
    "mountain_car": ("MountainCar", 2, 2, True),
    "pendulum": ("Pendulum", 2, 3, False),
}


def _random_state(name: str, n: int, gen: torch.Generator) -> torch.Tensor:
    if name.endswith("dummy"):
        return (torch.rand(n, 1, generator=gen) * 200 - 100)
    if name == "cartpole":
        return torch.randn(4, n, generator=gen) * torch.tensor([[1.0], [1.0], [2.0], [2.0]])
    if name == "mountain_car":
        p = torch.rand(1, n, generator=gen) * 1.8 - 1.2
        v = torch.rand(1, n, generator=gen) * 0.14 - 0.07
        return torch.vstack((p, v))
    th = torch.rand(1, n, generator=gen) * 20 - 10
    thd = torch.rand(1, n, generator=gen) * 16 - 8
    return torch.vstack((th, thd))


@pytest.mark.parametrize("name", list(ENVS))
@pytest.mark.parametrize("n", [4096, 1003])  # vectorised and scalar (ragged) paths
def test_env_step_matches_oracle(name: str, n: int) -> None:
    import rl8_b200.env as E

    cls_name, S, D, discrete = ENVS[name]
    gen = torch.Generator().manual_seed(sum(map(ord, name)) + n)
    env = getattr(E, cls_name)(n, 8, device=DEV)
    env.reset()
    state = _random_state(name, n, gen)
    oenv = O.OracleEnv(name, n)
    oenv.reset(state)
    obs0 = env.set_state(state.to(DEV))
    close(obs0, oenv.obs())
    for step in range(3):
        if discrete:
            a = torch.randint(0, 2 if "dummy" in name else 3, (n, 1), generator=gen)
        else:
            a = torch.randn(n, 1, generator=gen) * 2
        out = env.step(a.to(DEV))
        o_obs, o_r = oenv.step(a)
        # cos/sin of an angle that differs by 1 ulp (|theta| up to ~30) moves by ~4e-6
        close(out["obs"], o_obs.reshape(n, D), atol=8e-6)
        close(out["rewards"], o_r, atol=4e-6)
        close(env.state.reshape(-1), oenv.state.reshape(-1), atol=2e-6)
        assert out["obs"].shape == (n, D) and out["rewards"].shape == (n, 1)


def test_env_known_answers(kat: Golden) -> None:
    import rl8_b200.env as E

    for name, cls_name in (("cartpole", "CartPole"), ("pendulum", "Pendulum"), ("mountain_car", "MountainCar")):
        st = kat[f"{name}/state_in"]
        n = st.shape[1]
        env = getattr(E, cls_name)(n, 8, device=DEV)
        env.reset()
        env.set_state(st.to(DEV))
        out = env.step(kat[f"{name}/action"].to(DEV))
        close(env.state, kat[f"{name}/state_out"])
        close(out["rewards"], kat[f"{name}/reward"])
        if name != "mountain_car":
            close(out["obs"], kat[f"{name}/obs"])
    # exact-equality edge cases of mountain car (clip to the wall, goal reward)
    assert float(env.state[1, 1]) == 0.0 and float(out["rewards"][2, 0]) == 1.0


def test_env_reset_distributions() -> None:
    import rl8_b200.env as E

    torch.manual_seed(0)
    n = 1 << 16
    cp = E.CartPole(n, 8, device=DEV)
    obs = cp.reset()
    assert obs.shape == (n, 5) and abs(float(cp.state.std()) - 0.01) < 5e-4
    close(obs[:, 2], torch.cos(cp.state[2]))
    pd = E.Pendulum(n, 8, device=DEV)
    pd.reset()
    assert float(pd.state[0].min()) >= -math.pi and float(pd.state[0].max()) <= math.pi
    assert float(pd.state[1].abs().max()) <= 1.0
    dm = E.DiscreteDummyEnv(n, 8, device=DEV)
    o = dm.reset(config={"bounds": 3.0})
    assert o.shape == (n, 1) and float(o.abs().max()) <= 3.0
    mc = E.MountainCar(n, 8, device=DEV)
    mc.reset()
    assert abs(float(mc.state[0].mean()) + 0.5) < 1e-2


# ---------------------------------------------------------------------------------------
# Distributions
# ---------------------------------------------------------------------------------------


def test_categorical_matches_oracle_and_golden(kat: Golden) -> None:
    from rl8_b200.distributions import Categorical

    logits, q = kat["dist/logits"], kat["dist/q"]

    class Inj(Categorical):
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            return q.reshape(1, num, width).to(device)

    d = Inj({"logits": logits.to(DEV)}, None)
    a = d.sample()
    assert a.dtype == torch.int64 and a.shape == (512, 1)
    assert torch.equal(a.cpu(), kat["dist/cat_sample"]), "sampled discrete actions must be bit-exact"
    close(d.logp(a), kat["dist/cat_logp"])
    close(d.entropy(), kat["dist/cat_entropy"])
    assert torch.equal(d.deterministic_sample().cpu(), kat["dist/cat_mode"])
    # larger randomized check against the oracle, A = 2..5
    gen = torch.Generator().manual_seed(5)
    for A in (2, 3, 5):
        lg = torch.randn(20000, 1, A, generator=gen) * 3
        qq = torch.empty(20000, 1, A).exponential_(1, generator=gen)

        class Inj2(Categorical):
            @classmethod
            def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
                return qq.reshape(1, num, width).to(device)

        d2 = Inj2({"logits": lg.to(DEV)}, None)
        a2 = d2.sample()
        ref = O.categorical_sample(lg, qq)
        mismatch = int((a2.cpu() != ref).sum())
        assert mismatch == 0, f"{mismatch} of 20000 sampled actions differ (A={A})"
        close(d2.logp(a2), O.categorical_logp(lg, ref))
        close(d2.entropy(), O.categorical_entropy(lg))


def test_normal_and_squashed_match_golden(kat: Golden) -> None:
    from rl8_b200.distributions import Normal, SquashedNormal

    mean, log_std, z = kat["dist/mean"], kat["dist/log_std"], kat["dist/z"]

    class InjN(Normal):
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            return z.reshape(1, num).to(device)

    feats = {"mean": mean.to(DEV), "log_std": log_std.to(DEV)}
    d = InjN(feats, None)
    x = d.sample()
    close(x, kat["dist/normal_sample"])
    close(d.logp(kat["dist/normal_sample"].to(DEV)), kat["dist/normal_logp"])
    close(d.entropy(), kat["dist/normal_entropy"])
    close(d.deterministic_sample(), mean)
    s = SquashedNormal(feats, None)
    close(s.logp(kat["dist/squashed_x"].to(DEV)), kat["dist/squashed_logp"], rtol=2e-5, atol=2e-5)
    with pytest.raises(NotImplementedError):
        s.entropy()


# ---------------------------------------------------------------------------------------
# GAE
# ---------------------------------------------------------------------------------------


def test_gae_reference_known_answer_exact() -> None:
    """Upstream tests/test_nn/test_functional.py:14-49: exact equality, inplace aliasing."""
    from rl8_b200.nn import generalized_advantage_estimate

    N, T = 10, 5
    for inplace in (False, True):
        batch = {"rewards": torch.ones(N, T + 1, 1, device=DEV), "values": torch.ones(N, T + 1, 1, device=DEV)}
        und = torch.flip(torch.cumsum(batch["rewards"], dim=1), dims=(1,))
        out = generalized_advantage_estimate(
            batch, gae_lambda=1, gamma=1, inplace=inplace, normalize_advantages=False, return_returns=True
        )
        assert (out["advantages"] == und - 1).all()
        assert (out["returns"] == und).all()
        assert (out is batch) == inplace


@pytest.mark.parametrize("layout", ["env_major", "horizon_major"])
@pytest.mark.parametrize("shape", [(64, 5), (1000, 32), (1003, 33), (257, 70), (4096, 64)])
@pytest.mark.parametrize("normalize", [False, True])
def test_gae_matches_oracle(layout: str, shape: tuple[int, int], normalize: bool) -> None:
    from rl8_b200.nn import generalized_advantage_estimate

    N, T = shape
    gen = torch.Generator().manual_seed(N + T)
    r = torch.randn(N, T + 1, 1, generator=gen)
    v = torch.randn(N, T + 1, 1, generator=gen)
    o_r, o_adv, o_ret = O.gae(r, v, gamma=0.95, gae_lambda=0.9, reward_scale=1.7, normalize_advantages=normalize)
    if layout == "env_major":
        rd, vd = r.to(DEV), v.to(DEV)
    else:  # [T+1][N] storage presented as [N, T+1, 1]
        rd = r.squeeze(-1).T.contiguous().to(DEV).T.unsqueeze(-1)
        vd = v.squeeze(-1).T.contiguous().to(DEV).T.unsqueeze(-1)
    batch = {"rewards": rd, "values": vd}
    out = generalized_advantage_estimate(
        batch, gae_lambda=0.9, gamma=0.95, reward_scale=1.7, normalize_advantages=normalize
    )
    close(batch["rewards"], o_r)
    close(out["returns"], o_ret, atol=2e-6)
    close(out["advantages"], o_adv, atol=5e-6)
    if layout == "horizon_major" and not normalize:
        # same op order as the reference -> bit-identical
        assert torch.equal(out["advantages"].cpu(), o_adv)
        assert torch.equal(out["returns"].cpu(), o_ret)


@pytest.mark.parametrize("shape", [(64, 5), (1000, 32), (1003, 33), (4096, 64)])
def test_gae_scan_device_scale_forms(shape: tuple[int, int]) -> None:
    """rl8_gae_scan_dev (the form Algorithm.step launches: divisor read from device memory): with the scaled-reward
    write-back it equals rl8_gae_scan bit for bit; without it (write_scaled_rewards = 0) advantages / returns /
    moments are the same bits and `rewards` is left untouched (functional.py:106's side effect skipped)."""
    L, lib = _lib()
    N, T = shape
    gen = torch.Generator().manual_seed(N * 7 + T)
    r0 = torch.randn(T + 1, N, generator=gen).to(DEV)
    v = torch.randn(T + 1, N, generator=gen).to(DEV)
    scale = 1.7
    sdev = torch.tensor([scale, scale + 1e-8], dtype=torch.float32, device=DEV)  # what rl8_reward_scale leaves
    st = L.stream()
    res = []
    for form in ("host", "dev_write", "dev_keep"):
        r = r0.clone()
        adv, ret = torch.full_like(r, float("nan")), torch.full_like(r, float("nan"))
        mom = torch.zeros(3, dtype=torch.float64, device=DEV)
        if form == "host":
            rc = lib.rl8_gae_scan(L.ptr(r), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, 1, N, 0.95, 0.9, scale, L.ptr(mom), st)
        else:
            rc = lib.rl8_gae_scan_dev(L.ptr(r), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, 1, N, 0.95, 0.9, L.ptr(sdev),
                                      1 if form == "dev_write" else 0, L.ptr(mom), st)
        assert rc == 0
        res.append((r, adv[:T], ret[:T], mom))
    o_r, o_adv, o_ret = O.gae(r0.T.unsqueeze(-1).cpu(), v.T.unsqueeze(-1).cpu(), gamma=0.95, gae_lambda=0.9,
                              reward_scale=scale, normalize_advantages=False)
    assert torch.equal(res[0][1].T.cpu(), o_adv.squeeze(-1)[:, :T])
    for k in (1, 2):
        assert torch.equal(res[k][1], res[0][1]) and torch.equal(res[k][2], res[0][2])
        assert torch.equal(res[k][3], res[0][3])
    assert torch.equal(res[1][0], res[0][0]) and torch.equal(res[0][0].T.cpu(), o_r.squeeze(-1))
    assert torch.equal(res[2][0], r0)
    # the env-major layout has no keep-rewards variant
    rc = lib.rl8_gae_scan_dev(L.ptr(r0), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, T + 1, 1, 0.95, 0.9, L.ptr(sdev), 0,
                              L.ptr(mom), st)
    assert rc != 0


def test_gae_golden(kat: Golden) -> None:
    from rl8_b200.nn import generalized_advantage_estimate

    for tag, norm in (("raw", False), ("norm", True)):
        batch = {"rewards": kat["gae/rewards"].to(DEV), "values": kat["gae/values"].to(DEV)}
        out = generalized_advantage_estimate(
            batch, gae_lambda=0.95, gamma=0.95, reward_scale=2.0, normalize_advantages=norm
        )
        close(out["advantages"], kat[f"gae/{tag}/advantages"])
        close(out["returns"], kat[f"gae/{tag}/returns"])


# ---------------------------------------------------------------------------------------
# PPO losses (values + hand-derived gradients vs autograd of the oracle)
# ---------------------------------------------------------------------------------------


def test_ppo_losses_known_answer(kat: Golden) -> None:
    from rl8_b200.distributions import Categorical
    from rl8_b200.nn import ppo_losses

    for tag, dual in (("nodual", None), ("dual", 5.0)):
        buf = {
            "actions": kat["ppo/actions"].to(DEV),
            "logp": kat["ppo/logp_old"].to(DEV),
            "advantages": kat["ppo/advantages"].to(DEV),
            "returns": kat["ppo/returns"].to(DEV),
        }
        d = Categorical({"logits": kat["ppo/logits"].to(DEV)}, None)
        losses = ppo_losses(
            buf, {"values": kat["ppo/values"].to(DEV)}, d, clip_param=0.2, dual_clip_param=dual,
            entropy_coeff=0.01, vf_clip_param=5.0, vf_coeff=1.0,
        )
        for k in ("entropy", "policy", "vf", "total"):
            close(losses[k], kat[f"ppo/{tag}/{k}"].reshape(()))


@pytest.mark.parametrize("kind", ["categorical", "normal", "squashed_normal"])
@pytest.mark.parametrize("dual", [None, 3.0])
def test_ppo_losses_and_gradients_match_autograd(kind: str, dual: None | float) -> None:
    from rl8_b200 import distributions as Dm
    from rl8_b200.nn import ppo_losses

    B = 4096
    gen = torch.Generator().manual_seed(11)
    adv = torch.randn(B, 1, generator=gen)
    ret = torch.randn(B, 1, generator=gen) * 3
    values = torch.randn(B, 1, generator=gen) * 3
    ent_coeff = 0.0 if kind == "squashed_normal" else 0.02
    if kind == "categorical":
        logits = (torch.randn(B, 1, 3, generator=gen) * 1.5).requires_grad_(True)
        feats = {"logits": logits}
        actions = torch.randint(0, 3, (B, 1), generator=gen)
        bound = O.Dist(kind).bind(feats)
        logp_old = (bound.logp(actions) + torch.randn(B, 1, generator=gen) * 0.3).detach()
        leaves = [logits]
    else:
        mean = torch.randn(B, 1, generator=gen).requires_grad_(True)
        raw = torch.randn(B, 1, generator=gen).requires_grad_(True)
        feats = {"mean": mean, "log_std": torch.tanh(raw)}
        bound = O.Dist(kind).bind(feats)
        actions = bound.sample(torch.randn(B, 1, generator=gen)).detach()
        logp_old = (bound.logp(actions) + torch.randn(B, 1, generator=gen) * 0.3).detach()
        leaves = [mean, raw]
    v = values.clone().requires_grad_(True)
    ent = bound.entropy() if ent_coeff else None
    ref = O.ppo_losses(
        bound.logp(actions), v, ent, logp_old, adv, ret, clip_param=0.2, dual_clip_param=dual,
        entropy_coeff=ent_coeff, vf_clip_param=2.0, vf_coeff=0.7,
    )
    ref["total"].backward()
    cls = {"categorical": Dm.Categorical, "normal": Dm.Normal, "squashed_normal": Dm.SquashedNormal}[kind]
    d = cls({k: t.detach().to(DEV) for k, t in feats.items()}, None)
    buf = {"actions": actions.to(DEV), "logp": logp_old.to(DEV), "advantages": adv.to(DEV), "returns": ret.to(DEV)}
    got = ppo_losses(
        buf, {"values": values.to(DEV)}, d, clip_param=0.2, dual_clip_param=dual,
        entropy_coeff=ent_coeff, vf_clip_param=2.0, vf_coeff=0.7, return_grads=True,
    )
    for k in ("entropy", "policy", "vf", "total"):
        close(got[k], ref[k].detach().reshape(()), atol=1e-6)
    close(got["kl_div"], O.approx_kl(bound.logp(actions), logp_old).detach())
    ref_d = torch.cat([t.grad.reshape(B, -1) for t in leaves], dim=1)
    close(got["d_features"], ref_d, rtol=1e-4, atol=1e-9)
    close(got["d_values"], v.grad, rtol=1e-5, atol=1e-10)


# ---------------------------------------------------------------------------------------
# MLP forward, optimizer
# ---------------------------------------------------------------------------------------


def _policy(env_name: str, dist: None | str = None, n: int = 8):
    import rl8_b200.env as E
    from rl8_b200 import distributions as Dm
    from rl8_b200.policies import Policy

    env = getattr(E, ENVS[env_name][0])(n, 8, device=DEV)
    dcls = {None: None, "normal": Dm.Normal, "squashed_normal": Dm.SquashedNormal, "categorical": Dm.Categorical}[dist]
    return Policy(env.observation_spec, env.action_spec, distribution_cls=dcls, device=DEV)


@pytest.mark.parametrize("env_name", ["cartpole", "pendulum", "discrete_dummy"])
@pytest.mark.parametrize("rows", [1, 77, 4096])
def test_mlp_forward_fp32_matches_oracle(env_name: str, rows: int) -> None:
    torch.manual_seed(3)
    pol = _policy(env_name)
    params = {k: v.detach().cpu().clone() for k, v in pol.model.state_dict().items()}
    D = ENVS[env_name][2]
    obs = torch.randn(rows, D) * 2
    feats, value = O.model_forward(params, obs)
    out = pol.sample({"obs": obs.to(DEV).unsqueeze(1)}, kind="last", return_actions=False, return_values=True)
    for k, t in feats.items():
        close(out["features"][k], t, rtol=2e-5, atol=2e-6)
    close(out["values"], value, rtol=2e-5, atol=2e-6)
    # strided (horizon-major SoA) observations give the same numbers
    obs_soa = obs.T.contiguous().to(DEV)
    close(pol.forward_net(1, obs_soa.T), value, rtol=2e-5, atol=2e-6)


def test_reference_init_matches_oracle_init() -> None:
    """Same construction order / RNG consumption as the reference's default models."""
    for env_name, kind, A in (("cartpole", "discrete", 3), ("pendulum", "continuous", 1)):
        torch.manual_seed(123)
        pol = _policy(env_name)
        torch.manual_seed(123)
        ref = O.init_params(ENVS[env_name][2], kind, A)
        sd = pol.model.state_dict()
        assert set(sd) == set(ref)
        for k in ref:
            assert torch.equal(sd[k].cpu(), ref[k]), k


def test_clip_adam_matches_torch() -> None:
    L, lib = _lib()
    n = 135_684
    gen = torch.Generator().manual_seed(2)
    p0 = torch.randn(n, generator=gen)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    norm = torch.zeros(1, device=DEV)
    for step in range(1, 6):
        scale = 10.0 if step % 2 else 0.01  # clipped and un-clipped steps
        g = torch.randn(n, generator=gen) * scale
        p_ref.grad = g.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([p_ref], 5.0)
        opt.step()
        gd = g.to(DEV)
        rc = lib.rl8_clip_adam(L.ptr(p), L.ptr(gd), L.ptr(m), L.ptr(v), n, 5.0, 1e-3, 0.9, 0.999, 1e-8,
                               step, L.ptr(norm), L.stream())
        assert rc == 0
        close(norm, ref_norm.reshape(1), rtol=1e-6, atol=0)
        close(p, p_ref.detach(), rtol=1e-6, atol=1e-7)
    st = opt.state[p_ref]
    # moments mix +-O(1) terms: compare against their own scale, not element-wise ulp
    close(m, st["exp_avg"], rtol=1e-5, atol=1e-7)
    close(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-9)


def test_clip_adam_device_scalars_match_torch() -> None:
    """rl8_clip_adam_dev (learning rate and step count in device memory: the form a CUDA graph replays) against
    torch.optim.Adam + clip_grad_norm_, with the learning rate changed between steps."""
    L, lib = _lib()
    n = 135_684
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=gen)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p = p0.clone().to(DEV)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    scratch = torch.zeros(16, device=DEV)
    lr_dev = torch.full((1,), 1e-3, dtype=torch.float64, device=DEV)
    step_dev = torch.zeros(1, dtype=torch.int64, device=DEV)
    for step in range(1, 8):
        if step == 4:
            opt.param_groups[0]["lr"] = 2.5e-4
            lr_dev.fill_(2.5e-4)
        scale = 10.0 if step % 2 else 0.01
        g = torch.randn(n, generator=gen) * scale
        p_ref.grad = g.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([p_ref], 5.0)
        opt.step()
        gd = g.to(DEV)
        rc = lib.rl8_clip_adam_dev(L.ptr(p), L.ptr(gd), L.ptr(m), L.ptr(v), n, 5.0, L.ptr(lr_dev), 0.9, 0.999, 1e-8,
                                   L.ptr(step_dev), L.ptr(scratch), L.stream())
        assert rc == 0
        assert int(step_dev.item()) == step
        close(scratch[:1], ref_norm.reshape(1), rtol=1e-6, atol=0)
        close(p, p_ref.detach(), rtol=1e-6, atol=1e-7)
    st = opt.state[p_ref]
    close(m, st["exp_avg"], rtol=1e-5, atol=1e-7)
    close(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-9)
