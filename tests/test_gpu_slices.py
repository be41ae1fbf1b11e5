"""GPU: BASELINE.json's configs at FULL size, a 2 048-env slice replayed on the CPU oracle.

Environments never interact inside ``collect()``, so envs ``[e0, e0 + 2048)`` of a 65 536- / 262 144-env rollout
must equal the oracle's rollout of those 2 048 envs from the same injected initial state and sampling noise;
the stages with global statistics are compared teacher-forced:

* collect: every buffer field of the slice (discrete actions bit-exact) -- fp32 mode (``enable_amp=False``: split
  tcgen05 kernels) at the north-star tolerance;
* GAE: scaled rewards / returns of the slice from the device's global ``reward_scale``; normalised advantages from
  the device's own global moments;
* update: ``rl8_ppo_minibatch`` over exactly the slice's rows of the full-size buffer (row offset e0 * T) against
  the oracle's losses and autograd gradients on the slice.

``enable_amp=True`` (plain bf16 operands) is compared at bf16 tolerances, teacher-forced on the device's own
trajectories (one flipped action would otherwise change an env's whole future).  The recurrent config runs the
fp32 CUDA-core path against the recurrent oracle, and the bf16 LSTM cell kernel (``tc_lstm_cell_kernel``,
rows >= 512) against a bf16-operand emulation.
"""

from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

from oracle import ppo_oracle as O
from oracle import recurrent_oracle as R

pytestmark = pytest.mark.gpu
DEV = "cuda"
SLICE = 2048


def _build(env_name: str, dist_name: str, N: int, T: int, amp: bool, e0: int, seed: int, recurrent: bool = False,
           **kw):
    """Full-size algorithm whose env state and noise are injected; returns (algo, state0 slice, noise slice)."""
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig, RecurrentAlgorithmConfig
    from rl8_b200 import distributions as Dm

    env_cls = getattr(E, env_name)
    base = {"categorical": Dm.Categorical, "squashed_normal": Dm.SquashedNormal, "normal": Dm.Normal}[dist_name]
    g = torch.Generator(device=DEV).manual_seed(seed)
    discrete = dist_name == "categorical"
    width = 3 if discrete else 1
    noise = torch.empty(T, N, width, device=DEV)
    noise = noise.exponential_(1, generator=g) if discrete else noise.normal_(generator=g)
    S = {"CartPole": 4, "Pendulum": 2}[env_name]
    state0 = torch.randn(S, N, device=DEV, generator=g) * (0.05 if env_name == "CartPole" else 1.0)

    class InjDist(base):  # type: ignore[misc, valid-type]
        @classmethod
        def draw_noise(cls, steps, num, w, device):  # noqa: ANN001, ANN206
            if steps != T:
                return super().draw_noise(steps, num, w, device)
            return noise if discrete else noise.reshape(T, N)

    class InjEnv(env_cls):  # type: ignore[misc, valid-type]
        def reset(self, *, config=None):  # noqa: ANN001, ANN202
            super().reset(config=config)
            return self.set_state(state0.clone())

    InjEnv.__name__ = env_cls.__name__
    torch.manual_seed(seed)
    cfg = (RecurrentAlgorithmConfig if recurrent else AlgorithmConfig)(
        num_envs=N, horizon=T, enable_amp=amp, distribution_cls=InjDist, **kw)
    algo = cfg.build(InjEnv)
    sl = slice(e0, e0 + SLICE)
    nz = noise[:, sl].cpu()
    nz = nz.reshape(T, SLICE, 1, width) if discrete else nz.reshape(T, SLICE, 1)
    return algo, state0[:, sl].cpu(), nz


def _slice_buffers(algo, e0: int) -> dict[str, torch.Tensor]:
    return {k: algo.buffer[k][e0:e0 + SLICE].cpu() for k in
            ("obs", "rewards", "actions", "logp", "values", "advantages", "returns", "reversed_discounted_returns")}


CASES = [
    ("CartPole", "cartpole", "categorical", 65536, 32, 40_000),           # configs[1]
    ("Pendulum", "pendulum", "squashed_normal", 262144, 64, 200_000),     # configs[2]
]


@pytest.mark.parametrize("env_name,oname,dist_name,N,T,e0", CASES)
def test_full_size_slice_matches_oracle_fp32(env_name: str, oname: str, dist_name: str, N: int, T: int, e0: int) -> None:
    from rl8_b200 import _lib as L

    algo, state0, noise = _build(env_name, dist_name, N, T, False, e0, seed=21)
    assert algo.policy.precision == L.PREC_FP32_TC
    params = {k: v.detach().cpu().clone() for k, v in algo.policy.model.state_dict().items()}
    algo.collect()
    got = _slice_buffers(algo, e0)
    # ---- collect: the oracle rolls the slice's envs out by itself
    o_env = O.OracleEnv(oname, SLICE)
    kind = "discrete" if dist_name == "categorical" else "continuous"
    o_buf = O.new_buffer(SLICE, T, o_env.obs_dim, kind)
    dist = O.Dist(dist_name)
    O.collect(params, o_env, o_buf, dist, noise, reset_state=state0)
    if kind == "discrete":
        assert torch.equal(got["actions"], o_buf["actions"]), "sampled discrete actions must be bit-exact"
        for k in ("obs", "rewards", "logp", "values", "reversed_discounted_returns"):
            torch.testing.assert_close(got[k], o_buf[k], rtol=2e-5, atol=5e-6, msg=lambda m, k=k: f"{k}: {m}")
    else:
        # Continuous actions feed back into the pendulum's dynamics: a last-bit difference grows by orders of
        # magnitude over 64 steps (both runs are valid fp32 rollouts).  Free-running agreement is asserted for the
        # first 8 steps; every later step is checked teacher-forced on the stored trajectory.
        for k in ("obs", "rewards", "actions", "logp", "values"):
            torch.testing.assert_close(got[k][:, :8], o_buf[k][:, :8], rtol=2e-5, atol=5e-6, msg=lambda m, k=k: f"{k}: {m}")
        feats, values = O.model_forward(params, got["obs"].reshape(SLICE * (T + 1), -1))
        torch.testing.assert_close(got["values"].reshape(-1, 1), values, rtol=2e-5, atol=5e-6)
        d = dist.bind({k: v.reshape(SLICE, T + 1, 1)[:, :T].reshape(SLICE * T, 1) for k, v in feats.items()})
        z = noise.permute(1, 0, 2).reshape(SLICE * T, 1)
        torch.testing.assert_close(got["actions"][:, :T].reshape(-1, 1), d.sample(z), rtol=2e-5, atol=5e-6)
        torch.testing.assert_close(got["logp"][:, :T].reshape(-1, 1), d.logp(got["actions"][:, :T].reshape(-1, 1)),
                                   rtol=1e-4, atol=5e-5)  # log(1 - x^2 + eps) of the squashed action near +-1
        ob = got["obs"][:, :T].reshape(SLICE * T, 3)
        state = torch.stack((torch.atan2(ob[:, 1], ob[:, 0]), ob[:, 2]))
        _, ob1, r = O.pendulum_step(state, got["actions"][:, :T].reshape(-1, 1))
        torch.testing.assert_close(got["obs"][:, 1:].reshape(SLICE * T, 3), ob1, rtol=2e-5, atol=1e-5)
        torch.testing.assert_close(got["rewards"][:, :T].reshape(-1, 1), r.reshape(-1, 1), rtol=2e-5, atol=2e-5)
    # ---- GAE with the device's global reward scale / moments, teacher-forced on the slice's stored rewards / values
    o_buf = {k: v.clone() for k, v in got.items()}
    scale = algo.state.reward_scale
    hp, lib, buf = algo.hparams, algo._lib, algo.buffer
    algo._moments.zero_()
    lib.rl8_gae_scan(L.ptr(buf.hm["rewards"]), L.ptr(buf.hm["values"]), L.ptr(buf.hm["advantages"]),
                     L.ptr(buf.hm["returns"]), N, T, 1, N, hp.gamma, hp.gae_lambda, scale, L.ptr(algo._moments),
                     L.stream())
    raw_adv = buf["advantages"][e0:e0 + SLICE].cpu().clone()
    lib.rl8_gae_normalize(L.ptr(buf.hm["advantages"]), N, T, 1, N, L.ptr(algo._moments), L.stream())
    o_r, o_adv, o_ret = O.gae(o_buf["rewards"], o_buf["values"], reward_scale=scale, normalize_advantages=False)
    torch.testing.assert_close(buf["rewards"][e0:e0 + SLICE].cpu(), o_r, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(raw_adv, o_adv, rtol=1e-5, atol=5e-6)
    torch.testing.assert_close(buf["returns"][e0:e0 + SLICE].cpu(), o_ret, rtol=1e-5, atol=5e-6)
    s, s2, cnt = algo._moments.tolist()
    mean = s / cnt
    std = ((s2 - s * mean) / (cnt - 1)) ** 0.5
    all_adv = buf.hm["advantages"][:T]
    assert abs(float(all_adv.double().mean())) < 1e-5 and abs(float(all_adv.double().std()) - 1.0) < 1e-4
    want = (o_adv[:, :-1].double() - mean) / (std + 1e-8)
    torch.testing.assert_close(buf["advantages"][e0:e0 + SLICE, :-1].cpu().double(), want, rtol=1e-5, atol=5e-6)
    # ---- update: one minibatch = exactly the slice's rows of the full-size buffer
    model = algo.policy.model
    algo._grads.zero_()
    m, g = model.struct_for(model.flat_params), model.struct_for(algo._grads)
    M = SLICE * T
    ws = algo._workspace("ppo", int(lib.rl8_ppo_workspace(m, M, algo.policy.precision)))
    ppo = L.PpoHparams(hp.clip_param, 0.0, 0.0, hp.vf_clip_param, hp.vf_coeff, 1.0)
    sums = torch.zeros(5, dtype=torch.float64, device=DEV)
    rc = lib.rl8_ppo_minibatch(m, g, algo._batch_struct(), None, e0 * T, M, float(M), ppo, L.ptr(sums),
                               algo.policy.precision, L.ptr(ws), ws.numel(), L.stream())
    assert rc == 0, rc
    gsl = _slice_buffers(algo, e0)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    mb = {k: gsl[k][:, :-1].reshape(M, -1) for k in ("obs", "actions", "logp", "advantages", "returns")}
    losses, kl = O.minibatch_losses(p, dist, mb, clip_param=hp.clip_param, dual_clip_param=None, entropy_coeff=0.0,
                                    vf_clip_param=hp.vf_clip_param, vf_coeff=hp.vf_coeff)
    losses["total"].backward()
    s_ent, s_pol, s_vf, s_kl, n = sums.tolist()
    assert n == M
    assert s_pol / n == pytest.approx(float(losses["policy"]), rel=2e-5, abs=2e-6)
    assert s_vf / n == pytest.approx(float(losses["vf"]), rel=2e-5, abs=2e-6)
    assert s_kl / n == pytest.approx(float(kl), rel=2e-5, abs=2e-6)
    named = model.named_flat_views(algo._grads)
    gnorm = float(torch.cat([p[k].grad.flatten() for k in p]).double().norm())
    for k in p:
        ref = p[k].grad.double()
        err = float((named[k].cpu().double() - ref).norm())
        assert err <= max(2e-5 * float(ref.norm()), 2e-6 * gnorm), (k, err, float(ref.norm()), gnorm)


@pytest.mark.parametrize("env_name,oname,dist_name,N,T,e0", CASES)
def test_full_size_slice_bf16_teacher_forced(env_name: str, oname: str, dist_name: str, N: int, T: int, e0: int) -> None:
    """enable_amp=True: the stored log-probabilities / values of the slice are the oracle's functions of the stored
    observations and actions within bf16 operand rounding; the env transitions are exact given the stored actions."""
    algo, state0, noise = _build(env_name, dist_name, N, T, True, e0, seed=22)
    params = {k: v.detach().cpu().clone() for k, v in algo.policy.model.state_dict().items()}
    algo.collect()
    got = _slice_buffers(algo, e0)
    obs = got["obs"].reshape(SLICE * (T + 1), -1)
    feats, values = O.model_forward(params, obs)
    vscale = max(1.0, float(values.abs().max()))
    assert float((got["values"].reshape(-1, 1) - values).abs().max()) < 1e-2 * vscale
    d = O.Dist(dist_name).bind({k: v.reshape(SLICE, T + 1, *v.shape[1:])[:, :T].reshape(SLICE * T, *v.shape[1:])
                                for k, v in feats.items()})
    logp = d.logp(got["actions"][:, :T].reshape(SLICE * T, 1))
    assert float((got["logp"][:, :T].reshape(-1, 1) - logp).abs().max()) < 2e-2
    # env transitions replayed from the STORED actions are exact (the env kernels are fp32 in every mode)
    o_env = O.OracleEnv(oname, SLICE)
    ob = o_env.reset(state0)
    torch.testing.assert_close(got["obs"][:, 0], ob, rtol=2e-5, atol=5e-6)
    for t in range(4):
        ob, r = o_env.step(got["actions"][:, t])
        torch.testing.assert_close(got["obs"][:, t + 1], ob, rtol=2e-5, atol=5e-6)
        torch.testing.assert_close(got["rewards"][:, t], r, rtol=2e-5, atol=5e-6)


def test_full_size_slice_recurrent_fp32_matches_oracle() -> None:
    """configs[3] (CartPole LSTM, N = 65 536, T = 32), fp32 path: a 2 048-env slice against the recurrent oracle."""
    N, T, e0 = 65536, 32, 30_000
    algo, state0, noise = _build("CartPole", "categorical", N, T, False, e0, seed=23, recurrent=True)
    params = {k: v.detach().cpu().clone() for k, v in algo.policy.model.state_dict().items()}
    algo.collect()
    o_env = O.OracleEnv("cartpole", SLICE)
    o_buf = R.new_recurrent_buffer(SLICE, T, 5, "discrete")
    R.collect_recurrent(params, o_env, o_buf, O.Dist("categorical"), noise, seqs=0, reset_state=state0)
    assert torch.equal(algo.buffer["actions"][e0:e0 + SLICE].cpu(), o_buf["actions"])
    for k in ("obs", "rewards", "logp", "values"):
        torch.testing.assert_close(algo.buffer[k][e0:e0 + SLICE].cpu(), o_buf[k], rtol=2e-5, atol=5e-6,
                                   msg=lambda m, k=k: f"{k}: {m}")
    for k in ("hidden_states", "cell_states"):
        torch.testing.assert_close(algo.buffer["states"][k][e0:e0 + SLICE].cpu().reshape(o_buf[k].shape), o_buf[k],
                                   rtol=2e-5, atol=5e-6, msg=lambda m, k=k: f"{k}: {m}")


@pytest.mark.parametrize("rows", [512, 1000, 2048])
def test_lstm_cell_tensor_core_kernel_against_bf16_emulation(rows: int) -> None:
    """``tc_lstm_cell_kernel`` (enable_amp=True, rows >= 512: the gate GEMM with the cell in its epilogue): h', c'
    against the oracle's LSTM cell with h and W_hh rounded to bf16 where the kernel rounds them, and against the
    fp32 CUDA-core step at bf16 tolerance."""
    import rl8_b200.env as E
    from rl8_b200 import RecurrentAlgorithmConfig
    from rl8_b200 import _lib as L

    torch.manual_seed(31)
    algo = RecurrentAlgorithmConfig(num_envs=64, horizon=32, enable_amp=True).build(E.CartPole)
    pol = algo.policy
    params = {k: v.detach().cpu().clone() for k, v in pol.model.state_dict().items()}
    gen = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, 5, generator=gen)
    h = torch.tanh(torch.randn(rows, 256, generator=gen))
    c = torch.randn(rows, 256, generator=gen)
    assert pol.precision == L.PREC_BF16
    head16, val16, h16, c16 = pol.step_net(x.to(DEV), h.to(DEV), c.to(DEV))
    pol.precision = L.PREC_FP32
    head32, val32, h32, c32 = pol.step_net(x.to(DEV), h.to(DEV), c.to(DEV))
    # bf16 emulation: operands of the hidden-to-hidden contraction rounded, everything else fp32
    pe = dict(params)
    pe["lstm.weight_hh_l0"] = params["lstm.weight_hh_l0"].bfloat16().float()
    h_ref, c_ref = R.lstm_cell(pe, x, h.bfloat16().float(), c)
    # SFU sigmoid / tanh in the bf16 kernel (2 ulp exp, fast reciprocal): 2e-5 absolute on values in [-1, 1]
    torch.testing.assert_close(h16.cpu(), h_ref, rtol=1e-4, atol=5e-5)
    torch.testing.assert_close(c16.cpu(), c_ref, rtol=1e-4, atol=5e-5)
    h_fp, c_fp = R.lstm_cell(params, x, h, c)
    torch.testing.assert_close(h32.cpu(), h_fp, rtol=2e-5, atol=5e-6)
    torch.testing.assert_close(c32.cpu(), c_fp, rtol=2e-5, atol=5e-6)
    assert float((h16 - h32).abs().max()) < 2e-2 and float((c16 - c32).abs().max()) < 3e-2
    # heads of the bf16 step are fp32 functions of ITS h'
    want = F.linear(h16.cpu(), params["feature_head.weight"], params["feature_head.bias"])
    torch.testing.assert_close(head16.cpu(), want, rtol=2e-5, atol=5e-6)
