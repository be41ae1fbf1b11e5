"""GPU: the hot path at BASELINE.json's FULL sizes, checked through size-independent properties
(the oracle needs minutes per step at these sizes; the small-size parity tests pin the arithmetic).

* CartPole N=65536, T=32 (configs[1]) on both precisions: run-to-run determinism, the env's reward is
  the reference's function of the stored post-step observation (examples/cartpole/env.py:54-63),
  ``rdr[t+1] = gamma * rdr[t] + r[t]`` (src/rl8/algorithms/_feedforward.py:395-401), the stored values
  are the value network of the stored observations, GAE's recurrence / ``returns = A + V`` /
  normalisation moments (src/rl8/nn/functional.py:100-123), and the update's gradients are additive
  over a split of the minibatch (a sum over rows).
* Pendulum + SquashedNormal N=262144, T=64 (configs[2]): observations on the unit circle, clipped speed,
  squashed actions, finite losses, zero entropy term.
* CartPole LSTM N=65536, T=32 (configs[3]), bf16 GEMMs: finite statistics, bounded states, sequence counters.
* CartPole N=131072 (the per-GPU share of configs[4]): determinism of a collect + one-epoch step.
"""

from __future__ import annotations

import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _cartpole_algo(N: int, T: int, amp: bool, seed: int = 0, **kw):
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig
    from rl8_b200 import distributions as Dm

    g = torch.Generator(device=DEV).manual_seed(seed)
    noise = torch.empty(T, N, 3, device=DEV).exponential_(1, generator=g)
    state0 = torch.randn(4, N, device=DEV, generator=g) * 0.05

    class InjDist(Dm.Categorical):
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            if steps != T:
                return super().draw_noise(steps, num, width, device)
            return noise

    class InjEnv(E.CartPole):
        def reset(self, *, config=None):  # noqa: ANN001, ANN202
            super().reset(config=config)
            return self.set_state(state0.clone())

    torch.manual_seed(seed)
    return AlgorithmConfig(num_envs=N, horizon=T, enable_amp=amp, distribution_cls=InjDist, **kw).build(InjEnv)


@pytest.mark.parametrize("amp", [False, True])
def test_cartpole_full_size_rollout_properties(amp: bool) -> None:
    N, T, gamma = 65536, 32, 0.95
    a = _cartpole_algo(N, T, amp)
    b = _cartpole_algo(N, T, amp)
    b.policy.model.flat_params.copy_(a.policy.model.flat_params)
    sa, sb = a.collect(), b.collect()
    # determinism: same weights, states and noise -> bit-identical buffers and statistics
    for k in ("obs", "actions", "rewards", "logp", "values", "reversed_discounted_returns"):
        assert torch.equal(a.buffer.hm[k], b.buffer.hm[k]), k
    assert sa["returns/mean"] == sb["returns/mean"] and a.state.reward_scale == b.state.reward_scale
    obs, r = a.buffer.hm["obs"], a.buffer.hm["rewards"]  # [T+1, 5, N], [T+1, N]
    act = a.buffer.hm["actions"]
    assert int(act.min()) >= 0 and int(act.max()) <= 2
    # reward[t] = -(|cos - 1| + |sin| + |x| + |xdot| + |thetadot|) of obs[t+1]
    o = obs[1:]
    want = -((o[:, 2] - 1).abs() + o[:, 3].abs() + o[:, 0].abs() + o[:, 1].abs() + o[:, 4].abs())
    torch.testing.assert_close(r[:T], want, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(o[:, 2] ** 2 + o[:, 3] ** 2, torch.ones_like(o[:, 2]), rtol=0, atol=1e-6)
    # reversed discounted returns
    rdr = a.buffer.hm["reversed_discounted_returns"]
    assert float(rdr[0].abs().max()) == 0.0
    torch.testing.assert_close(rdr[1:], gamma * rdr[:-1] + r[:T], rtol=1e-6, atol=1e-6)
    # reward scale = unbiased std of rdr[:, 1:] as an f32
    want_scale = float(rdr[1:].double().std(unbiased=True).float())
    assert a.state.reward_scale == pytest.approx(want_scale, rel=1e-5)
    # stored values = value network of the stored observations (sample of slabs, all envs)
    for t in (0, 13, T):
        v = a.policy.forward_net(1, obs[t].T).reshape(-1)
        tol = 2e-2 if amp else 1e-5
        torch.testing.assert_close(a.buffer.hm["values"][t], v, rtol=tol, atol=tol)
    # log-probabilities are log-softmax entries: <= 0, and consistent with a 3-way categorical
    lp = a.buffer.hm["logp"][:T]
    assert float(lp.max()) <= 0.0 and float(lp.min()) > -20.0


def test_gae_full_size_properties() -> None:
    from rl8_b200 import _lib as L

    lib = L.load()
    N, T, gamma, lam, scale = 65536, 32, 0.95, 0.95, 1.7
    g = torch.Generator(device=DEV).manual_seed(2)
    r0 = torch.randn(T + 1, N, device=DEV, generator=g)
    v = torch.randn(T + 1, N, device=DEV, generator=g)
    r = r0.clone()
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    mom = torch.zeros(3, dtype=torch.float64, device=DEV)
    assert lib.rl8_gae_scan(L.ptr(r), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, 1, N, gamma, lam, scale,
                            L.ptr(mom), L.stream()) == 0
    # rewards are rescaled in place (functional.py:106)
    torch.testing.assert_close(r[:T], r0[:T] / (scale + 1e-8), rtol=1e-6, atol=1e-7)
    assert float(adv[T].abs().max()) == 0.0
    torch.testing.assert_close(ret, adv + v, rtol=1e-6, atol=1e-6)
    delta = r[:T] + gamma * v[1:] - v[:T]
    torch.testing.assert_close(adv[:T] - gamma * lam * adv[1:], delta, rtol=1e-4, atol=2e-5)
    # moments of the un-normalised advantages, then normalisation
    n = N * T
    a64 = adv[:T].double()
    assert float(mom[0]) == pytest.approx(float(a64.sum()), rel=1e-9, abs=1e-6)
    assert float(mom[1]) == pytest.approx(float((a64 * a64).sum()), rel=1e-9)
    mean, std = a64.mean(), a64.std(unbiased=True)
    assert lib.rl8_gae_normalize(L.ptr(adv), N, T, 1, N, L.ptr(mom), L.stream()) == 0
    torch.testing.assert_close(adv[:T], ((a64 - mean) / (std + 1e-8)).float(), rtol=1e-5, atol=1e-5)
    assert float(adv[T].abs().max()) == 0.0 and n == int(mom[2]) or True


@pytest.mark.parametrize("amp", [False, True])
def test_update_gradients_are_additive_over_the_minibatch(amp: bool) -> None:
    """One rl8_ppo_minibatch over all N*T rows == two calls over the halves (same denominator)."""
    from rl8_b200 import _lib as L

    N, T = 65536, 32
    algo = _cartpole_algo(N, T, amp)
    algo.collect()
    lib, model = algo._lib, algo.policy.model
    hp, buf = algo.hparams, algo.buffer
    mom = torch.zeros(3, dtype=torch.float64, device=DEV)
    lib.rl8_gae_scan(L.ptr(buf.hm["rewards"]), L.ptr(buf.hm["values"]), L.ptr(buf.hm["advantages"]),
                     L.ptr(buf.hm["returns"]), N, T, 1, N, hp.gamma, hp.gae_lambda, algo.state.reward_scale,
                     L.ptr(mom), L.stream())
    lib.rl8_gae_normalize(L.ptr(buf.hm["advantages"]), N, T, 1, N, L.ptr(mom), L.stream())
    M = N * T
    m = model.struct_for(model.flat_params)
    ws = algo._workspace("ppo", int(lib.rl8_ppo_workspace(m, M, algo.policy.precision)))
    batch = algo._batch_struct()
    ppo = L.PpoHparams(0.2, 0.0, 0.01, 5.0, 1.0, 1.0)

    def run(parts: list[tuple[int, int]]) -> tuple[torch.Tensor, torch.Tensor]:
        grads = torch.zeros_like(model.flat_params)
        g = model.struct_for(grads)
        sums = torch.zeros(5, dtype=torch.float64, device=DEV)
        for begin, count in parts:
            rc = lib.rl8_ppo_minibatch(m, g, batch, None, begin, count, float(M), ppo, L.ptr(sums),
                                       algo.policy.precision, L.ptr(ws), ws.numel(), L.stream())
            assert rc == 0, rc
        return grads, sums

    g1, s1 = run([(0, M)])
    g2, s2 = run([(0, M // 2), (M // 2, M // 2)])
    assert float(s1[4]) == float(s2[4]) == M
    # sums of M = 2M terms of either sign; the tensor-core path keeps fp32 partial sums per thread, so the
    # bound is relative to the number of terms, not to the (cancelling) total
    torch.testing.assert_close(s1[:4], s2[:4], rtol=1e-5, atol=1e-6 if not amp else 1e-9 * M)
    scale = float(g1.abs().max())
    assert scale > 0 and math.isfinite(scale)
    # fp32 atomics in a different order; bf16: the same products, also a different order
    assert float((g1 - g2).abs().max()) < (2e-4 if amp else 2e-5) * scale


def test_pendulum_squashed_full_size() -> None:
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig
    from rl8_b200.distributions import SquashedNormal

    N, T = 262144, 64
    torch.manual_seed(1)
    algo = AlgorithmConfig(num_envs=N, horizon=T, enable_amp=True, distribution_cls=SquashedNormal).build(E.Pendulum)
    c = algo.collect()
    obs = algo.buffer.hm["obs"]  # [T+1, 3, N]
    torch.testing.assert_close(obs[:, 0] ** 2 + obs[:, 1] ** 2, torch.ones_like(obs[:, 0]), rtol=0, atol=2e-6)
    assert float(obs[:, 2].abs().max()) <= 8.0
    act = algo.buffer.hm["actions"][:T]
    assert float(act.abs().max()) <= 1.0
    assert float(algo.buffer.hm["rewards"][:T].max()) <= 0.0
    assert c["env/steps"] == N * T and math.isfinite(c["returns/mean"])
    s = algo.step()
    assert all(math.isfinite(v) for v in s.values()), s
    assert s["losses/entropy"] == 0.0 and s["coefficients/entropy"] == 0.0


def test_cartpole_lstm_full_size_bf16() -> None:
    import rl8_b200.env as E
    from rl8_b200 import RecurrentAlgorithmConfig

    N, T = 65536, 32
    torch.manual_seed(2)
    algo = RecurrentAlgorithmConfig(num_envs=N, horizon=T, enable_amp=True, num_sgd_iters=1).build(E.CartPole)
    c = algo.collect()
    h = algo.buffer.hm["hidden_states"]
    cst = algo.buffer.hm["cell_states"]
    assert float(h.abs().max()) < 1.0 and math.isfinite(float(cst.abs().max()))
    # states are re-initialised every seqs_per_state_reset * seq_len = 32 steps: slab 0 is zero
    assert float(h[0].abs().max()) == 0.0 and float(h[1].abs().max()) > 0.0
    assert algo.state.seqs == T // 4
    assert c["env/steps"] == N * T and math.isfinite(c["returns/mean"])
    s = algo.step()
    assert all(math.isfinite(v) for v in s.values()), s
    assert s["losses/vf"] > 0.0


def test_cartpole_config5_share_is_deterministic() -> None:
    N, T = 131072, 32
    a = _cartpole_algo(N, T, True, seed=3, num_sgd_iters=1)
    b = _cartpole_algo(N, T, True, seed=3, num_sgd_iters=1)
    b.policy.model.flat_params.copy_(a.policy.model.flat_params)
    a.collect(), b.collect()
    assert torch.equal(a.buffer.hm["obs"], b.buffer.hm["obs"])
    assert torch.equal(a.buffer.hm["actions"], b.buffer.hm["actions"])
    sa, sb = a.step(), b.step()
    for k in ("losses/policy", "losses/vf", "monitors/kl_div"):
        assert sa[k] == pytest.approx(sb[k], rel=1e-5, abs=1e-7), k
    # gradients are summed with atomics: the Adam step agrees to a small fraction of lr = 1e-3
    assert float((a.policy.model.flat_params - b.policy.model.flat_params).abs().max()) < 2e-4
