"""GPU: the tensor-core (tcgen05, bf16 operands / fp32 accumulate) path.

bf16 GEMMs cannot match the fp32 reference to 1e-5, so parity is teacher-forced
(SURVEY.md §7 "hard parts"): every stage is checked against the CPU oracle evaluated on
the SAME inputs with the SAME operand rounding (H1 and W2 rounded to bf16, fp32 accumulate),
and everything that does not depend on the GEMM (env transitions given the recorded actions,
reward bookkeeping) is checked at the fp32 tolerance.
"""

from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """Round to nearest (ties away from zero) onto the tf32 grid: PTX cvt.rna.tf32.f32."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def emulated_forward(p: dict[str, torch.Tensor], obs: torch.Tensor, tf32_layer1: bool = False):
    """Oracle forward with the tensor-core path's operand rounding.  ``tf32_layer1``: layer 1 as the
    tensor-core kernels compute it (one kind::tf32 MMA: observations, W1 and b1 rounded to tf32, fp32
    accumulation); otherwise plain fp32.  The two differ by a quarter of a bf16 ulp of H1."""
    def net(prefix: str) -> torch.Tensor:
        w1, b1 = p[f"{prefix}.0.0.weight"], p[f"{prefix}.0.0.bias"]
        if tf32_layer1:
            z1 = (tf32_round(obs).double() @ tf32_round(w1).double().T + tf32_round(b1).double()).float()
            h1 = F.relu(z1)
        else:
            h1 = F.relu(F.linear(obs, w1, b1))
        z2 = F.linear(bf16_round(h1).double(), bf16_round(p[f"{prefix}.0.2.weight"]).double()).float()
        return F.relu(z2 + p[f"{prefix}.0.2.bias"])

    value = F.linear(net("vf_model"), p["vf_model.2.weight"], p["vf_model.2.bias"])
    if "feature_model.2.weight" in p:
        logits = F.linear(net("feature_model"), p["feature_model.2.weight"], p["feature_model.2.bias"])
        return {"logits": logits.reshape(-1, 1, logits.shape[-1])}, value
    z = net("latent_model")
    mean = F.linear(z, p["action_mean.weight"], p["action_mean.bias"])
    raw = F.linear(z, p["action_log_std.weight"], p["action_log_std.bias"])
    return {"mean": mean, "log_std": torch.tanh(raw)}, value


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("N,K", [(256, 256), (256, 128), (128, 64), (8, 128), (16, 16)])
def test_tcgen05_descriptor_selftest(a_mn: int, b_mn: int, N: int, K: int) -> None:
    from rl8_b200 import _lib as L

    lib = L.load()
    gen = torch.Generator().manual_seed(N * 7 + K + a_mn * 2 + b_mn)
    A = torch.randn(128, K, generator=gen)
    B = torch.randn(N, K, generator=gen)
    Ad, Bd = A.to(DEV), B.to(DEV)
    D = torch.full((128, N), float("nan"), device=DEV)
    rc = lib.rl8_tc_selftest(L.ptr(Ad), L.ptr(Bd), L.ptr(D), N, K, a_mn, b_mn, L.stream())
    assert rc == 0
    ref = (bf16_round(A).double() @ bf16_round(B).double().T).float()
    torch.testing.assert_close(D.cpu(), ref, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("N", [16, 128, 256])
def test_tcgen05_tf32_selftest(N: int) -> None:
    from rl8_b200 import _lib as L

    lib = L.load()
    gen = torch.Generator().manual_seed(N)
    A = torch.randn(128, 8, generator=gen) * 3
    B = torch.randn(N, 8, generator=gen)
    Ad, Bd = A.to(DEV), B.to(DEV)
    D = torch.full((128, N), float("nan"), device=DEV)
    assert lib.rl8_tc_selftest_tf32(L.ptr(Ad), L.ptr(Bd), L.ptr(D), N, L.stream()) == 0
    ref = (tf32_round(A).double() @ tf32_round(B).double().T).float()
    torch.testing.assert_close(D.cpu(), ref, rtol=1e-5, atol=1e-5)


def test_tmem_store_load_roundtrip() -> None:
    from rl8_b200 import _lib as L

    lib = L.load()
    x = torch.randint(-(2**31), 2**31 - 1, (128, 128), dtype=torch.int64).to(torch.int32).to(DEV)
    y = torch.zeros_like(x)
    assert lib.rl8_tc_selftest_tmem(L.ptr(x), L.ptr(y), L.stream()) == 0
    assert torch.equal(x.cpu(), y.cpu())


def test_tmem_16x256b_register_layout() -> None:
    """tcgen05.ld.16x256b.x4: thread (g = lane / 4, t = lane % 4) of a warp receives, from the 16 lanes it addresses,
    rows {g, g + 8} and the column pairs {8 k + 2 t, 8 k + 2 t + 1} -- the layout x3_update_f_kernel's gW3 pass
    assumes (rl8_b200/csrc/tc.cuh: tmem_ld_16x256b_x4)."""
    from rl8_b200 import _lib as L

    lib = L.load()
    x = (torch.arange(128).view(128, 1) * 256 + torch.arange(32).view(1, 32)).to(torch.int32).to(DEV)
    y = torch.zeros(128, 32, dtype=torch.int32, device=DEV)
    assert lib.rl8_tc_selftest_tmem_16x256b(L.ptr(x), L.ptr(y), L.stream()) == 0
    y = y.cpu()
    for tid in range(128):
        warp, lane = divmod(tid, 32)
        g, t = divmod(lane, 4)
        for h in range(2):
            for k in range(4):
                for e in range(4):
                    row = 32 * warp + 16 * h + g + 8 * (e // 2)
                    col = 8 * k + 2 * t + (e % 2)
                    assert int(y[tid, 16 * h + 4 * k + e]) == row * 256 + col


def _algo(env_name: str, dist=None, n: int = 256, t: int = 8, **kw):
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig

    return AlgorithmConfig(num_envs=n, horizon=t, enable_amp=True, distribution_cls=dist, **kw).build(
        getattr(E, env_name)
    )


@pytest.mark.parametrize("env_name", ["CartPole", "Pendulum", "DiscreteDummyEnv"])
@pytest.mark.parametrize("rows", [1, 130, 5000])
def test_forward_matches_emulated_oracle(env_name: str, rows: int) -> None:
    torch.manual_seed(1)
    algo = _algo(env_name)
    pol = algo.policy
    params = {k: v.detach().cpu().clone() for k, v in pol.model.state_dict().items()}
    D = algo.env.observation_spec.shape[0]
    obs = torch.randn(rows, D) * 2
    feats, value = emulated_forward(params, obs, tf32_layer1=True)
    out = pol.sample({"obs": obs.to(DEV).unsqueeze(1)}, return_actions=False, return_values=True)
    # An H1 element that sits on a bf16 rounding boundary may round the other way on the GPU
    # (fmaf chain vs the CPU's dot product): allow one bf16 ulp of one hidden unit.
    for k, ref in feats.items():
        torch.testing.assert_close(out["features"][k].cpu(), ref, rtol=2e-3, atol=5e-4)
    torch.testing.assert_close(out["values"].cpu(), value, rtol=2e-3, atol=5e-4)
    # and the bf16 path stays close to the fp32 oracle (operand rounding only)
    f32_feats, f32_value = O.model_forward(params, obs)
    torch.testing.assert_close(out["values"].cpu(), f32_value, rtol=3e-2, atol=3e-3)


@pytest.mark.parametrize(
    "env_name,oname,dist",
    [("CartPole", "cartpole", "categorical"), ("MountainCar", "mountain_car", "categorical"),
     ("Pendulum", "pendulum", "squashed_normal"), ("ContinuousDummyEnv", "continuous_dummy", "normal"),
     ("DiscreteDummyEnv", "discrete_dummy", "categorical")],
)
@pytest.mark.parametrize("n", [256, 300])
def test_rollout_kernel_teacher_forced(env_name: str, oname: str, dist: str, n: int) -> None:
    import rl8_b200.env as E
    from rl8_b200 import distributions as Dm

    T = 12
    gen = torch.Generator().manual_seed(n)
    dcls = {"categorical": Dm.Categorical, "normal": Dm.Normal, "squashed_normal": Dm.SquashedNormal}[dist]
    P = {"cartpole": 3, "mountain_car": 3, "discrete_dummy": 2}.get(oname, 1)
    noise = (torch.empty(T, n, 1, P).exponential_(1, generator=gen) if dist == "categorical"
             else torch.randn(T, n, 1, generator=gen))

    class Inj(dcls):  # type: ignore[misc, valid-type]
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            if steps != T:
                return super().draw_noise(steps, num, width, device)
            return noise.reshape(steps, num, -1).squeeze(-1).contiguous().to(device) if dist != "categorical" \
                else noise.reshape(steps, num, width).to(device)

    torch.manual_seed(0)
    algo = _algo(env_name, Inj, n=n, t=T)
    params = {k: v.detach().cpu().clone() for k, v in algo.policy.model.state_dict().items()}
    algo.collect()
    buf = {k: algo.buffer[k].cpu() for k in algo.buffer.keys()}
    # (1) env transitions / rewards / rdr replayed on the CPU oracle with the recorded actions
    oenv = O.OracleEnv(oname, n)
    first_obs = buf["obs"][:, 0]
    # recover the initial state the env kernel produced (state is not in the buffer): replay from
    # the env's own reset is not possible, so compare transitions obs[t] -> obs[t+1] instead for
    # envs whose observation determines the state.
    if oname in ("discrete_dummy", "continuous_dummy", "mountain_car"):
        state = first_obs.T.contiguous() if oname == "mountain_car" else first_obs.clone()
        oenv.reset(state)
        rdr = torch.zeros(n, 1)
        for t in range(T):
            o_obs, o_r = oenv.step(buf["actions"][:, t])
            torch.testing.assert_close(buf["obs"][:, t + 1], o_obs.reshape(n, -1), rtol=1e-5, atol=2e-6)
            torch.testing.assert_close(buf["rewards"][:, t], o_r, rtol=1e-5, atol=2e-6)
            rdr = 0.95 * rdr + o_r
            torch.testing.assert_close(buf["reversed_discounted_returns"][:, t + 1], rdr, rtol=1e-5, atol=1e-5)
    else:
        # state is recoverable from the observation (theta = atan2(sin, cos)): check every
        # single transition obs[t] --action[t]--> obs[t+1], reward[t] on its own
        for t in range(T):
            ob = buf["obs"][:, t]
            if oname == "cartpole":
                state = torch.stack((ob[:, 0], ob[:, 1], torch.atan2(ob[:, 3], ob[:, 2]), ob[:, 4]))
            else:
                state = torch.stack((torch.atan2(ob[:, 1], ob[:, 0]), ob[:, 2]))
            oenv.reset(state)
            o_obs, o_r = oenv.step(buf["actions"][:, t])
            torch.testing.assert_close(buf["obs"][:, t + 1], o_obs.reshape(n, -1), rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(buf["rewards"][:, t], o_r, rtol=1e-4, atol=1e-5)
    # (2) policy outputs recomputed on the recorded observations with the same operand rounding
    flat_obs = buf["obs"].reshape(n * (T + 1), -1)
    feats, value = emulated_forward(params, flat_obs, tf32_layer1=True)  # every tensor-core kernel: tf32 layer 1
    torch.testing.assert_close(buf["values"].reshape(-1, 1), value, rtol=2e-3, atol=5e-4)
    d = O.Dist(dist).bind({k: v.reshape(n, T + 1, *v.shape[1:])[:, :T].reshape(n * T, *v.shape[1:]) for k, v in feats.items()})
    nz = noise.permute(1, 0, 2, 3).reshape(n * T, 1, P) if dist == "categorical" else noise.permute(1, 0, 2).reshape(n * T, 1)
    ref_actions = d.sample(nz)
    got_actions = buf["actions"][:, :T].reshape(n * T, 1)
    if dist == "categorical":
        mism = int((ref_actions != got_actions).sum())
        assert mism <= max(1, n * T // 500), f"{mism} of {n * T} discrete actions differ"
        torch.testing.assert_close(buf["logp"][:, :T].reshape(-1, 1), d.logp(got_actions), rtol=2e-3, atol=5e-4)
    else:
        torch.testing.assert_close(got_actions, ref_actions, rtol=2e-3, atol=5e-4)
        torch.testing.assert_close(buf["logp"][:, :T].reshape(-1, 1), d.logp(got_actions), rtol=5e-3, atol=5e-3)


def _twin_algos(env_name: str, dist=None, n: int = 512, t: int = 16, **kw):
    """An fp32 and a bf16 algorithm with identical parameters and identical rollout buffers."""
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig

    algos = []
    for amp in (False, True):
        torch.manual_seed(7)
        algos.append(
            AlgorithmConfig(num_envs=n, horizon=t, enable_amp=amp, distribution_cls=dist,
                            shuffle_minibatches=False, **kw).build(getattr(E, env_name))
        )
    a32, a16 = algos
    # W2 is rounded to bf16 in BOTH twins: the systematic part of the operand rounding (a fixed
    # perturbation of the weights, which shifts e.g. the mean value prediction) then cancels and
    # what is left is the unbiased rounding of the activation tiles.
    sd = {k: (bf16_round(v) if k.endswith(".0.2.weight") else v.clone())
          for k, v in a32.policy.model.state_dict().items()}
    a32.policy.model.load_state_dict(sd)
    a16.policy.model.load_state_dict(sd)
    torch.manual_seed(8)
    a32.collect()
    a16.buffer._raw.copy_(a32.buffer._raw)
    a16.state.buffered, a16.state.horizons = True, a32.state.horizons
    a16.state.reward_scale = a32.state.reward_scale
    return a32, a16


@pytest.mark.parametrize(
    "env_name,dist,kw",
    [("CartPole", None, {"entropy_coeff": 0.01}),
     ("MountainCar", None, {"sgd_minibatch_size": 2048, "accumulate_grads": True}),
     ("Pendulum", "squashed_normal", {"dual_clip_param": 3.0}),
     ("ContinuousDummyEnv", "normal", {"entropy_coeff": 0.01}),
     ("DiscreteDummyEnv", None, {})],
)
def test_update_kernels_match_fp32_path(env_name: str, dist, kw) -> None:
    """Fused tcgen05 forward/loss/backward vs the fp32 kernels on the SAME buffer and weights:
    losses to 1e-3, every gradient tensor to ~1% (bf16 operand rounding), cosine > 0.999."""
    from rl8_b200 import distributions as Dm

    dcls = {None: None, "normal": Dm.Normal, "squashed_normal": Dm.SquashedNormal}[dist]
    a32, a16 = _twin_algos(env_name, dcls, num_sgd_iters=1, **kw)
    grads: list[dict[str, torch.Tensor]] = [{}, {}]
    for algo, g in zip((a32, a16), grads):
        algo._on_grads = lambda named, g=g: g.update({k: v.detach().cpu().clone() for k, v in named.items()}) if not g else None
    s32, s16 = a32.step(), a16.step()
    # bf16 operands: 2^-9 relative per hidden unit; the dummy envs feed |obs| up to 100, so
    # their value loss moves by a few 1e-3 relative
    for k in ("losses/policy", "losses/vf", "losses/total", "losses/entropy", "monitors/kl_div"):
        assert s16[k] == pytest.approx(s32[k], rel=1e-2, abs=5e-4), (k, s16[k], s32[k])
    assert set(grads[0]) == set(grads[1]) and grads[0]
    keys = sorted(grads[0])
    full32 = torch.cat([grads[0][k].flatten() for k in keys]).double()
    full16 = torch.cat([grads[1][k].flatten() for k in keys]).double()
    gnorm = float(full32.norm())
    assert float((full16 - full32).norm()) / gnorm < 2e-2
    assert float((full16 * full32).sum() / (gnorm * full16.norm())) > 0.9995
    for k in keys:
        g32, g16 = grads[0][k].double(), grads[1][k].double()
        err = float((g16 - g32).norm())
        # per tensor: 5 % of its own norm, or (tiny / cancelling tensors) 0.5 % of the global norm
        assert err < max(5e-2 * float(g32.norm()), 5e-3 * gnorm), (k, err, float(g32.norm()), gnorm)
    p32 = a32.policy.model.flat_params.cpu()
    p16 = a16.policy.model.flat_params.cpu()
    assert float((p32 - p16).abs().max()) < 2.5e-3  # one Adam step moves each weight by <= lr = 1e-3


def test_ragged_minibatch_tiles_and_shuffle_bf16() -> None:
    """Row counts that are not multiples of the 128-row tile, shuffled minibatches."""
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig

    torch.manual_seed(3)
    algo = AlgorithmConfig(num_envs=100, horizon=9, enable_amp=True, num_sgd_iters=2,
                           sgd_minibatch_size=300).build(E.CartPole)
    for _ in range(2):
        c = algo.collect()
        s = algo.step()
        assert all(v == v and abs(v) < 1e6 for v in s.values()), s
        assert c["env/steps"] == 900


def test_bf16_training_improves_returns() -> None:
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig, Trainer

    torch.manual_seed(0)
    algo = AlgorithmConfig(num_envs=2048, horizon=16, enable_amp=True, sgd_minibatch_size=8192).build(
        E.DiscreteDummyEnv
    )
    trainer = Trainer(algo)
    first = trainer.step(env_config={"bounds": 4.0})["returns/mean"]
    last = first
    for _ in range(30):
        last = trainer.step(env_config={"bounds": 4.0})["returns/mean"]
    assert last > first + 1.0, (first, last)


# ---------------------------------------------------------------------------------------------------
# generic tensor-core GEMM (gemm_tc.cu) and the recurrent path on it (enable_amp=True)
# ---------------------------------------------------------------------------------------------------


@pytest.mark.parametrize(
    "a_k,b_k,M,N,K,splits",
    [
        (1, 1, 300, 1024, 256, 1),    # gates = h W_hh^T            (ragged M)
        (1, 1, 1000, 1024, 256, 1),   # ... at M >= 512: the B-resident variant (ragged last tile)
        (1, 1, 40000, 1000, 256, 1),  # ... many tiles per CTA, ragged N
        (1, 0, 257, 256, 1024, 1),    # dh = dG W_hh
        (0, 0, 1024, 256, 1000, 7),   # gW_hh += dG^T h             (split-K over ragged K)
        (1, 1, 128, 256, 64, 1),
        (0, 0, 128, 252, 72, 2),      # N, K tails
        (1, 0, 5, 8, 16, 1),
    ],
)
def test_tc_gemm_matches_bf16_emulation(a_k: int, b_k: int, M: int, N: int, K: int, splits: int) -> None:
    """``rl8_tc_gemm`` = fp32 GEMM of the bf16-rounded operands (fp32 accumulation): 2e-4 of the
    result's scale against a float64 product of the rounded operands, both operand majors, ragged
    M / N / K, split-K accumulation on top of existing contents."""
    from rl8_b200 import _lib as L

    lib = L.load()
    g = torch.Generator(device=DEV).manual_seed(M * 7 + K)
    A = torch.randn((M, K) if a_k else (K, M), generator=g, device=DEV)
    B = torch.randn((N, K) if b_k else (K, N), generator=g, device=DEV)
    base = torch.randn(M, N, generator=g, device=DEV)
    C = base.clone()
    acc = int(splits > 1)
    rc = lib.rl8_tc_gemm(a_k, b_k, acc, L.ptr(A), L.ptr(B), L.ptr(C), M, N, K, A.stride(0), B.stride(0),
                         C.stride(0), splits, L.stream())
    assert rc == 0, rc
    Ar = (A if a_k else A.T).bfloat16().double()
    Br = (B if b_k else B.T).bfloat16().double()
    ref = Ar @ Br.T + (base.double() if acc else 0)
    err = float((C.double() - ref).abs().max())
    assert err < 2e-4 * float(ref.abs().max()), (err, float(ref.abs().max()))


@pytest.mark.parametrize("N", [192, 1100])
@pytest.mark.parametrize("env_name,dist", [("CartPole", None), ("Pendulum", "squashed_normal")])
def test_recurrent_amp_matches_fp32_path(env_name: str, dist, N: int) -> None:
    """RecurrentAlgorithm with ``enable_amp=True`` (LSTM GEMMs in bf16 on tcgen05) against the fp32 path
    from the same weights, env states and noise: rollout values / states to bf16 accuracy, identical
    discrete actions except where two logits tie within bf16 noise, losses 1e-2, gradient cosine > 0.999.
    N = 192 runs the generic bf16 GEMMs; N = 1100 (>= 512 rows per step, a ragged last tile: 2 200 sequences) the fused
    kernels -- tc_lstm_cell_kernel forward, lstm_cell_bwd_tc / lstm_dh_tc / lstm_wgrad_tc backward (lstm_tc.cu), with
    every gradient tensor compared one by one."""
    import rl8_b200.env as E
    from rl8_b200 import RecurrentAlgorithmConfig
    from rl8_b200 import distributions as Dm

    T, L = 8, 4
    env_cls = getattr(E, env_name)
    base = {None: None, "squashed_normal": Dm.SquashedNormal}[dist] or Dm.Categorical
    torch.manual_seed(11)
    width = 3 if base is Dm.Categorical else 1
    noise = torch.empty(T, N, width)
    noise = noise.exponential_(1) if base is Dm.Categorical else noise.normal_()
    S = 4 if env_name == "CartPole" else 2
    state0 = torch.randn(S, N) * 0.05

    class InjDist(base):  # type: ignore[misc, valid-type]
        @classmethod
        def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
            if steps != T:
                return super().draw_noise(steps, num, width, device)
            return noise.reshape(steps, num, -1).squeeze(-1).to(device) if base is not Dm.Categorical else noise.to(device)

    class InjEnv(env_cls):  # type: ignore[misc, valid-type]
        def reset(self, *, config=None):  # noqa: ANN001, ANN202
            super().reset(config=config)
            return self.set_state(state0.to(DEV))

    algos = []
    for amp in (False, True):
        torch.manual_seed(5)
        algos.append(RecurrentAlgorithmConfig(num_envs=N, horizon=T, seq_len=L, seqs_per_state_reset=2, enable_amp=amp,
                                              num_sgd_iters=1,
                                              distribution_cls=InjDist, shuffle_minibatches=False).build(InjEnv))
    a32, a16 = algos
    a16.policy.model.flat_params.copy_(a32.policy.model.flat_params)
    grads: list[dict[str, torch.Tensor]] = [{}, {}]
    stats = []
    for algo, g in zip(algos, grads):
        algo._on_grads = lambda named, g=g: g.update({k: v.detach().cpu().clone() for k, v in named.items()}) if not g else None
        algo.collect()
    v32, v16 = a32.buffer["values"].cpu(), a16.buffer["values"].cpu()
    assert float((v32 - v16).abs().max()) < 2e-2 * max(1.0, float(v32.abs().max()))
    h32 = a32.buffer["states"]["hidden_states"].cpu()
    h16 = a16.buffer["states"]["hidden_states"].cpu()
    assert float((h32 - h16).abs().max()) < 2e-2
    if base is Dm.Categorical:
        same = (a32.buffer["actions"] == a16.buffer["actions"]).float().mean()
        assert float(same) > 0.98
        # make the update inputs identical so the comparison isolates the update kernels
        for k in ("obs", "actions", "logp", "values", "rewards"):
            a16.buffer[k].copy_(a32.buffer[k])
        for k in ("hidden_states", "cell_states"):
            a16.buffer["states"][k].copy_(a32.buffer["states"][k])
        a16.state.reward_scale = a32.state.reward_scale
    for algo in algos:
        stats.append(algo.step())
    s32, s16 = stats
    for k in ("losses/policy", "losses/vf", "losses/total", "monitors/kl_div"):
        assert s16[k] == pytest.approx(s32[k], rel=2e-2, abs=1e-3), (k, s16[k], s32[k])
    keys = sorted(grads[0])
    assert keys and set(keys) == set(grads[1])
    if base is Dm.Categorical:
        full32 = torch.cat([grads[0][k].flatten() for k in keys]).double()
        full16 = torch.cat([grads[1][k].flatten() for k in keys]).double()
        cos = float((full16 * full32).sum() / (full32.norm() * full16.norm()))
        assert cos > 0.999, cos
        assert float((full16 - full32).norm() / full32.norm()) < 3e-2
        for k in keys:  # every tensor on its own (a wrong small tensor -- biases, W_ih, heads -- hides in the norm above)
            a, b = grads[0][k].double(), grads[1][k].double()
            assert float((a - b).norm()) <= 5e-2 * float(a.norm()) + 1e-7, (k, float((a - b).norm()), float(a.norm()))
