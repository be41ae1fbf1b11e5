"""GPU: the fp32-accurate tensor-core mode (RL8_PREC_FP32_TC: split operands -- two fp16 pieces per fp32 value in the
kernels, three bf16 pieces in the first form kept as a build option -- on tcgen05 pair MMAs).

The pair selftest pins the cta_group::2 plumbing and bounds each piece-product set against fp64;
the fused forward kernel is compared with the CPU oracle (the reference's fp32 ``nn.Linear`` chain,
src/rl8/models/_feedforward.py:263-375) at the north-star tolerance, at row counts on both sides of
every tile boundary (128-row CTA halves, 256-row pair tiles, more tiles than CTA pairs).
"""

from __future__ import annotations

import pytest
import torch

from oracle import ppo_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL, ATOL = 1e-5, 1e-6


def _lib():
    from rl8_b200 import _lib as L

    return L, L.load()


def _rel(got: torch.Tensor, ref: torch.Tensor) -> float:
    return float((got.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("K", [32, 256])
def test_pair_mma_term_sets_against_fp64(K: int) -> None:
    L, lib = _lib()
    gen = torch.Generator().manual_seed(K)
    A = torch.randn(256, K, generator=gen).to(DEV)
    B = torch.randn(256, K, generator=gen).to(DEV)
    ref = A.double() @ B.double().T
    err = {}
    for terms in (1, 7, 63):
        D = torch.full((256, 256), float("nan"), device=DEV)
        assert lib.rl8_tc3_selftest(L.ptr(A), L.ptr(B), L.ptr(D), K, terms, L.stream()) == 0
        torch.cuda.synchronize()
        err[terms] = _rel(D, ref)
    assert 2e-4 < err[1] < 8e-3, err      # one bf16 product: operands really are single pieces
    assert err[7] < 1.2e-5, err           # x2: two pieces, three products (gradient contractions)
    assert err[63] < 4e-6, err            # x3: three pieces, six products (forward contraction)
    # bit-reproducible
    D2 = torch.empty(256, 256, device=DEV)
    lib.rl8_tc3_selftest(L.ptr(A), L.ptr(B), L.ptr(D2), K, 63, L.stream())
    torch.cuda.synchronize()
    assert torch.equal(D, D2)


@pytest.mark.parametrize("K", [32, 256])
def test_pair_mma_fp16_pieces_against_fp64(K: int) -> None:
    """Two fp16 pieces per operand, three piece products (the kernels' form, split_tc.cuh): 22 bits per operand once
    the operands are scaled by a power of two into fp16's range -- also for elements 2^-12 of the largest one; the
    bound is the one of the six bf16 products (the fp32 accumulation in tensor memory dominates both)."""
    L, lib = _lib()
    gen = torch.Generator().manual_seed(100 + K)
    A = torch.randn(256, K, generator=gen)
    B = torch.randn(256, K, generator=gen)
    A[:, ::3] *= 2.0**-12  # wide dynamic range inside one operand
    B[::2] *= 2.0**-9
    A, B = A.to(DEV), B.to(DEV)
    ref = A.double() @ B.double().T
    sa = 2.0 ** (14 - int(torch.ceil(torch.log2(A.abs().max()))))
    sb = 2.0 ** (14 - int(torch.ceil(torch.log2(B.abs().max()))))
    err = {}
    for terms in (64 + 1, 64 + 7, 64 + 15):
        D = torch.full((256, 256), float("nan"), device=DEV)
        assert lib.rl8_tc3_selftest(L.ptr(A * sa), L.ptr(B * sb), L.ptr(D), K, terms, L.stream()) == 0
        torch.cuda.synchronize()
        err[terms] = _rel(D / (sa * sb), ref)
    assert 2e-5 < err[65] < 2e-3, err     # one fp16 product: operands really are single pieces
    assert err[71] < 4e-6, err            # the update kernels' three products
    assert err[79] < 4e-6, err
    # per-element: rows of B scaled down by 2^-9 keep their relative accuracy (fp16 subnormals do not bite)
    small = (D[:, ::2] / (sa * sb)).double()
    assert float((small - ref[:, ::2]).abs().max() / ref[:, ::2].abs().max()) < 4e-6
    assert lib.rl8_tc3_selftest(L.ptr(A), L.ptr(B), L.ptr(D), K, 64 + 16, L.stream()) != 0  # no third fp16 piece


def test_pair_mma_maps_rows_and_columns_of_both_ctas() -> None:
    """Exact integer-valued operands: every D[m][n] identifies its row and column (leader rows 0..127,
    peer rows 128..255; B half of the leader = columns 0..127)."""
    L, lib = _lib()
    K = 32
    A = torch.zeros(256, K)
    B = torch.zeros(256, K)
    A[:, 0] = torch.arange(256, dtype=torch.float32)  # row id
    A[:, 1] = 1.0
    B[:, 0] = 1.0
    B[:, 1] = torch.arange(256, dtype=torch.float32) * 256.0  # column id
    A, B = A.to(DEV), B.to(DEV)
    D = torch.empty(256, 256, device=DEV)
    assert lib.rl8_tc3_selftest(L.ptr(A), L.ptr(B), L.ptr(D), K, 63, L.stream()) == 0
    want = torch.arange(256.0).view(256, 1) + 256.0 * torch.arange(256.0).view(1, 256)
    assert torch.equal(D.cpu(), want)


def _policy(env_name: str, dist: None | str = None):
    import rl8_b200.env as E
    from rl8_b200 import distributions as Dm
    from rl8_b200.policies import Policy

    env = getattr(E, env_name)(8, 8, device=DEV)
    dcls = {None: None, "normal": Dm.Normal, "squashed_normal": Dm.SquashedNormal}[dist]
    return Policy(env.observation_spec, env.action_spec, distribution_cls=dcls, device=DEV)


@pytest.mark.parametrize("env_name,D", [("CartPole", 5), ("Pendulum", 3), ("DiscreteDummyEnv", 1)])
@pytest.mark.parametrize("rows", [1, 127, 128, 129, 256, 257, 1000, 19_001, 70_000])
def test_forward_split_matches_oracle(env_name: str, D: int, rows: int) -> None:
    L, _ = _lib()
    torch.manual_seed(rows + D)
    pol = _policy(env_name)
    pol.precision = L.PREC_FP32_TC
    params = {k: v.detach().cpu().clone() for k, v in pol.model.state_dict().items()}
    obs = torch.randn(rows, D) * 2
    feats, value = O.model_forward(params, obs)
    head = pol.forward_net(0, obs.to(DEV))
    vals = pol.forward_net(1, obs.to(DEV))
    ref_head = torch.cat([t.reshape(rows, -1) for t in feats.values()], dim=1)
    if "log_std" in feats:  # continuous head: the kernel applies tanh to column 1 like the model does
        ref_head = torch.cat([feats["mean"].reshape(rows, 1), feats["log_std"].reshape(rows, 1)], dim=1)
    torch.testing.assert_close(head.cpu(), ref_head, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(vals.cpu(), value.reshape(rows, 1), rtol=RTOL, atol=ATOL)
    # strided (horizon-major SoA) observations give the same numbers, bit for bit
    obs_soa = obs.T.contiguous().to(DEV)
    assert torch.equal(pol.forward_net(1, obs_soa.T), vals)


@pytest.mark.parametrize("obs_scale", [1e-4, 1.0, 3e3, 1e6])
def test_forward_split_operand_scales_follow_the_observation_range(obs_scale: float) -> None:
    """fp16 pieces: the power-of-two scale of H1 comes from max |obs| of the rows of the call, so observations of any
    magnitude -- and rows a thousand times smaller than the largest one in the same call -- keep the fp32 path's accuracy
    relative to the size of the outputs."""
    L, _ = _lib()
    torch.manual_seed(int(obs_scale * 7) % 1000 + 3)
    pol = _policy("CartPole")
    pol.precision = L.PREC_FP32_TC
    params = {k: v.detach().cpu().clone() for k, v in pol.model.state_dict().items()}
    rows = 1500
    obs = torch.randn(rows, 5) * obs_scale
    obs[::7] *= 1e-3
    feats, value = O.model_forward(params, obs)
    ref_head = torch.cat([t.reshape(rows, -1) for t in feats.values()], dim=1)
    head = pol.forward_net(0, obs.to(DEV)).cpu()
    vals = pol.forward_net(1, obs.to(DEV)).cpu()
    assert torch.isfinite(head).all() and torch.isfinite(vals).all()
    for got, ref in ((head, ref_head), (vals, value.reshape(rows, 1))):
        big = float(ref.abs().max())
        torch.testing.assert_close(got, ref, rtol=RTOL, atol=max(ATOL, 1e-5 * big))
        small = slice(0, rows, 7)  # the small rows on their own scale
        torch.testing.assert_close(got[small], ref[small], rtol=RTOL, atol=max(ATOL, 1e-5 * float(ref[small].abs().max())))


def test_forward_split_trained_scale_weights_against_fp64() -> None:
    """O(1) head weights and large activations (what a trained policy looks like): the split forward
    stays within the fp32 path's own distance from fp64."""
    L, _ = _lib()
    torch.manual_seed(11)
    pol = _policy("CartPole")
    with torch.no_grad():
        for name, p in pol.model.named_parameters():
            if name.startswith(("feature_model.2", "vf_model.2")):
                p.copy_(torch.randn_like(p) * 0.5)
    rows = 3000
    obs = (torch.randn(rows, 5) * 3).to(DEV)
    sd = {k: v.detach().double() for k, v in pol.model.state_dict().items()}

    def mlp64(prefix: str, head: str) -> torch.Tensor:
        x = obs.double()
        x = torch.relu(x @ sd[f"{prefix}.0.0.weight"].T + sd[f"{prefix}.0.0.bias"])
        x = torch.relu(x @ sd[f"{prefix}.0.2.weight"].T + sd[f"{prefix}.0.2.bias"])
        return x @ sd[f"{head}.weight"].T + sd[f"{head}.bias"]

    ref_pi, ref_vf = mlp64("feature_model", "feature_model.2"), mlp64("vf_model", "vf_model.2")
    out = {}
    for prec in (L.PREC_FP32, L.PREC_FP32_TC):
        pol.precision = prec
        out[prec] = (_rel(pol.forward_net(0, obs), ref_pi), _rel(pol.forward_net(1, obs), ref_vf))
    assert max(out[L.PREC_FP32_TC]) < 5e-6, out
    assert max(out[L.PREC_FP32]) < 5e-6, out


# ---------------------------------------------------------------------------------------------------------
# update: the three split kernels against the CUDA-core fp32 update on the SAME buffer and weights
# ---------------------------------------------------------------------------------------------------------


def _twins(env_name: str, dist, n: int, t: int, **kw):  # noqa: ANN001, ANN202
    import rl8_b200.env as E
    from rl8_b200 import AlgorithmConfig

    L, _ = _lib()
    algos = []
    for prec in (L.PREC_FP32, L.PREC_FP32_TC):
        torch.manual_seed(7)
        a = AlgorithmConfig(num_envs=n, horizon=t, distribution_cls=dist, **kw).build(getattr(E, env_name))
        a.policy.precision = prec
        algos.append(a)
    ref, tc = algos
    tc.policy.model.load_state_dict(ref.policy.model.state_dict())
    torch.manual_seed(8)
    tc.collect()  # the buffer comes from the split forward: logp_new == logp_old on the first minibatch
    ref.buffer._raw.copy_(tc.buffer._raw)
    ref.state.buffered, ref.state.horizons = True, tc.state.horizons
    ref.state.reward_scale = tc.state.reward_scale
    return ref, tc


@pytest.mark.parametrize(
    "env_name,dist,n,t,kw",
    [("CartPole", None, 1000, 16, {"entropy_coeff": 0.01}),
     ("CartPole", None, 300, 7, {"sgd_minibatch_size": 700, "shuffle_minibatches": True}),   # ragged tiles, row lists
     ("MountainCar", None, 1024, 16, {"sgd_minibatch_size": 2048, "accumulate_grads": True}),
     ("Pendulum", "squashed_normal", 777, 8, {"dual_clip_param": 3.0}),
     ("ContinuousDummyEnv", "normal", 512, 16, {"entropy_coeff": 0.01}),
     # Raw rewards / advantages: max |dOut| is two orders of magnitude above the normalised case (the dZ2 operand scale)
     # and every advantage has the same sign, so each policy gradient is a difference of sums ~10^3 times its size.  The
     # tensor pipe truncates its accumulator after every instruction where the CUDA cores round to nearest: measured
     # 1.8e-3 of the tensor's norm on the policy's gW2 -- the bar of this case is 5e-3, and Adam's normalisation of such
     # gradients is not compared.
     ("Pendulum", "normal", 512, 8, {"normalize_advantages": False, "normalize_rewards": False, "vf_coeff": 5.0,
                                     "_cancelling": True}),
     ("DiscreteDummyEnv", None, 70_000, 4, {})],                                               # more tiles than CTA pairs
)
def test_update_split_matches_fp32_cuda_cores(env_name: str, dist, n: int, t: int, kw) -> None:  # noqa: ANN001
    from rl8_b200 import distributions as Dm

    dcls = {None: None, "normal": Dm.Normal, "squashed_normal": Dm.SquashedNormal}[dist]
    kw = {"shuffle_minibatches": False, **kw}
    cancelling = kw.pop("_cancelling", False)
    ref, tc = _twins(env_name, dcls, n, t, num_sgd_iters=1, **kw)
    grads: list[dict[str, torch.Tensor]] = [{}, {}]
    for algo, g in zip((ref, tc), grads):
        algo._on_grads = (lambda named, g=g: g.update({k: v.detach().double().cpu().clone() for k, v in named.items()})
                          if not g else None)
    torch.manual_seed(9)
    s_ref = ref.step()
    torch.manual_seed(9)  # same minibatch permutation
    s_tc = tc.step()
    for k in ("losses/policy", "losses/vf", "losses/total", "losses/entropy", "monitors/kl_div"):
        assert s_tc[k] == pytest.approx(s_ref[k], rel=2e-5, abs=2e-6), (k, s_tc[k], s_ref[k])
    assert set(grads[0]) == set(grads[1]) and grads[0]
    gnorm = float(torch.cat([v.flatten() for v in grads[0].values()]).norm())
    for k in sorted(grads[0]):
        a, b = grads[0][k], grads[1][k]
        err = float((a - b).norm())
        # 2e-5 of the tensor's own norm, or -- tensors that are sums of cancelling terms -- 2e-6 of the global norm
        assert err <= max((5e-3 if cancelling else 2e-5) * float(a.norm()), 2e-6 * gnorm), (k, err, float(a.norm()), gnorm)
    if cancelling:
        return
    p_ref, p_tc = ref.policy.model.flat_params, tc.policy.model.flat_params
    # Adam turns a gradient error dg into lr * eps / (|g| + eps)^2 * dg: only elements with |g| ~ eps = 1e-8 move
    assert float((p_ref - p_tc).abs().max()) < 2e-4
    assert float(((p_ref - p_tc).abs() > 5e-6).float().mean()) < 2e-3


def test_fp32_modes_share_the_golden_case(monkeypatch: pytest.MonkeyPatch) -> None:
    """RL8_FP32_SIMT=1 selects the CUDA-core GEMMs for enable_amp=False (the cross-check of the split kernels)."""
    L, _ = _lib()
    assert L.precision_for(False) == L.PREC_FP32_TC and L.precision_for(True) == L.PREC_BF16
    monkeypatch.setenv("RL8_FP32_SIMT", "1")
    assert L.precision_for(False) == L.PREC_FP32
