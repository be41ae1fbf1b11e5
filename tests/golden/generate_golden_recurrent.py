"""Golden vectors for the RECURRENT path from the UNMODIFIED upstream reference.

Container-only tool (the reference does not exist on the GPU box)::

    PYTHONPATH=oracle/refshim:/root/reference/src:/root/reference TORCHDYNAMO_DISABLE=1 \
        python tests/golden/generate_golden_recurrent.py

For every case: build the reference ``RecurrentAlgorithm`` through
``RecurrentAlgorithmConfig(...).build(env_cls)``, inject the initial env state and the
per-step sampling noise through the reference's plug-in points (see
``generate_golden.py``), run ``collect()`` + ``step()`` (reference code, torch fp32, eager
CPU, ``nn.LSTM``), replay the same inputs through ``oracle/recurrent_oracle.py`` and assert
agreement (discrete actions / counters exact, fp32 values to 1e-5 relative: the
reference's ``nn.LSTM`` runs in oneDNN whose summation order differs from the oracle's
explicit gates by ~1 ulp per step), then store inputs + reference outputs in
``tests/golden/rec_<case>.npz``.
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
sys.path.insert(0, HERE)
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

from generate_golden import (  # noqa: E402
    DIST_INFO,
    ENV_INFO,
    SUBSAMPLE,
    _NoiseQueue,
    injected_env,
)
from rl8 import RecurrentAlgorithmConfig  # noqa: E402  (upstream)
from rl8.data import DataKeys  # noqa: E402

from oracle import ppo_oracle as O  # noqa: E402
from oracle import recurrent_oracle as R  # noqa: E402

RTOL, ATOL = 1e-5, 2e-6


def close(name: str, a: torch.Tensor, b: torch.Tensor, rtol: float = RTOL, atol: float = ATOL) -> None:
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.dtype in (torch.int64, torch.int32):
        assert torch.equal(a, b), f"{name}: integer mismatch"
        return
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol, msg=lambda m: f"{name}: {m}")


def flat_buffer(buf: Any) -> dict[str, torch.Tensor]:
    out: dict[str, torch.Tensor] = {}
    for k, v in buf.items():
        if k == DataKeys.STATES:
            for kk, vv in v.items():
                out[kk] = vv
        else:
            out[k] = v
    return out


def run_case(
    name: str,
    env_name: str,
    dist_name: str,
    *,
    N: int,
    T: int,
    seed: int,
    rounds: int = 1,
    horizons_per_env_reset: int = 1,
    **algo_kwargs: Any,
) -> None:
    print(f"== {name}")
    torch.manual_seed(seed)
    ref_env_cls = ENV_INFO[env_name]
    states: list[torch.Tensor] = []
    env_cls = injected_env(ref_env_cls, states)
    algo_kwargs = dict(algo_kwargs)
    algo_kwargs.setdefault("shuffle_minibatches", False)
    algo = RecurrentAlgorithmConfig(
        num_envs=N,
        horizon=T,
        horizons_per_env_reset=horizons_per_env_reset,
        distribution_cls=DIST_INFO[dist_name],
        device="cpu",
        **algo_kwargs,
    ).build(env_cls)
    hp = algo.hparams
    params0 = {k: v.detach().clone() for k, v in algo.policy.model.state_dict().items()}
    # validate() leaves junk in the buffer (incl. states[:, 1]); the product starts clean
    # and the reference overwrites everything it reads, except states[:, -1] which is zero.
    o_env = O.OracleEnv(env_name, N)
    o_p = {k: v.clone() for k, v in params0.items()}
    o_dist = O.Dist(dist_name)
    o_buf = R.new_recurrent_buffer(N, T, o_env.obs_dim, o_env.action_kind)
    o_opt: dict[str, Any] = {}
    o_seqs = 0
    A = o_env.num_actions

    out: dict[str, np.ndarray] = {f"param0/{k}": v.numpy() for k, v in params0.items()}
    meta = dict(
        env=env_name, dist=dist_name, N=N, T=T, rounds=rounds,
        horizons_per_env_reset=horizons_per_env_reset, subsample=SUBSAMPLE,
        seq_len=hp.seq_len, seqs_per_state_reset=hp.seqs_per_state_reset,
        **{k: v for k, v in algo_kwargs.items() if k not in ("seq_len", "seqs_per_state_reset")},
    )
    gen = torch.Generator().manual_seed(seed + 1000)
    for rnd in range(rounds):
        will_reset = (algo.state.horizons % horizons_per_env_reset) == 0
        if will_reset:
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

        # ---- reference collect ---------------------------------------------------------
        cstats = algo.collect()
        assert not _NoiseQueue.items
        ref_buf = flat_buffer(algo.buffer)
        for k, v in ref_buf.items():
            out[f"r{rnd}/collect/{k}"] = v.detach().numpy().copy()
        for k, v in cstats.items():
            if not k.startswith("profiling"):
                out[f"r{rnd}/collect_stats/{k}"] = np.float64(v)
        out[f"r{rnd}/reward_scale"] = np.float64(algo.state.reward_scale)
        out[f"r{rnd}/seqs"] = np.int64(algo.state.seqs)

        # ---- oracle collect ------------------------------------------------------------
        o_stats, o_seqs = R.collect_recurrent(
            o_p, o_env, o_buf, o_dist, noise,
            seqs=o_seqs, seq_len=hp.seq_len, seqs_per_state_reset=hp.seqs_per_state_reset,
            gamma=hp.gamma, reset=will_reset, reset_state=state0 if will_reset else None,
        )
        assert o_seqs == algo.state.seqs
        for k in ("actions", "obs", "rewards", "logp", "values", "reversed_discounted_returns",
                  "hidden_states", "cell_states"):
            close(f"{name}/r{rnd}/collect/{k}", o_buf[k], ref_buf[k])
        for k, v in o_stats.items():
            ref_v = algo.state.reward_scale if k == "reward_scale" else cstats[k]
            assert abs(v - ref_v) <= 1e-5 * max(1.0, abs(ref_v)), (k, v, ref_v)

        # ---- reference step; record the first optimizer step's gradients ------------------
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

        # ---- oracle step -----------------------------------------------------------------
        o_grads: dict[str, torch.Tensor] = {}

        def hook(g: dict[str, torch.Tensor]) -> None:
            if not o_grads:
                o_grads.update(g)

        o_sstats = R.step_recurrent(
            o_p, o_buf, o_dist, o_opt,
            reward_scale=algo.state.reward_scale, seq_len=hp.seq_len,
            gamma=hp.gamma, gae_lambda=hp.gae_lambda,
            normalize_advantages=hp.normalize_advantages,
            sgd_minibatch_size=hp.sgd_minibatch_size, num_sgd_iters=hp.num_sgd_iters,
            shuffle=False, accumulate_grads=hp.accumulate_grads, clip_param=hp.clip_param,
            dual_clip_param=hp.dual_clip_param, entropy_coeff=algo.entropy_scheduler.coeff,
            vf_clip_param=hp.vf_clip_param, vf_coeff=hp.vf_coeff,
            target_kl_div=hp.target_kl_div, max_grad_norm=hp.max_grad_norm, grad_hook=hook,
        )
        for k in grads_first:
            scale = float(grads_first[k].abs().max())
            close(f"{name}/r{rnd}/grad/{k}", o_grads[k], grads_first[k], rtol=1e-4, atol=1e-5 * scale + 1e-9)
        for k in params1:
            # Adam normalises tiny gradients to +-lr steps: compare at a fraction of lr.
            close(f"{name}/r{rnd}/param1/{k}", o_p[k], params1[k], rtol=1e-4, atol=2e-4)
        for k, v in o_sstats.items():
            assert abs(v - sstats[k]) <= 1e-4 * max(1.0, abs(sstats[k])), (k, v, sstats[k])
        print(f"   round {rnd}: collect+step oracle ~= reference")
        print("   ", {k: round(float(v), 6) for k, v in sstats.items() if "profiling" not in k})

    out["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


def main() -> None:
    torch.set_num_threads(1)
    run_case("rec_discrete_dummy", "discrete_dummy", "categorical", N=32, T=8, seed=11,
             seq_len=4, seqs_per_state_reset=2, entropy_coeff=0.01, num_sgd_iters=2)
    run_case("rec_cartpole", "cartpole", "categorical", N=32, T=16, seed=12, rounds=2,
             horizons_per_env_reset=2, seq_len=4, seqs_per_state_reset=8, num_sgd_iters=2,
             sgd_minibatch_size=64)
    run_case("rec_pendulum_squashed", "pendulum", "squashed_normal", N=32, T=8, seed=13,
             seq_len=2, seqs_per_state_reset=4, num_sgd_iters=2, sgd_minibatch_size=64,
             accumulate_grads=True)
    run_case("rec_continuous_dummy_normal", "continuous_dummy", "normal", N=32, T=8, seed=14,
             seq_len=4, seqs_per_state_reset=-1, entropy_coeff=0.01, num_sgd_iters=2,
             dual_clip_param=5.0, rounds=2)


if __name__ == "__main__":
    main()
