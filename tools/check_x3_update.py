"""RL8_PREC_FP32_TC update (split-operand pair kernels) beside the CUDA-core fp32 update on the SAME buffer and
weights: loss statistics and every first-step gradient tensor, printed as relative errors.

    python tools/check_x3_update.py [N] [T]          (RL8_X3_STAGES=1|3|7 limits the kernels that run)
"""

from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rl8_b200.env as E  # noqa: E402
from rl8_b200 import AlgorithmConfig, _lib  # noqa: E402
from rl8_b200 import distributions as Dm  # noqa: E402

CASES = [
    ("CartPole", None, {"entropy_coeff": 0.01}),
    ("Pendulum", Dm.SquashedNormal, {"dual_clip_param": 3.0}),
    ("ContinuousDummyEnv", Dm.Normal, {"entropy_coeff": 0.01}),
    ("MountainCar", None, {"sgd_minibatch_size": 2048, "accumulate_grads": True}),
]


def twins(env_name, dist, n, t, **kw):  # noqa: ANN001, ANN201
    algos = []
    for prec in (_lib.PREC_FP32, _lib.PREC_FP32_TC):
        torch.manual_seed(7)
        a = AlgorithmConfig(num_envs=n, horizon=t, distribution_cls=dist, shuffle_minibatches=False,
                            num_sgd_iters=1, **kw).build(getattr(E, env_name))
        a.policy.precision = prec
        algos.append(a)
    ref, tc = algos
    tc.policy.model.load_state_dict(ref.policy.model.state_dict())
    torch.manual_seed(8)
    ref.collect()
    tc.buffer._raw.copy_(ref.buffer._raw)
    tc.state.buffered, tc.state.horizons = True, ref.state.horizons
    tc.state.reward_scale = ref.state.reward_scale
    return ref, tc


def main() -> None:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    t = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    for env_name, dist, kw in CASES:
        ref, tc = twins(env_name, dist, n, t, **kw)
        grads: list[dict[str, torch.Tensor]] = [{}, {}]
        for algo, g in zip((ref, tc), grads):
            algo._on_grads = (lambda named, g=g: g.update({k: v.detach().double().cpu().clone()
                                                           for k, v in named.items()}) if not g else None)
        s_ref, s_tc = ref.step(), tc.step()
        print(f"== {env_name} N={n} T={t} {kw}")
        for k in ("losses/policy", "losses/vf", "losses/total", "losses/entropy", "monitors/kl_div"):
            d = abs(s_tc[k] - s_ref[k]) / max(abs(s_ref[k]), 1e-12)
            print(f"   {k:18s} ref {s_ref[k]: .8e}  tc {s_tc[k]: .8e}  rel {d:.2e}")
        gn = float(torch.cat([v.flatten() for v in grads[0].values()]).norm())
        for k in sorted(grads[0]):
            a, b = grads[0][k], grads[1][k]
            print(f"   grad {k:28s} |ref| {float(a.norm()):.3e}  err/|ref| {float((a - b).norm() / max(a.norm(), 1e-30)):.2e}"
                  f"  err/|all| {float((a - b).norm()) / gn:.2e}  max|err| {float((a - b).abs().max()):.2e}")


if __name__ == "__main__":
    main()
