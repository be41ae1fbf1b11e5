"""One CartPole ``collect()`` + ``step()`` at bench size for profilers, and per-kernel CUDA-event timings.

    python tools/profile_step.py [fp32|bf16] [num_envs] [horizon] [steps]

Plain run: prints the CUDA-event time of collect(), step() and of ONE rl8_ppo_minibatch call, and -- for the
fp32 tensor-core mode -- of each of its three kernels alone (RL8_X3_STAGES).  Under ``ncu`` the same command gives
the launch list (`--metrics gpu__time_duration.sum`) or the full capture of one kernel (`-k regex:...`).
"""

from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl8_b200 import AlgorithmConfig, _lib as L  # noqa: E402
from rl8_b200.env import CartPole  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2

lib = L.load()
torch.manual_seed(0)
algo = AlgorithmConfig(num_envs=N, horizon=T, enable_amp=prec == "bf16").build(CartPole)


def timed(fn) -> float:  # noqa: ANN001
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)


for i in range(steps):
    c = timed(algo.collect)
    s = timed(algo.step)
    print(f"step {i}: collect {c:.3f} ms  step {s:.3f} ms  -> {N * T / (c + s) / 1e3:.2f} M transitions/s")

algo.collect()
model = algo.policy.model
m, g = model.struct_for(model.flat_params), model.struct_for(algo._grads)
M = N * T
ws = algo._workspace("ppo", int(lib.rl8_ppo_workspace(m, M, algo.policy.precision)))
batch = algo._batch_struct()
ppo = L.PpoHparams(0.2, 0.0, 0.0, 5.0, 1.0, 1.0)
sums = torch.zeros(5, dtype=torch.float64, device=algo.device)


def minibatch() -> None:
    rc = lib.rl8_ppo_minibatch(m, g, batch, None, 0, M, float(M), ppo, L.ptr(sums), algo.policy.precision,
                               L.ptr(ws), ws.numel(), L.stream())
    assert rc == 0, rc


minibatch()
print(f"rl8_ppo_minibatch over {M} rows: {min(timed(minibatch) for _ in range(3)):.3f} ms")
if algo.policy.precision == L.PREC_FP32_TC:
    for name, bit in (("x3_update_f (forward + loss)", 1), ("x3_update_b (input gradient)", 2),
                      ("x3_update_w (weight gradient)", 4)):
        os.environ["RL8_X3_STAGES"] = str(bit)
        minibatch()
        print(f"   {name:32s} {min(timed(minibatch) for _ in range(3)):.3f} ms (incl. the max |obs| and W2 piece-image launches)")
    os.environ.pop("RL8_X3_STAGES")
    # optional: sweep of the policy / value split of the 74 CTA pairs per kernel (RL8_X3_SWEEP=1)
    if os.environ.get("RL8_X3_SWEEP"):
        for name, bit, var in (("f", 1, "RL8_X3_POLICY_PAIRS"), ("b", 2, "RL8_X3_POLICY_PAIRS_B"),
                               ("w", 4, "RL8_X3_POLICY_PAIRS_W")):
            os.environ["RL8_X3_STAGES"] = str(bit)
            res = []
            for n_pi in range(30, 52, 2):
                os.environ[var] = str(n_pi)
                minibatch()
                res.append(f"{n_pi}: {min(timed(minibatch) for _ in range(3)):.3f}")
            os.environ.pop(var)
            print(f"   sweep {name} (policy pairs of 74: ms)  " + "  ".join(res))
        os.environ.pop("RL8_X3_STAGES")
