"""One CartPole RecurrentAlgorithm ``collect()`` + ``step()`` at bench size (BASELINE config 3) for profilers.

    python tools/profile_lstm.py [fp32|bf16] [num_envs] [horizon] [steps] [sgd_iters]

Plain run: CUDA-event time of collect() and step().  Under ``ncu --metrics gpu__time_duration.sum`` the same command
gives the launch list of the recurrent path.
"""

from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl8_b200 import RecurrentAlgorithmConfig  # noqa: E402
from rl8_b200.env import CartPole  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 4

torch.manual_seed(0)
algo = RecurrentAlgorithmConfig(num_envs=N, horizon=T, enable_amp=prec == "bf16", num_sgd_iters=iters).build(CartPole)


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
    print(f"step {i}: collect {c:.3f} ms  step {s:.3f} ms ({iters} epochs)  -> {N * T / (c + s) / 1e3:.2f} M transitions/s")
