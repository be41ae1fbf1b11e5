"""Soak run: many Trainer.step() iterations on the fp32 tensor-core path (the default: split-operand kernels, update
replayed from a CUDA graph) and on the bf16 path (feedforward and recurrent), checking that every statistic stays
finite and that returns improve -- a cheap guard against races that a single step would hide."""
import math
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

import rl8_b200.env as E  # noqa: E402
from rl8_b200 import AlgorithmConfig, RecurrentAlgorithmConfig, RecurrentTrainer, Trainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 500
torch.manual_seed(0)
for name, trainer, n in (
    ("CartPole ff fp32-tc N=65536", Trainer(AlgorithmConfig(num_envs=65536, horizon=32).build(E.CartPole)), max(20, steps // 4)),
    ("Pendulum squashed fp32-tc N=8192, 4 shuffled minibatches",
     Trainer(AlgorithmConfig(num_envs=8192, horizon=32, sgd_minibatch_size=65536).build(E.Pendulum)), max(20, steps // 4)),
    ("CartPole ff bf16 N=65536", Trainer(AlgorithmConfig(num_envs=65536, horizon=32, enable_amp=True).build(E.CartPole)), steps),
    ("CartPole ff bf16 N=1000 (ragged tiles), shuffled minibatches",
     Trainer(AlgorithmConfig(num_envs=1000, horizon=16, enable_amp=True, sgd_minibatch_size=3200).build(E.CartPole)), steps),
    ("Pendulum squashed bf16 N=8192", Trainer(AlgorithmConfig(num_envs=8192, horizon=32, enable_amp=True).build(E.Pendulum)), steps),
    ("CartPole LSTM bf16 N=4096", RecurrentTrainer(RecurrentAlgorithmConfig(num_envs=4096, horizon=32, enable_amp=True).build(E.CartPole)), max(10, steps // 10)),
):
    t0 = time.time()
    first = last = None
    for i in range(n):
        s = trainer.step()
        bad = {k: v for k, v in s.items() if isinstance(v, float) and not math.isfinite(v)}
        assert not bad, (name, i, bad)
        if i == 0:
            first = s["returns/mean"]
        last = s["returns/mean"]
    torch.cuda.synchronize()
    print(f"{name}: {n} steps in {time.time() - t0:.1f} s, returns/mean {first:.3f} -> {last:.3f}")
