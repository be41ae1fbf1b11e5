"""Host-side breakdown of ``Trainer.step()`` (wall clock, synchronised): where the e2e number of
bench.py spends its time beyond the kernels."""
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

from rl8_b200 import AlgorithmConfig, Trainer  # noqa: E402
from rl8_b200.distributions import Categorical  # noqa: E402
from rl8_b200.env import CartPole  # noqa: E402

N, T = 65536, 32
host_noise = torch.empty(T, N, 3).exponential_(1).pin_memory()
dev_noise = torch.empty(T, N, 3, device="cuda")


class HostNoise(Categorical):
    @classmethod
    def draw_noise(cls, steps, num, width, device):  # noqa: ANN001, ANN206
        if steps != T:
            return super().draw_noise(steps, num, width, device)
        dev_noise.copy_(host_noise, non_blocking=True)
        return dev_noise


algo = AlgorithmConfig(num_envs=N, horizon=T, enable_amp=True, distribution_cls=HostNoise).build(CartPole)
tr = Trainer(algo)
for _ in range(3):
    tr.step()
torch.cuda.synchronize()


def clock(fn, n=10):  # noqa: ANN001, ANN201
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


print("Trainer.step          %.3f ms" % clock(tr.step))
print("memory_stats          %.3f ms" % clock(algo.memory_stats))
print("collect               %.3f ms" % clock(lambda: (algo.collect(), setattr(algo.state, "buffered", True))))
algo.collect()
t = []
for _ in range(10):
    algo.collect()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    algo.step()
    torch.cuda.synchronize()
    t.append(1e3 * (time.perf_counter() - t0))
print("step                  %.3f ms" % (sum(t) / len(t)))
print("H2D 25 MB pinned      %.3f ms" % clock(lambda: dev_noise.copy_(host_noise, non_blocking=True)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("flush.zero_ 256 MB    %.3f ms" % clock(flush.zero_))


def per_step(fn, n=12):  # noqa: ANN001, ANN201
    out = []
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        out.append(round(1e3 * (time.perf_counter() - t0), 2))
    return out


print("Trainer.step each            ", per_step(tr.step))
print("memory_stats each            ", per_step(algo.memory_stats))
real_mem = algo.memory_stats
algo.memory_stats = lambda: {}
print("Trainer.step w/o memory_stats", per_step(tr.step))
algo.memory_stats = real_mem
HostNoise.draw_noise = classmethod(lambda cls, steps, num, width, device: dev_noise)
print("Trainer.step w/o H2D copy    ", per_step(tr.step))
algo.memory_stats = lambda: {}
print("Trainer.step w/o both        ", per_step(tr.step))
import gc  # noqa: E402

gc.disable()
print("  ... and gc disabled        ", per_step(tr.step))
