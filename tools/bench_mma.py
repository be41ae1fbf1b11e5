"""tcgen05 MMA pace and round-trip latency (one CTA): cycles per 128xNx16 instruction by N and operand
major, and the fixed cost of one issue -> commit -> mbarrier-wait round trip."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from rl8_b200 import _lib as L  # noqa: E402

lib = L.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")


def run(N, K, reps, a_mn, b_mn):  # noqa: ANN001, ANN201
    best = None
    for _ in range(3):
        rc = lib.rl8_tc_bench_mma(L.ptr(out), N, K, reps, a_mn, b_mn, L.stream())
        assert rc == 0, rc
        torch.cuda.synchronize()
        c = out.tolist()
        best = c if best is None or c[0] < best[0] else best
    return best


print("round trip of ONE instruction (K=16):")
for N in (16, 64, 128, 256):
    print(f"  N={N:3d}: {run(N, 16, 1, 0, 0)[0]} cycles")
print("pace: cycles per instruction over 64 GEMMs of K=256 (1024 instructions), K-major A and B:")
for N in (16, 32, 64, 128, 256):
    t, i = run(N, 256, 64, 0, 0)
    print(f"  N={N:3d}: {t / 1024:7.1f} cycles/instr (issue loop {i / 1024:5.1f}), {128 * N * 16 * 2 * 1024 / t:7.0f} FLOP/cycle")
print("same with B MN-major (the backward's view of W2):")
for N in (64, 128, 256):
    t, i = run(N, 256, 64, 0, 1)
    print(f"  N={N:3d}: {t / 1024:7.1f} cycles/instr, {128 * N * 16 * 2 * 1024 / t:7.0f} FLOP/cycle")
print("thin GEMMs: A MN-major (K=128 rows), N=16:")
t, i = run(16, 128, 128, 1, 0)
print(f"  N= 16: {t / 1024:7.1f} cycles/instr")
t, i = run(256, 128, 128, 1, 1)
print(f"  N=256 (gW2 shape, A and B MN-major): {t / 1024:7.1f} cycles/instr")
print("kind::tf32, K=8 (the layer-1 instruction): round trip of one instruction, and pace over 256 instructions:")
for N in (128, 256):
    print(f"  N={N:3d}: round trip {run(N, 8, 1, 0, 0)[0]} cycles, pace {run(N, 8, 256, 0, 0)[0] / 256:6.1f} cycles/instr")
print("gW2 shape (A and B MN-major, K=128 rows) by N:")
for N in (64, 128, 256):
    t, i = run(N, 128, 128, 1, 1)
    print(f"  N={N:3d}: {t / 1024:7.1f} cycles/instr, {128 * N * 16 * 2 * 1024 / t:7.0f} FLOP/cycle")
