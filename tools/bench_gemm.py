"""Time the recurrent path's three tensor-core GEMM shapes (rl8_tc_gemm) at 65 536 rows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl8_b200 import _lib as L  # noqa: E402

lib = L.load()
R = 65536


def bench(name, a_k, b_k, acc, M, N, K, splits):  # noqa: ANN001, ANN201
    A = torch.randn((M, K) if a_k else (K, M), device="cuda")
    B = torch.randn((N, K) if b_k else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    f = lambda: lib.rl8_tc_gemm(a_k, b_k, acc, L.ptr(A), L.ptr(B), L.ptr(C), M, N, K, A.stride(0), B.stride(0),  # noqa: E731
                                C.stride(0), splits, L.stream())
    for _ in range(3):
        assert f() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms * 1e3:7.1f} us, {2.0 * M * N * K / ms / 1e9:6.1f} TFLOP/s")


bench("gates  = h W_hh^T   [65536 x 1024 x 256]", 1, 1, 0, R, 1024, 256, 1)
bench("dh     = dG W_hh    [65536 x 256 x 1024]", 1, 0, 0, R, 256, 1024, 1)
bench("gW_hh += dG^T h     [1024 x 256 x 65536]", 0, 0, 1, 1024, 256, R, 64)

# one whole LSTM step (gate GEMM + cell + both heads) through rl8_lstm_forward at 65 536 rows
import rl8_b200.env as E  # noqa: E402
from rl8_b200 import RecurrentAlgorithmConfig  # noqa: E402

for amp in (True, False):
    algo = RecurrentAlgorithmConfig(num_envs=256, horizon=8, seq_len=4, seqs_per_state_reset=2,
                                    enable_amp=amp).build(E.CartPole)
    pol = algo.policy
    obs = torch.randn(R, 5, device="cuda")
    h = torch.randn(R, 256, device="cuda") * 0.1
    c = torch.randn(R, 256, device="cuda") * 0.1
    for _ in range(3):
        pol.step_net(obs, h, c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pol.step_net(obs, h, c)
    e1.record()
    e1.synchronize()
    print(f"rl8_lstm_forward ({'bf16' if amp else 'fp32'}), 65536 rows (incl. 2 output allocations): "
          f"{e0.elapsed_time(e1) * 100:.1f} us")
