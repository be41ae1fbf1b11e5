"""Pace of tcgen05.mma.cta_group::2 (M=256, K=16, bf16) issued back to back, on 1 pair and on all 74 pairs at once:
cycles and nanoseconds per instruction (their ratio is the SM clock the tensor pipe actually ran at)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl8_b200 import _lib as L  # noqa: E402

lib = L.load()
out = torch.zeros(2 * 74, dtype=torch.int64, device="cuda")
for pairs in (1, 74):
    for n_cols in (256, 128):
        for terms in (6, 3, 1):
            for reps in (8, 4096, 32768):
                best = None
                for _ in range(3):
                    out.zero_()
                    rc = lib.rl8_tc3_bench_pace(L.ptr(out), pairs, reps, terms, n_cols, L.stream())
                    assert rc == 0, rc
                    torch.cuda.synchronize()
                    o = out[: 2 * pairs].view(pairs, 2).double()
                    cyc, ns = float(o[:, 0].max()), float(o[:, 1].max())
                    best = (cyc, ns) if best is None or ns < best[1] else best
                n_instr = reps * 2 * terms
                cyc, ns = best
                flop = 2.0 * 256 * n_cols * 16 * n_instr * pairs
                print(f"pairs {pairs:2d} N {n_cols} terms {terms} reps {reps:5d}: {cyc / n_instr:7.1f} cycles/instr "
                      f"{ns / n_instr:7.1f} ns/instr  clock {cyc / ns * 1e3:6.0f} MHz  {flop / ns / 1e3:7.1f} TFLOP/s",
                      flush=True)

# both operands MN-major (the weight-gradient kernel's tiles): terms + 16
for terms in (6, 3, 2):
    best = None
    for _ in range(3):
        out.zero_()
        rc = lib.rl8_tc3_bench_pace(L.ptr(out), 74, 4096, terms + 16, 256, L.stream())
        assert rc == 0, rc
        torch.cuda.synchronize()
        o = out[: 2 * 74].view(74, 2).double()
        cyc, ns = float(o[:, 0].max()), float(o[:, 1].max())
        best = (cyc, ns) if best is None or ns < best[1] else best
    n_instr = 4096 * 2 * terms
    cyc, ns = best
    print(f"MN-major operands: pairs 74 N 256 terms {terms} reps 4096: {cyc / n_instr:7.1f} cycles/instr "
          f"{ns / n_instr:7.1f} ns/instr  clock {cyc / ns * 1e3:6.0f} MHz", flush=True)
