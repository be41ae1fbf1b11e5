"""TMEM read throughput microbenchmark (one CTA): bytes per SM cycle by warp count and load shape."""
import torch

from rl8_b200 import _lib as L

lib = L.load()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
iters = 2000
for mode, per_round in ((0, 32 * 32 * 4), (1, 2 * 32 * 32 * 4), (2, 32 * 16 * 4)):
    for nw in (1, 4, 8, 16):
        lib.rl8_tc_bench_tmem(L.ptr(out), nw, iters, mode, L.stream())
        torch.cuda.synchronize()
        cyc = int(out[0])
        print(f"mode {mode} warps {nw:2d}: {cyc / iters:8.1f} cycles/round, "
              f"{nw * per_round * iters / cyc:8.1f} B/cycle/SM")
