"""The HBM-bound kernels at the sizes of bench.py's `stream_rooflines`, for `ncu --set full`:

    python tools/profile_streams.py            # CUDA-event GB/s of each kernel (L2 flushed before every launch)
    ncu --set full -k regex:'gae_scan_hm|gae_normalize|env_step_kernel|view_windows' -s 15 -c 5 ... (same command)

Launch order per round: gae_scan_hm_kernel<4> (functional form), gae_scan_hm_kernel<4> (Algorithm.step form), gae_normalize_kernel, env_step_kernel<cartpole>, view_windows_kernel<u32>;
three warm-up rounds, then one measured round.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl8_b200 import _lib as L  # noqa: E402
from rl8_b200.env import CartPole  # noqa: E402

lib = L.load()
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = L.stream()
N, T = 1 << 20, 32
r = torch.randn(T + 1, N, device=dev)
v = torch.randn(T + 1, N, device=dev)
adv, ret = torch.empty_like(r), torch.empty_like(r)
mom = torch.zeros(3, dtype=torch.float64, device=dev)
sdev = torch.ones(2, device=dev)
Ne = 1 << 24
env = CartPole(Ne, 32, device=dev)
env.reset()
act = torch.randint(0, 3, (Ne,), device=dev)
cfg = env.rl8_cfg()
Nv, T1, D, size = 1 << 19, 33, 5, 4
x = torch.randn(T1, D, Nv, device=dev).permute(2, 0, 1)
o = torch.empty(Nv * T1, size, D, device=dev)
mk = torch.empty(Nv * T1, size, dtype=torch.bool, device=dev)

kernels = [
    ("gae_scan_hm_kernel<4>  N=2^20 T=32", 16.0 * N * T + 8.0 * N, 20.0 * N * T + 20.0 * N,
     lambda: lib.rl8_gae_scan(L.ptr(r), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, 1, N, 0.95, 0.95, 1.0, L.ptr(mom), st)),
    ("gae_scan_hm_kernel<4>  step form (no r write-back)", 16.0 * N * T + 8.0 * N, 16.0 * N * T + 16.0 * N,
     lambda: lib.rl8_gae_scan_dev(L.ptr(r), L.ptr(v), L.ptr(adv), L.ptr(ret), N, T, 1, N, 0.95, 0.95, L.ptr(sdev), 0,
                                  L.ptr(mom), st)),
    ("gae_normalize_kernel   N=2^20 T=32", 8.0 * N * T, 8.0 * N * T,
     lambda: lib.rl8_gae_normalize(L.ptr(adv), N, T, 1, N, L.ptr(mom), st)),
    ("env_step_kernel<cartpole> N=2^24", 64.0 * Ne, 64.0 * Ne,
     lambda: lib.rl8_env_step(env.rl8_kind, cfg, L.ptr(env.state), L.ptr(act), L.ptr(env._obs), 1, Ne,
                              L.ptr(env._reward), Ne, st)),
    ("view_windows_kernel<u32> N=2^19 T+1=33 D=5 size=4", 4.0 * Nv * T1 * D * (1 + size) + 1.0 * Nv * T1 * size,
     4.0 * Nv * T1 * D * (1 + size) + 1.0 * Nv * T1 * size,
     lambda: lib.rl8_view_windows(L.ptr(x), 4, Nv, T1, D, x.stride(0), x.stride(1), x.stride(2), size, -(size - 1), T1,
                                  L.ptr(o), L.ptr(mk), st)),
]
for rnd in range(4):
    for name, survey_bytes, kernel_bytes, fn in kernels:
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if rnd == 3:
            ms = e0.elapsed_time(e1)
            print(f"{name:52s} {ms * 1e3:8.1f} us  SURVEY §8d bytes {survey_bytes / 1e6:8.1f} MB -> {survey_bytes / ms / 1e6:7.0f} GB/s"
                  f"  | bytes the kernel moves {kernel_bytes / 1e6:8.1f} MB -> {kernel_bytes / ms / 1e6:7.0f} GB/s")
