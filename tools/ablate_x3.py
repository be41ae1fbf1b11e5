"""Development: per-kernel times of the RL8_PREC_FP32_TC update with parts of the worker code switched off
(RL8_X3_ABL bits; results are then wrong, only the timing is meaningful).

    python tools/ablate_x3.py [num_envs] [horizon]
"""

from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl8_b200 import AlgorithmConfig, _lib as L  # noqa: E402
from rl8_b200.env import CartPole  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lib = L.load()
torch.manual_seed(0)
algo = AlgorithmConfig(num_envs=N, horizon=T).build(CartPole)
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


def timed(fn) -> float:  # noqa: ANN001
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1)


CASES = [
    (1, 0, "f full"), (1, 1, "f: no H1 production"), (1, 2, "f: no pass 1 math"), (1, 4, "f: no gW3 pass"),
    (1, 6, "f: no epilogue math"), (1, 7, "f: MMAs + barriers only"),
    (1, 7 + 64, "f: ... and no W2 bulk copies"), (1, 7 + 128, "f: ... and no loss phase"),
    (1, 7 + 64 + 128, "f: ... neither"), (1, 64, "f: full but no W2 bulk copies"),
    (2, 0, "b full"), (2, 8, "b: no dZ2 production"), (2, 16, "b: no epilogue math"), (2, 24, "b: MMAs + barriers only"),
    (4, 0, "w full"), (4, 32, "w: no production"),
    (7, 0, "all three"),
]
extra = [a.split("=") for a in sys.argv[3:]]
for k, v in extra:
    os.environ[k] = v
minibatch()
for stages, abl, name in CASES:
    os.environ["RL8_X3_STAGES"], os.environ["RL8_X3_ABL"] = str(stages), str(abl)
    minibatch()
    print(f"{name:32s} {min(timed(minibatch) for _ in range(3)):.3f} ms", flush=True)

# where the MMA warp of pair 0 waits (x3_update_f_kernel)
for abl in (0, 7 + 64 + 128):
    os.environ["RL8_X3_STAGES"], os.environ["RL8_X3_ABL"] = "1", str(abl)
    counters = torch.zeros(64, dtype=torch.int64, device=algo.device)
    lib.rl8_x3_debug_buffer(L.ptr(counters))
    minibatch()
    torch.cuda.synchronize()
    lib.rl8_x3_debug_buffer(None)
    c = counters.tolist()
    for net, label in ((0, "policy"), (1, "value")):
        d = c[16 * net: 16 * net + 16]
        tiles = max(d[10], 1)
        print(f"abl {abl} {label}: {d[9] / tiles:8.0f} cycles / tile ({tiles} tiles); MMA warp waits per tile: full[kc] "
              + " ".join(f"{x / tiles:6.0f}" for x in d[:8]) + f"  acc_empty {d[8] / tiles:6.0f}")
        for off, who in ((0, "warp 5 (no loss rows)"), (32, "warp 1 (loss rows)")):
            w = c[16 * net + off + 11: 16 * net + off + 16]
            print(f"      {who}: cycles / tile in produce {w[0] / tiles:6.0f}  pass 1 (+ accumulator wait) {w[1] / tiles:6.0f}"
                  f"  barrier 1 {w[2] / tiles:6.0f}  loss + barrier 2 {w[3] / tiles:6.0f}  pass 2 {w[4] / tiles:6.0f}")
