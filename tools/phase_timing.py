"""Where a tile's time goes inside ``tc_update_h_kernel``: SM cycles per phase and tile.

    python tools/phase_timing.py [num_envs] [horizon]

Runs ``rl8_ppo_minibatch`` (bf16 tensor-core path) over a freshly collected CartPole buffer with
the ``rl8_tc_phase_buffer`` debug hook armed and prints the average cycles CTA 0 of each network
spends per tile in each phase (see the kernel's A..J comments).
"""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))

from rl8_b200 import AlgorithmConfig, _lib as L  # noqa: E402
from rl8_b200.env import CartPole  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lib = L.load()
algo = AlgorithmConfig(num_envs=N, horizon=T, enable_amp=True).build(CartPole)
algo.collect()
model = algo.policy.model
m, g = model.struct_for(model.flat_params), model.struct_for(algo._grads)
M = N * T
ws = algo._workspace("ppo", int(lib.rl8_ppo_workspace(m, M, algo.policy.precision)))
batch = L.Batch()
batch.dist_kind, batch.T, batch.N = 0, T, N
for k in ("obs", "actions", "logp", "advantages", "returns"):
    setattr(batch, k, algo.buffer.hm[k].data_ptr())
ppo = L.PpoHparams(0.2, 0.0, 0.0, 5.0, 1.0, 1.0)
sums = torch.zeros(5, dtype=torch.float64, device=algo.device)
counters = torch.zeros(24, dtype=torch.int64, device=algo.device)


def run() -> None:
    rc = lib.rl8_ppo_minibatch(m, g, batch, None, 0, M, float(M), ppo, L.ptr(sums), algo.policy.precision,
                               L.ptr(ws), ws.numel(), L.stream())
    assert rc == 0, rc


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.rl8_tc_phase_buffer(L.ptr(counters))
reps = 5
for _ in range(reps):
    run()
torch.cuda.synchronize()
lib.rl8_tc_phase_buffer(None)
c = counters.cpu().tolist()
names = ["A+B stage, Z1 MMA", "C   H1 epilogue", "D   Z2 MMA + H2 epi", "E   row loss", "F   gW3 + G MMA",
         "G   dZ2 epilogue", "H+I dH1 MMA + dZ1 epi", "J   gW1 thin MMA"]
chunk = min(M, 1 << 21)
ntiles = -(-chunk // 128)
n_pi = int(__import__('os').environ.get('RL8_H_POLICY_CTAS', 80))
for net, label, nct in ((0, "policy", n_pi), (1, "value", 148 - n_pi)):
    tiles = reps * (-(-M // chunk)) * (-(-ntiles // nct))  # tiles CTA 0 of this network processed
    tot = sum(c[8 * net: 8 * net + 8])
    print(f"{label}: {tot / tiles:8.0f} cycles / tile  ({tiles} tiles)")
    for i, nm in enumerate(names):
        v = c[8 * net + i]
        print(f"   {nm:24s} {v / tiles:8.0f}  {100 * v / tot:5.1f}%")

# weight-gradient kernel, CTA 0 (policy network, unit half 0): 37 CTAs per role
tiles = reps * (-(-M // chunk)) * (-(-ntiles // 37))
print(f"tc_update_w CTA 0 ({tiles} tiles):")
for i, nm in ((16, "wait Z1 / free stage"), (17, "H1 epilogue + barrier"), (18, "issuer: wait dZ2 bulk copy"),
              (19, "issuer: wait previous gW2"), (20, "issuer: issue Z1 + gW2")):
    print(f"   {nm:28s} {c[i] / tiles:8.0f}")
