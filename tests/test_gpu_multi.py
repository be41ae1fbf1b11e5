"""GPU (>= 2 devices; skipped on a single-GPU box): the env-sharded data-parallel path against the single-process
reference -- W ranks x N/W envs reproduce the vectors the unmodified reference recorded with N envs
(tools/check_multi_gpu_equivalence.py), and replicas stay bit-identical (tools/check_multi_gpu.py)."""

from __future__ import annotations

import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script: str, *args: str, world: int = 2) -> str:
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", script), *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    return out.stdout


@pytest.mark.parametrize("mode", ["tensor_core", "cuda_core"])
@pytest.mark.parametrize("case", ["ff_cartpole", "ff_discrete_dummy", "ff_pendulum_squashed"])
def test_two_ranks_reproduce_the_single_process_reference(case: str, mode: str) -> None:
    out = _torchrun("check_multi_gpu_equivalence.py", case, mode)
    assert "EQUIVALENCE OK" in out, out[-2000:]


def test_replicas_stay_identical_on_two_ranks() -> None:
    out = _torchrun("check_multi_gpu.py")
    assert out.count("parameters identical on 2 ranks: True") == 3, out[-2000:]
