"""Accuracy of the split-bf16 pair GEMM (rl8_tc3_selftest) by term set and contraction length, against fp64
and beside a plain fp32 matmul of the same operands.  Run on a B200:  python tools/check_split_accuracy.py"""

from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl8_b200 import _lib  # noqa: E402

TERM_SETS = {"a0b0 (bf16)": 1, "x2: 3 terms": 7, "4 terms": 15, "x3: 6 terms": 63}


def errors(got: torch.Tensor, ref: torch.Tensor) -> tuple[float, float]:
    """(max |err| / max |ref|,  rms err / rms ref)"""
    e = (got.double() - ref).abs()
    return float(e.max() / ref.abs().max()), float(e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())


def main() -> None:
    lib = _lib.load()
    torch.manual_seed(0)
    dev = "cuda"
    for kind in ("randn", "relu x uniform weights (all products of one sign half the time)"):
        for K in (32, 256, 1024):
            if kind == "randn":
                A = torch.randn(256, K, device=dev)
                B = torch.randn(256, K, device=dev)
            else:
                A = torch.randn(256, K, device=dev).relu()
                B = (torch.rand(256, K, device=dev) - 0.3) / 16
            ref = A.double() @ B.double().T
            torch.backends.cuda.matmul.allow_tf32 = False
            print(f"{kind}  K={K}: fp32 matmul (cuBLAS)  max/max %.3e  rms/rms %.3e" % errors(A @ B.T, ref))
            for name, terms in TERM_SETS.items():
                D = torch.empty(256, 256, device=dev)
                rc = lib.rl8_tc3_selftest(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), K, terms, _lib.stream())
                _lib.check(rc, "rl8_tc3_selftest")
                torch.cuda.synchronize()
                print(f"    {name:14s} max/max %.3e  rms/rms %.3e" % errors(D, ref))


if __name__ == "__main__":
    main()
