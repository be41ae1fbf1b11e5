"""CPU: the arithmetic behind the fp16-piece form of RL8_PREC_FP32_TC (rl8_b200/csrc/split_tc.cuh), emulated in numpy.

An fp32 operand scaled by a power of two into fp16's range is h0 + h1 up to 2^-22 of its size (h0 = f16(x),
h1 = f16(x - h0), the subtraction exact in fp32), and a0 b0 + a0 b1 + a1 b0 reproduces the dot product to ~1e-7 of its
scale at K = 256: coarser than the six products of three bf16 pieces (~1e-8), finer than the fp32-accumulated dot product of
the reference's own nn.Linear (~5e-7) -- the numbers DESIGN.md section 2 quotes; the GPU twin of this test is
tests/test_gpu_split.py::test_pair_mma_fp16_pieces_against_fp64.
"""

from __future__ import annotations

import numpy as np


def _pow2_scale_for(bound: float) -> float:
    """split_tc.cuh: the largest power of two s with bound * s <= 2^14 (capped at 2^60); 1 for 0 / inf / nan."""
    if not (bound > 0.0) or not (bound < 1.0e38):
        return 1.0
    _, e = np.frexp(np.float32(bound))
    return float(np.ldexp(np.float32(1.0), min(14 - int(e), 60)))


def _f16_pieces(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    h0 = x.astype(np.float16)
    r = x - h0.astype(np.float32)  # exact: the residual of a rounding
    assert np.array_equal(r.astype(np.float64), x.astype(np.float64) - h0.astype(np.float64))
    return h0, r.astype(np.float16)


def _bf16(x: np.ndarray) -> np.ndarray:  # round to nearest even on the upper 16 bits
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return u.astype(np.uint32).view(np.float32)


def test_scale_is_a_power_of_two_that_fits_the_bound() -> None:
    for bound in (1e-30, 3e-9, 0.0625, 1.0, 4.7, 65504.0, 1e6, 3e37):
        s = _pow2_scale_for(bound)
        m, _ = np.frexp(s)
        assert m == 0.5 and np.isfinite(s) and s > 0  # a power of two
        assert bound * s <= 2.0**14
        assert bound * s >= 2.0**13 or s == 2.0**60  # the largest such power, unless capped
        assert np.float32(1.0) / np.float32(s) > 0  # its inverse is representable
    for bound in (0.0, float("inf"), float("nan"), -1.0):
        assert _pow2_scale_for(bound) == 1.0


def test_two_fp16_pieces_and_three_products_match_six_bf16_products() -> None:
    rng = np.random.default_rng(0)
    K = 256
    a = rng.standard_normal((64, K)).astype(np.float32)
    b = (rng.standard_normal((48, K)) * 0.06).astype(np.float32)
    a[:, ::5] *= np.float32(2.0**-11)  # wide range inside one operand
    b[::3] *= np.float32(2.0**-9)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    sa, sb = _pow2_scale_for(float(np.abs(a).max())), _pow2_scale_for(float(np.abs(b).max()))
    xa, xb = a * np.float32(sa), b * np.float32(sb)
    assert np.array_equal(xa.astype(np.float64), a.astype(np.float64) * sa)  # scaling by a power of two is exact
    a0, a1 = _f16_pieces(xa)
    b0, b1 = _f16_pieces(xb)
    # representation: two pieces carry 22 bits (elements far below the bound: an absolute error of an fp16 subnormal)
    rep = np.abs(a0.astype(np.float64) + a1.astype(np.float64) - xa.astype(np.float64))
    assert np.all(rep <= np.maximum(2.0**-22 * np.abs(xa), 2.0**-25))
    A0, A1, B0, B1 = (p.astype(np.float64) for p in (a0, a1, b0, b1))
    got3 = (A0 @ B0.T + A0 @ B1.T + A1 @ B0.T) / (sa * sb)
    err3 = float(np.abs(got3 - ref).max() / np.abs(ref).max())
    # three bf16 pieces, six products, no scales (the first form of the path)
    p0 = _bf16(a)
    p1 = _bf16(a - p0)
    p2 = _bf16(a - p0 - p1)
    q0 = _bf16(b)
    q1 = _bf16(b - q0)
    q2 = _bf16(b - q0 - q1)
    P0, P1, P2, Q0, Q1, Q2 = (p.astype(np.float64) for p in (p0, p1, p2, q0, q1, q2))
    got6 = P0 @ Q0.T + P0 @ Q1.T + P1 @ Q0.T + P1 @ Q1.T + P0 @ Q2.T + P2 @ Q0.T
    err6 = float(np.abs(got6 - ref).max() / np.abs(ref).max())
    # ... two bf16 pieces, three products (fails the golden vectors) ... and the reference's own arithmetic: fp32 products
    # accumulated in fp32
    got_x2 = P0 @ Q0.T + P0 @ Q1.T + P1 @ Q0.T
    err_x2 = float(np.abs(got_x2 - ref).max() / np.abs(ref).max())
    acc = np.zeros_like(ref, dtype=np.float32)
    for k in range(K):
        acc = (acc + a[:, k : k + 1] * b[:, k][None, :]).astype(np.float32)
    err_fp32 = float(np.abs(acc - ref).max() / np.abs(ref).max())
    assert err6 < 2e-8 and err3 < 2e-7 and err3 < err_fp32, (err6, err3, err_fp32)
    assert err_x2 > 10 * err3, (err_x2, err3)


def test_relu_mask_from_the_sign_of_the_negated_sum() -> None:
    """h1_chunk<true>: with -[W1 | b1] staged (and no entry -0) the mask bit of z > 0 is the sign bit of the sum."""
    rng = np.random.default_rng(1)
    w = rng.standard_normal((256, 5)).astype(np.float32)
    bias = rng.standard_normal(256).astype(np.float32)
    bias[::7] = 0.0
    obs = rng.standard_normal((100, 5)).astype(np.float32)
    obs[::9] = 0.0  # all-zero rows: z = b exactly, z = +0 where the bias is zero
    nw, nb = (np.float32(-1.0) * w) + np.float32(0.0), (np.float32(-1.0) * bias) + np.float32(0.0)
    assert not np.signbit(nb[bias == 0]).any()  # the staging never leaves a -0
    z = np.tile(bias, (100, 1))
    nz = np.tile(nb, (100, 1))
    for d in range(5):  # bias first, then d ascending, one rounding per term (the kernels' FMA chain up to fusion)
        z = (z.astype(np.float64) + obs[:, d : d + 1].astype(np.float64) * w[:, d].astype(np.float64)).astype(np.float32)
        nz = (nz.astype(np.float64) + obs[:, d : d + 1].astype(np.float64) * nw[:, d].astype(np.float64)).astype(np.float32)
    assert np.array_equal(np.signbit(nz), z > 0)
    assert np.array_equal(np.minimum(nz, 0), -np.maximum(z, 0))
