"""ctypes binding of ``librl8_b200.so`` (``include/rl8_b200.h``).

The product has no CPU path and no PyTorch-eager stand-in: if the CUDA library is missing
or a call fails, this module raises.  Build with ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C rl8_b200/csrc``.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Any

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librl8_b200.so")

ABI_VERSION = 1

# enums (include/rl8_b200.h)
ENV_DISCRETE_DUMMY, ENV_CONTINUOUS_DUMMY, ENV_CARTPOLE, ENV_MOUNTAIN_CAR, ENV_PENDULUM = range(5)
DIST_CATEGORICAL, DIST_NORMAL, DIST_SQUASHED_NORMAL = range(3)
PREC_FP32, PREC_BF16, PREC_FP32_TC = range(3)



def precision_for(enable_amp: bool) -> int:
    """Kernel precision of an ``enable_amp`` setting: ``True`` -> bf16 tcgen05 GEMMs; ``False`` -> fp32
    results, on tcgen05 through split operands -- two fp16 pieces per fp32 value (``PREC_FP32_TC``), or -- with ``RL8_FP32_SIMT=1`` in
    the environment -- the CUDA-core fp32 GEMMs the split path is cross-checked against."""
    if enable_amp:
        return PREC_BF16
    return PREC_FP32 if os.environ.get("RL8_FP32_SIMT", "0") == "1" else PREC_FP32_TC


_ERRORS = {
    -1: "RL8_ERR_ARG (bad argument)",
    -2: "RL8_ERR_CUDA (launch failed)",
    -3: "RL8_ERR_UNSUPPORTED (outside the fused path)",
    -4: "RL8_ERR_WORKSPACE (workspace too small)",
}


class EnvCfg(C.Structure):
    _fields_ = [("p", C.c_float * 16)]


class Model(C.Structure):
    _fields_ = [("D", C.c_int32), ("H", C.c_int32), ("P", C.c_int32)] + [
        (f"{net}_{name}", C.c_void_p)
        for net in ("pi", "vf")
        for name in ("w1", "b1", "w2", "b2", "w3", "b3")
    ]


class PpoHparams(C.Structure):
    _fields_ = [
        ("clip_param", C.c_float),
        ("dual_clip_param", C.c_float),
        ("entropy_coeff", C.c_float),
        ("vf_clip_param", C.c_float),
        ("vf_coeff", C.c_float),
        ("loss_scale", C.c_float),
    ]


class Rollout(C.Structure):
    _fields_ = [
        ("env_kind", C.c_int32),
        ("dist_kind", C.c_int32),
        ("T", C.c_int32),
        ("deterministic", C.c_int32),
        ("N", C.c_int64),
        ("gamma", C.c_float),
        ("normalize_rewards", C.c_int32),
        ("env_cfg", EnvCfg),
        ("env_state", C.c_void_p),
        ("obs", C.c_void_p),
        ("actions", C.c_void_p),
        ("logp", C.c_void_p),
        ("values", C.c_void_p),
        ("rewards", C.c_void_p),
        ("rdr", C.c_void_p),
        ("noise", C.c_void_p),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("dist_kind", C.c_int32),
        ("T", C.c_int32),
        ("N", C.c_int64),
        ("obs", C.c_void_p),
        ("actions", C.c_void_p),
        ("logp", C.c_void_p),
        ("advantages", C.c_void_p),
        ("returns", C.c_void_p),
    ]


class LstmModel(C.Structure):
    _fields_ = [("D", C.c_int32), ("H", C.c_int32), ("P", C.c_int32)] + [
        (name, C.c_void_p)
        for name in ("w_ih", "w_hh", "b_ih", "b_hh", "pi_w", "pi_b", "vf_w", "vf_b")
    ]


class RecurrentRollout(C.Structure):
    _fields_ = [
        ("ro", Rollout),
        ("hidden", C.c_void_p),
        ("cell", C.c_void_p),
        ("seq_len", C.c_int32),
        ("seqs_per_state_reset", C.c_int32),
        ("seqs", C.c_int64),
    ]


class RecurrentBatch(C.Structure):
    _fields_ = [
        ("b", Batch),
        ("hidden", C.c_void_p),
        ("cell", C.c_void_p),
        ("seq_len", C.c_int32),
    ]


_i32, _i64, _f32, _f64, _vp, _int = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p, C.c_int

# name -> (restype, argtypes); mirrors include/rl8_b200.h one for one.
SIGNATURES: dict[str, tuple[Any, list[Any]]] = {
    "rl8_abi_version": (_int, []),
    "rl8_last_error": (C.c_char_p, []),
    "rl8_env_reset": (_int, [_int, C.POINTER(EnvCfg), _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "rl8_env_observe": (_int, [_int, _vp, _vp, _i64, _i64, _i64, _vp]),
    "rl8_env_step": (_int, [_int, C.POINTER(EnvCfg), _vp, _vp, _vp, _i64, _i64, _vp, _i64, _vp]),
    "rl8_dist_sample": (_int, [_int, _vp, _i32, _vp, _int, _vp, _vp, _i64, _vp]),
    "rl8_dist_logp_entropy": (_int, [_int, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "rl8_gae_scan": (
        _int,
        [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _f64, _f64, _f64, _vp, _vp],
    ),
    "rl8_gae_normalize": (_int, [_vp, _i64, _i32, _i64, _i64, _vp, _vp]),
    "rl8_reward_scale": (_int, [_vp, _f64, _int, _vp, _vp]),
    "rl8_gae_scan_dev": (
        _int,
        [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _f64, _f64, _vp, _int, _vp, _vp],
    ),
    "rl8_collect_stats": (_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "rl8_collect_stats_from": (_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "rl8_lstm_collect_workspace": (_i64, [C.POINTER(LstmModel), _i64, _i32, _int]),
    "rl8_lstm_collect": (
        _int, [C.POINTER(LstmModel), C.POINTER(RecurrentRollout), _int, _vp, _i64, _vp]
    ),
    "rl8_lstm_forward": (
        _int,
        [C.POINTER(LstmModel), _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _int,
         _vp, _i64, _vp],
    ),
    "rl8_lstm_ppo_workspace": (_i64, [C.POINTER(LstmModel), _i64, _i32, _int]),
    "rl8_lstm_ppo_minibatch": (
        _int,
        [C.POINTER(LstmModel), C.POINTER(LstmModel), C.POINTER(RecurrentBatch), _vp, _i64, _i64,
         _f64, C.POINTER(PpoHparams), _vp, _int, _vp, _i64, _vp],
    ),
    "rl8_mlp_forward_workspace": (_i64, [_i32, _i64]),
    "rl8_mlp_forward": (
        _int,
        [C.POINTER(Model), _int, _vp, _i64, _i64, _i64, _vp, _int, _int, _vp, _i64, _vp],
    ),
    "rl8_collect_workspace": (_i64, [C.POINTER(Model), _i64, _i32, _int]),
    "rl8_collect": (_int, [C.POINTER(Model), C.POINTER(Rollout), _int, _vp, _i64, _vp]),
    "rl8_ppo_workspace": (_i64, [C.POINTER(Model), _i64, _int]),
    "rl8_ppo_minibatch": (
        _int,
        [
            C.POINTER(Model),
            C.POINTER(Model),
            C.POINTER(Batch),
            _vp,
            _i64,
            _i64,
            _f64,
            C.POINTER(PpoHparams),
            _vp,
            _int,
            _vp,
            _i64,
            _vp,
        ],
    ),
    "rl8_ppo_losses": (
        _int,
        [_int, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f64, C.POINTER(PpoHparams), _vp, _vp,
         _vp, _vp],
    ),
    "rl8_ppo_losses_direct": (
        _int,
        [_int, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _f64, C.POINTER(PpoHparams), _vp, _vp,
         _vp, _vp],
    ),
    "rl8_view_windows": (_int, [_vp, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _i32, _i64, _i64, _vp, _vp, _vp]),
    "rl8_tc_selftest": (_int, [_vp, _vp, _vp, _i32, _i32, _int, _int, _vp]),
    "rl8_tc_selftest_tf32": (_int, [_vp, _vp, _vp, _i32, _vp]),
    "rl8_tc_selftest_tmem": (_int, [_vp, _vp, _vp]),
    "rl8_tc_selftest_tmem_16x256b": (_int, [_vp, _vp, _vp]),
    "rl8_tc3_selftest": (_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "rl8_tc_bench_tmem": (_int, [_vp, _i32, _i32, _i32, _vp]),
    "rl8_tc_phase_buffer": (_int, [_vp]),
    "rl8_tc_gemm": (_int, [_int, _int, _int, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _i64, _i64, _i32, _vp]),
    "rl8_tc_bench_mma": (_int, [_vp, _i32, _i32, _i32, _int, _int, _vp]),
    "rl8_tc3_bench_pace": (_int, [_vp, _i32, _i32, _i32, _i32, _vp]),
    "rl8_x3_debug_buffer": (_int, [_vp]),
    "rl8_clip_adam": (
        _int,
        [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _f64, _i64, _vp, _vp],
    ),
    "rl8_clip_adam_dev": (
        _int,
        [_vp, _vp, _vp, _vp, _i64, _f64, _vp, _f64, _f64, _f64, _vp, _vp, _vp],
    ),
    "rl8_clip_grads": (_int, [_vp, _i64, _f64, _vp, _vp]),
}

_lib: None | C.CDLL = None


def load() -> C.CDLL:
    """Load the shared library (once) and declare every prototype.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: rl8_b200 has no CPU or PyTorch path. Build it with"
            " `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.rl8_abi_version() != ABI_VERSION:
        raise RuntimeError(
            f"librl8_b200.so ABI {lib.rl8_abi_version()} != binding ABI {ABI_VERSION}; rebuild."
        )
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    detail = ""
    if rc == -2 and _lib is not None:
        detail = ": " + _lib.rl8_last_error().decode()
    if rc == -3:
        raise NotImplementedError(f"{what}: {_ERRORS[rc]}")
    raise RuntimeError(f"{what} failed with {_ERRORS.get(rc, rc)}{detail}")


def ptr(t: None | torch.Tensor) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"`{name}` lives on {t.device}; rl8_b200 runs on CUDA (sm_100a) only and has no CPU path."
        )
