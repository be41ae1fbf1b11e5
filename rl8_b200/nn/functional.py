"""GAE and PPO losses as functions (src/rl8/nn/functional.py:50-123, 259-363)."""

from __future__ import annotations

from typing import Any, MutableMapping

import torch

from .. import _lib
from ..data import DataKeys
from ..distributions import Distribution


def _col(t: torch.Tensor) -> torch.Tensor:
    """[B, T+1, 1] -> [B, T+1] view."""
    return t.squeeze(-1) if t.dim() == 3 else t


def generalized_advantage_estimate(
    batch: MutableMapping[str, torch.Tensor],
    /,
    *,
    gae_lambda: float = 0.95,
    gamma: float = 0.95,
    inplace: bool = False,
    normalize_advantages: bool = True,
    return_returns: bool = True,
    reward_scale: float = 1.0,
) -> MutableMapping[str, torch.Tensor]:
    """GAE over ``batch["rewards"]`` / ``batch["values"]`` of shape ``[B, T+1, 1]``.

    Semantics of the reference, including its side effect: ``batch["rewards"]`` is replaced
    by the scaled rewards.  Env-major inputs (the reference layout) take the warp-scan
    kernel, horizon-major views (``Algorithm.buffer``) the sequential coalesced one.
    """
    lib = _lib.load()
    rewards, values = batch[DataKeys.REWARDS], batch[DataKeys.VALUES]
    _lib.require_cuda(rewards, "rewards")
    out: MutableMapping[str, torch.Tensor] = batch if inplace else {}
    r2, v2 = _col(rewards), _col(values)
    B, T1 = r2.shape
    # Scaled rewards are a new tensor in the reference (batch["rewards"] = rewards / c).
    r_new = torch.empty_like(r2).copy_(r2) if not inplace else r2
    if DataKeys.ADVANTAGES not in out:
        out[DataKeys.ADVANTAGES] = torch.zeros_like(rewards)
    adv = _col(out[DataKeys.ADVANTAGES])
    ret = None
    if return_returns:
        out[DataKeys.RETURNS] = torch.empty_like(rewards)
        ret = _col(out[DataKeys.RETURNS])
    for t in (r_new, v2, adv) + ((ret,) if ret is not None else ()):
        if t.stride() != r_new.stride():
            raise ValueError("rewards / values / advantages / returns must share one layout")
    moments = torch.zeros(3, dtype=torch.float64, device=rewards.device)
    rc = lib.rl8_gae_scan(
        _lib.ptr(r_new), _lib.ptr(v2), _lib.ptr(adv), _lib.ptr(ret), B, T1 - 1,
        r_new.stride(0), r_new.stride(1), gamma, gae_lambda, reward_scale, _lib.ptr(moments),
        _lib.stream(),
    )
    _lib.check(rc, "rl8_gae_scan")
    if normalize_advantages:
        rc = lib.rl8_gae_normalize(
            _lib.ptr(adv), B, T1 - 1, adv.stride(0), adv.stride(1), _lib.ptr(moments), _lib.stream()
        )
        _lib.check(rc, "rl8_gae_normalize")
    batch[DataKeys.REWARDS] = r_new.reshape(rewards.shape) if not inplace else rewards
    return out


def ppo_losses(
    buffer_batch: MutableMapping[str, torch.Tensor],
    sample_batch: MutableMapping[str, torch.Tensor],
    sample_distribution: Distribution,
    /,
    *,
    clip_param: float = 0.2,
    dual_clip_param: None | float = 5.0,
    entropy_coeff: float = 0.0,
    vf_clip_param: float = 1.0,
    vf_coeff: float = 1.0,
    return_grads: bool = False,
) -> dict[str, Any]:
    """PPO loss components ``{"entropy", "policy", "vf", "total"}`` (0-dim tensors).

    ``losses["policy"]`` is the clipped surrogate *objective*; ``total = vf_coeff * vf -
    policy - entropy_coeff * entropy``.  With ``return_grads`` the hand-derived gradients of
    ``total`` w.r.t. the distribution features (``"d_features" [B, P]``) and the values
    (``"d_values" [B, 1]``) are returned too -- what autograd would give the reference.
    """
    lib = _lib.load()
    feats = sample_distribution._packed()
    B, P = feats.shape
    kind = sample_distribution.rl8_kind
    discrete = kind == _lib.DIST_CATEGORICAL
    if entropy_coeff != 0 and kind == _lib.DIST_SQUASHED_NORMAL:
        sample_distribution.entropy()  # raises NotImplementedError like the reference
    col = lambda k, dt=torch.float32: buffer_batch[k].reshape(-1).to(dt).contiguous()  # noqa: E731
    values = sample_batch[DataKeys.VALUES].reshape(-1).float().contiguous()
    actions = col(DataKeys.ACTIONS, torch.int64 if discrete else torch.float32)
    logp_old, adv, ret = col(DataKeys.LOGP), col(DataKeys.ADVANTAGES), col(DataKeys.RETURNS)
    hp = _lib.PpoHparams(clip_param, dual_clip_param or 0.0, entropy_coeff, vf_clip_param, vf_coeff, 1.0)
    sums = torch.zeros(5, dtype=torch.float64, device=feats.device)
    d_f = torch.empty_like(feats) if return_grads else None
    d_v = torch.empty(B, 1, device=feats.device) if return_grads else None
    rc = lib.rl8_ppo_losses(
        kind, _lib.ptr(feats), P, _lib.ptr(values), _lib.ptr(actions), _lib.ptr(logp_old),
        _lib.ptr(adv), _lib.ptr(ret), B, float(B), hp, _lib.ptr(sums), _lib.ptr(d_f), _lib.ptr(d_v),
        _lib.stream(),
    )
    _lib.check(rc, "rl8_ppo_losses")
    means = (sums[:4] / sums[4]).float()
    entropy, policy, vf = means[0], means[1], means[2]
    total = vf_coeff * vf - policy
    if entropy_coeff != 0:
        total = total - entropy_coeff * entropy
    out: dict[str, Any] = {"entropy": entropy, "policy": policy, "vf": vf, "total": total,
                           "kl_div": means[3]}
    if return_grads:
        out["d_features"], out["d_values"] = d_f, d_v
    return out
