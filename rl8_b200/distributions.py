"""Action distributions (src/rl8/distributions.py:18-170) on CUDA kernels.

``Distribution(features, model)`` keeps the reference's constructor and its
``sample / deterministic_sample / logp / entropy`` methods, each backed by
``rl8_dist_sample`` / ``rl8_dist_logp_entropy``.  Inside :class:`rl8_b200.Algorithm` the
distribution is *fused* into the rollout and update kernels; the class then only selects
the kernel (``rl8_kind``) and supplies the sampling noise through :meth:`draw_noise` --
the hook tests use to inject identical draws into this engine and the CPU oracle.
"""

from __future__ import annotations

from typing import Any, ClassVar, Mapping

import torch

from . import _lib
from .specs import Categorical as CategoricalSpec
from .specs import TensorSpec, Unbounded


class Distribution:
    """Policy component that turns model features into actions."""

    #: Kernel selector (``rl8_dist_kind`` in include/rl8_b200.h).
    rl8_kind: ClassVar[int]

    def __init__(self, features: Mapping[str, torch.Tensor], model: Any, /) -> None:
        self.features = features
        self.model = model
        self._lib = _lib.load()

    @staticmethod
    def default_dist_cls(action_spec: TensorSpec, /) -> type["Distribution"]:
        """Categorical for discrete specs, Normal for continuous ones
        (src/rl8/distributions.py:54-73)."""
        if isinstance(action_spec, CategoricalSpec):
            return Categorical
        if isinstance(action_spec, Unbounded):
            return Normal
        raise TypeError(f"Action spec {action_spec} has no default distribution support.")

    # -- noise ---------------------------------------------------------------------------
    @classmethod
    def draw_noise(cls, steps: int, num: int, width: int, device: Any) -> torch.Tensor:
        """Noise consumed by ``steps`` sampling steps of ``num`` rows.

        Categorical: ``q ~ Exp(1)`` of shape ``[steps, num, width]`` (``multinomial`` ==
        ``argmax(probs / q)``); Normal / SquashedNormal: ``z ~ N(0, 1)`` of shape
        ``[steps, num]``.  Override to inject pre-drawn noise.
        """
        raise NotImplementedError

    # -- packed features [B, P] ------------------------------------------------------------
    def _packed(self) -> torch.Tensor:
        raise NotImplementedError

    def _width(self) -> int:
        return self._packed().shape[-1]

    def _sample(self, deterministic: bool) -> torch.Tensor:
        feats = self._packed()
        B, P = feats.shape
        discrete = self.rl8_kind == _lib.DIST_CATEGORICAL
        out = torch.empty(B, 1, device=feats.device, dtype=torch.int64 if discrete else torch.float32)
        noise = None
        if not deterministic:
            noise = self.draw_noise(1, B, P, feats.device).reshape(-1).contiguous()
        rc = self._lib.rl8_dist_sample(
            self.rl8_kind, _lib.ptr(feats), P, _lib.ptr(noise), int(deterministic), _lib.ptr(out),
            None, B, _lib.stream(),
        )
        _lib.check(rc, "rl8_dist_sample")
        return out

    def sample(self) -> torch.Tensor:
        return self._sample(False)

    def deterministic_sample(self) -> torch.Tensor:
        return self._sample(True)

    def _logp_entropy(self, samples: None | torch.Tensor, want_entropy: bool) -> torch.Tensor:
        feats = self._packed()
        B, P = feats.shape
        out = torch.empty(B, 1, device=feats.device)
        discrete = self.rl8_kind == _lib.DIST_CATEGORICAL
        act = None
        if samples is not None:
            act = samples.reshape(-1).to(torch.int64 if discrete else torch.float32).contiguous()
        rc = self._lib.rl8_dist_logp_entropy(
            self.rl8_kind, _lib.ptr(feats), P, _lib.ptr(act),
            None if want_entropy else _lib.ptr(out), _lib.ptr(out) if want_entropy else None,
            B, _lib.stream(),
        )
        _lib.check(rc, "rl8_dist_logp_entropy")
        return out

    def logp(self, samples: torch.Tensor) -> torch.Tensor:
        return self._logp_entropy(samples, False)

    def entropy(self) -> torch.Tensor:
        return self._logp_entropy(None, True)


class Categorical(Distribution):
    """Discrete actions from ``features["logits"]`` of shape ``[B, 1, A]``."""

    rl8_kind = _lib.DIST_CATEGORICAL

    @classmethod
    def draw_noise(cls, steps: int, num: int, width: int, device: Any) -> torch.Tensor:
        return torch.empty(steps, num, width, device=device).exponential_(1)

    def _packed(self) -> torch.Tensor:
        logits = self.features["logits"]
        return logits.reshape(logits.shape[0], -1).contiguous().float()


class Normal(Distribution):
    """Continuous actions from ``features["mean"]`` and ``features["log_std"]`` (``[B, 1]``)."""

    rl8_kind = _lib.DIST_NORMAL

    @classmethod
    def draw_noise(cls, steps: int, num: int, width: int, device: Any) -> torch.Tensor:
        return torch.randn(steps, num, device=device)

    def _packed(self) -> torch.Tensor:
        mean, log_std = self.features["mean"], self.features["log_std"]
        if mean.shape[-1] != 1:
            raise NotImplementedError("the fused distributions support one action dimension")
        return torch.cat((mean.reshape(-1, 1), log_std.reshape(-1, 1)), dim=1).contiguous().float()


class SquashedNormal(Normal):
    """Normal squashed into ``[-1, 1]`` by ``tanh``; has no entropy
    (src/rl8/distributions.py:147-170)."""

    rl8_kind = _lib.DIST_SQUASHED_NORMAL

    def entropy(self) -> torch.Tensor:
        raise NotImplementedError(
            f"Entropy isn't defined for {type(self).__name__}. Set the entropy coefficient to"
            " `0` to avoid this error during training."
        )
