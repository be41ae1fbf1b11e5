"""Default feedforward models (src/rl8/models/_feedforward.py:234-383) for the fused path.

Both default models are two independent MLPs, ``obs -> 256 -> ReLU -> 256 -> ReLU -> head``:
a policy network (``feature_model`` with an ``A``-way logits head, or ``latent_model`` with
``action_mean`` / ``action_log_std`` heads) and a value network ``vf_model``.  They are
ordinary ``nn.Module`` trees with the reference's parameter names (so ``state_dict()`` is
interchangeable) and the reference's initialisation order (``nn.Linear`` defaults, heads
U(+-1e-3) with zero bias, drawn from torch's global CPU generator).  Every parameter is a
view into ONE flat fp32 device buffer laid out for the kernels (``rl8_model`` in
include/rl8_b200.h), so the fused Adam and the NCCL gradient all-reduce touch a single
contiguous range.
"""

from __future__ import annotations

from typing import Any, Sequence

import torch
import torch.nn as nn

from . import _lib
from .data import DataKeys, Device
from .specs import Categorical, TensorSpec, Unbounded

_SEG_ALIGN = 4  # floats; keeps every segment 16-byte aligned for 128-bit loads


def _mlp(in_dim: int, hiddens: Sequence[int]) -> nn.Sequential:
    layers: list[nn.Module] = []
    for h in hiddens[:-1]:
        layers += [nn.Linear(in_dim, h), nn.ReLU()]
        in_dim = h
    layers.append(nn.Linear(in_dim, hiddens[-1]))
    return nn.Sequential(*layers)


def _small_head(in_dim: int, out_dim: int) -> nn.Linear:
    head = nn.Linear(in_dim, out_dim)
    nn.init.uniform_(head.weight, a=-1e-3, b=1e-3)
    nn.init.zeros_(head.bias)
    return head


class GenericModel(nn.Module):
    """User-defined feedforward model: the reference's ``Model`` plug-in point
    (src/rl8/models/_feedforward.py:19-231).

    Subclass it like the reference's ``Model``: build torch modules in ``__init__``, return the
    distribution's features from ``forward(batch)`` (``{"logits": [B, 1, A]}`` for
    ``Categorical``, ``{"mean": [B, 1], "log_std": [B, 1]}`` for ``Normal`` /
    ``SquashedNormal``) and the value estimate ``[B, 1]`` from ``value_function()``.
    ``view_requirements`` maps batch keys to :class:`~rl8_b200.views.ViewRequirement` (default: the
    most recent observation).  Such a model runs through torch (autograd included) on the GPU while
    everything around it -- env step, sampling, log-probabilities, GAE, the PPO losses and their
    gradients w.r.t. the model outputs -- stays on this library's kernels; any ``torch.optim``
    optimizer can drive it.  The default models below take the fully fused path instead.
    """

    def __init__(self, observation_spec: TensorSpec, action_spec: TensorSpec, /, **config: Any) -> None:
        super().__init__()
        from .views import ViewRequirement

        self.observation_spec = observation_spec
        self.action_spec = action_spec
        self.config = config
        self.view_requirements = {DataKeys.OBS: ViewRequirement(shift=0)}

    def apply_view_requirements(self, batch: Any, /, *, kind: str = "last") -> dict[str, Any]:
        """``{key: view}`` of ``batch [B, T, ...]``: ``kind="last"`` -> ``[B, ...]`` (sampling),
        ``kind="all"`` -> ``[B * T, ...]`` (training); src/rl8/models/_feedforward.py:58-101."""
        out = {}
        for key, vr in self.view_requirements.items():
            out[key] = vr.apply_all(key, batch) if kind == "all" else vr.apply_last(key, batch)
        return out

    @property
    def drop_size(self) -> int:
        return next(iter(vr.drop_size for vr in self.view_requirements.values()))

    def validate_view_requirements(self) -> None:
        drops = {k: vr.drop_size for k, vr in self.view_requirements.items()}
        if len(set(drops.values())) > 1:
            raise RuntimeError(
                f"{self} view requirements with drop sizes {drops} result in an ambiguous batch size."
            )

    def forward(self, batch: Any, /) -> dict[str, torch.Tensor]:  # pragma: no cover - abstract
        raise NotImplementedError

    def value_function(self) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError


class Model(nn.Module):
    """Base of the fused default models."""

    #: Number of policy-head outputs the kernels see (A logits, or 2 = mean + log_std).
    head_width: int

    def __init__(self, observation_spec: TensorSpec, action_spec: TensorSpec, /, **config: Any):
        super().__init__()
        self.observation_spec = observation_spec
        self.action_spec = action_spec
        self.config = config
        self._value: None | torch.Tensor = None
        self._flat: None | torch.Tensor = None
        self._segments: list[tuple[str, list[nn.Parameter]]] = []

    @staticmethod
    def default_model_cls(observation_spec: TensorSpec, action_spec: TensorSpec, /) -> type["Model"]:
        """src/rl8/models/_feedforward.py:103-133."""
        if not isinstance(observation_spec, Unbounded):
            raise TypeError(f"Observation spec {observation_spec} has no default model support.")
        if len(observation_spec.shape) != 1:
            raise TypeError("Default models support 1D observations only.")
        if isinstance(action_spec, Categorical):
            return DefaultDiscreteModel
        if isinstance(action_spec, Unbounded):
            return DefaultContinuousModel
        raise TypeError(f"Action spec {action_spec} has no default model support.")

    # -- flat parameter storage ----------------------------------------------------------------
    def _kernel_segments(self) -> list[tuple[str, list[nn.Parameter]]]:
        """(rl8_model field, parameters concatenated into it) in struct order."""
        raise NotImplementedError

    def flatten_(self, device: Device) -> "Model":
        """Move the parameters into one flat device buffer and re-point them at views."""
        segs = self._kernel_segments()
        sizes = [sum(p.numel() for p in ps) for _, ps in segs]
        padded = [(s + _SEG_ALIGN - 1) // _SEG_ALIGN * _SEG_ALIGN for s in sizes]
        flat = torch.zeros(sum(padded), device=device, dtype=torch.float32)
        self._offsets: dict[str, int] = {}
        off = 0
        with torch.no_grad():
            for (field, ps), pad in zip(segs, padded):
                self._offsets[field] = off
                o = off
                for p in ps:
                    view = flat[o : o + p.numel()].view(p.shape)
                    view.copy_(p.detach())
                    p.data = view
                    o += p.numel()
                off += pad
        self._flat = flat
        self._segments = segs
        return self

    @property
    def flat_params(self) -> torch.Tensor:
        assert self._flat is not None, "call flatten_(device) first"
        return self._flat

    def struct_for(self, flat: torch.Tensor) -> _lib.Model:
        """``rl8_model`` whose pointers address ``flat`` (the parameters, or a gradient /
        moment buffer with the same layout)."""
        m = _lib.Model()
        m.D = self.observation_spec.shape[0]
        m.H = self.hidden
        m.P = self.head_width
        base = flat.data_ptr()
        for field, off in self._offsets.items():
            setattr(m, field, base + 4 * off)
        return m

    def named_flat_views(self, flat: torch.Tensor) -> dict[str, torch.Tensor]:
        """Views of ``flat`` keyed like ``named_parameters()`` (e.g. to read gradients)."""
        names = {id(p): n for n, p in self.named_parameters()}
        out: dict[str, torch.Tensor] = {}
        for field, ps in self._segments:
            o = self._offsets[field]
            for p in ps:
                out[names[id(p)]] = flat[o : o + p.numel()].view(p.shape)
                o += p.numel()
        return out

    def load_state_dict(self, state_dict: Any, strict: bool = True, assign: bool = False) -> Any:
        """``nn.Module.load_state_dict`` that writes THROUGH the parameter views into the flat kernel buffer (the
        parameters are never re-pointed, so ``assign`` is refused); ``strict`` and the returned
        ``_IncompatibleKeys`` behave as in torch."""
        from torch.nn.modules.module import _IncompatibleKeys

        if assign:
            raise NotImplementedError("assign=True would detach the parameters from the flat kernel buffer")
        own = dict(self.named_parameters())
        missing = [k for k in own if k not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        errors = []
        for k, v in state_dict.items():
            if k in own and tuple(v.shape) != tuple(own[k].shape):
                errors.append(f"size mismatch for {k}: copying a param with shape {tuple(v.shape)} from checkpoint, "
                              f"the shape in current model is {tuple(own[k].shape)}.")
        if strict and (missing or unexpected):
            errors.insert(0, f"Missing key(s) in state_dict: {missing}. Unexpected key(s) in state_dict: {unexpected}.")
        if errors:
            raise RuntimeError(f"Error(s) in loading state_dict for {type(self).__name__}:\n\t" + "\n\t".join(errors))
        with torch.no_grad():
            for k, v in state_dict.items():
                if k in own:
                    own[k].copy_(v)
        return _IncompatibleKeys(missing, unexpected)

    def value_function(self) -> torch.Tensor:
        assert self._value is not None
        return self._value


def _check_config(hiddens: Sequence[int], activation_fn: str, bias: bool) -> int:
    if tuple(hiddens) != (256, 256) or activation_fn != "relu" or not bias:
        raise NotImplementedError(
            "the fused path implements the reference's default architecture only: hiddens=(256,"
            f" 256), activation_fn='relu', bias=True (got {tuple(hiddens)}, {activation_fn!r},"
            f" {bias})"
        )
    return int(hiddens[-1])


class DefaultDiscreteModel(Model):
    """1-D observations, one discrete action (src/rl8/models/_feedforward.py:313-383)."""

    def __init__(
        self,
        observation_spec: Unbounded,
        action_spec: Categorical,
        /,
        *,
        hiddens: Sequence[int] = (256, 256),
        activation_fn: str = "relu",
        bias: bool = True,
    ) -> None:
        super().__init__(observation_spec, action_spec)
        self.hidden = _check_config(hiddens, activation_fn, bias)
        d = observation_spec.shape[0]
        n_act = max(1, action_spec.shape[0] if len(action_spec.shape) else 1)
        if n_act != 1:
            raise NotImplementedError("the fused path supports a single discrete action")
        self.head_width = action_spec.space.n
        self.feature_model = nn.Sequential(_mlp(d, hiddens), nn.ReLU())
        self.feature_model.append(_small_head(hiddens[-1], self.head_width))
        self.vf_model = nn.Sequential(_mlp(d, hiddens), nn.ReLU(), nn.Linear(hiddens[-1], 1))

    def _kernel_segments(self) -> list[tuple[str, list[nn.Parameter]]]:
        f, v = self.feature_model, self.vf_model
        return [
            ("pi_w1", [f[0][0].weight]), ("pi_b1", [f[0][0].bias]),
            ("pi_w2", [f[0][2].weight]), ("pi_b2", [f[0][2].bias]),
            ("pi_w3", [f[2].weight]), ("pi_b3", [f[2].bias]),
            ("vf_w1", [v[0][0].weight]), ("vf_b1", [v[0][0].bias]),
            ("vf_w2", [v[0][2].weight]), ("vf_b2", [v[0][2].bias]),
            ("vf_w3", [v[2].weight]), ("vf_b3", [v[2].bias]),
        ]

    def features_from_head(self, head: torch.Tensor) -> dict[str, torch.Tensor]:
        return {"logits": head.reshape(-1, 1, self.head_width)}


class DefaultContinuousModel(Model):
    """1-D observations, one continuous action (src/rl8/models/_feedforward.py:234-310)."""

    def __init__(
        self,
        observation_spec: Unbounded,
        action_spec: Unbounded,
        /,
        *,
        hiddens: Sequence[int] = (256, 256),
        activation_fn: str = "relu",
        bias: bool = True,
    ) -> None:
        super().__init__(observation_spec, action_spec)
        self.hidden = _check_config(hiddens, activation_fn, bias)
        d = observation_spec.shape[0]
        if action_spec.shape[0] != 1:
            raise NotImplementedError("the fused path supports a single continuous action")
        self.head_width = 2
        self.latent_model = nn.Sequential(_mlp(d, hiddens), nn.ReLU())
        self.action_mean = _small_head(hiddens[-1], 1)
        self.action_log_std = _small_head(hiddens[-1], 1)
        self.vf_model = nn.Sequential(_mlp(d, hiddens), nn.ReLU(), nn.Linear(hiddens[-1], 1))

    def _kernel_segments(self) -> list[tuple[str, list[nn.Parameter]]]:
        f, v = self.latent_model, self.vf_model
        return [
            ("pi_w1", [f[0][0].weight]), ("pi_b1", [f[0][0].bias]),
            ("pi_w2", [f[0][2].weight]), ("pi_b2", [f[0][2].bias]),
            ("pi_w3", [self.action_mean.weight, self.action_log_std.weight]),
            ("pi_b3", [self.action_mean.bias, self.action_log_std.bias]),
            ("vf_w1", [v[0][0].weight]), ("vf_b1", [v[0][0].bias]),
            ("vf_w2", [v[0][2].weight]), ("vf_b2", [v[0][2].bias]),
            ("vf_w3", [v[2].weight]), ("vf_b3", [v[2].bias]),
        ]

    def features_from_head(self, head: torch.Tensor) -> dict[str, torch.Tensor]:
        return {"mean": head[:, 0:1].contiguous(), "log_std": head[:, 1:2].contiguous()}


__all__ = ["GenericModel", "Model", "DefaultDiscreteModel", "DefaultContinuousModel", "DataKeys"]
