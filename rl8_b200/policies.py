"""Policy = model + action distribution (src/rl8/policies/_feedforward.py:20-176)."""

from __future__ import annotations

from typing import Any, Literal, Mapping

import torch

from . import _lib
from .data import DataKeys, Device
from .distributions import Distribution
from .models import Model
from .specs import TensorSpec

ViewKind = Literal["last", "all"]


class Policy:
    """Union of a feedforward model and an action distribution.

    ``sample`` keeps the reference signature; the forward pass runs through
    ``rl8_mlp_forward`` and the distribution through ``rl8_dist_sample``.  Only the default
    models with shift-0 view requirements are on the fused path.
    """

    def __init__(
        self,
        observation_spec: TensorSpec,
        action_spec: TensorSpec,
        /,
        *,
        model: None | Model = None,
        model_cls: None | type[Model] = None,
        model_config: None | dict[str, Any] = None,
        distribution_cls: None | type[Distribution] = None,
        device: Device = "cpu",
    ) -> None:
        self.model_config = model_config or {}
        if model and model_cls:
            raise ValueError(
                "`model` and `model_cls` args are mutually exclusive."
                "Provide one or the other, but not both."
            )
        if torch.device(device).type != "cuda":
            raise RuntimeError("rl8_b200 policies run on CUDA only; there is no CPU path.")
        if model is None:
            model_cls = model_cls or Model.default_model_cls(observation_spec, action_spec)
            model = model_cls(observation_spec, action_spec, **self.model_config)
        if not isinstance(model, Model):
            raise NotImplementedError(
                "custom model classes are outside the fused hot path; use the default models"
            )
        self.model = model.flatten_(device)
        self.distribution_cls = distribution_cls or Distribution.default_dist_cls(action_spec)
        if not (isinstance(self.distribution_cls, type) and issubclass(self.distribution_cls, Distribution)):
            raise NotImplementedError("custom distributions must subclass rl8_b200 distributions")
        self.device = torch.device(device)
        #: GEMM precision of the kernels (``_lib.PREC_FP32`` | ``_lib.PREC_BF16``).
        self.precision = _lib.PREC_FP32
        self._lib = _lib.load()
        self._ws: None | torch.Tensor = None

    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward_net(self, which: int, obs: torch.Tensor, out: None | torch.Tensor = None) -> torch.Tensor:
        """``out [B, P]`` = policy (``which=0``) or value (``which=1``) network on ``obs [B, D]``
        (any strides)."""
        _lib.require_cuda(obs, "obs")
        if obs.dtype != torch.float32:
            obs = obs.float()
        B = obs.shape[0]
        width = 1 if which else self.model.head_width
        if out is None:
            out = torch.empty(B, width, device=obs.device)
        nbytes = int(self._lib.rl8_mlp_forward_workspace(self.model.hidden, B))
        ws = self._workspace(nbytes)
        m = self.model.struct_for(self.model.flat_params)
        rc = self._lib.rl8_mlp_forward(
            m, which, _lib.ptr(obs), obs.stride(0), obs.stride(1), B, _lib.ptr(out),
            int(which == 0 and self.model.head_width == 2 and self._continuous()),
            self.precision, _lib.ptr(ws), ws.numel(), _lib.stream(),
        )
        _lib.check(rc, "rl8_mlp_forward")
        return out

    def _continuous(self) -> bool:
        return self.distribution_cls.rl8_kind != _lib.DIST_CATEGORICAL

    def sample(
        self,
        batch: Mapping[str, torch.Tensor],
        /,
        *,
        kind: ViewKind = "last",
        deterministic: bool = False,
        inplace: bool = False,
        requires_grad: bool = False,
        return_actions: bool = True,
        return_logp: bool = False,
        return_values: bool = False,
        return_views: bool = False,
    ) -> dict[str, Any]:
        """Sample the policy on ``batch["obs"]`` of shape ``[B, T, D]``.

        ``kind="last"`` uses the most recent observation of every row (``[B, D]``),
        ``kind="all"`` flattens ``B`` and ``T`` (row ``b * T + t``), as the shift-0 view
        requirements of the default models do (src/rl8/views.py:408-445).
        """
        if requires_grad:
            raise NotImplementedError(
                "autograd through Policy.sample is not part of the fused path; gradients are"
                " produced by Algorithm.step's fused backward"
            )
        if DataKeys.VIEWS in batch:
            obs = batch[DataKeys.VIEWS][DataKeys.OBS]
        else:
            obs = batch[DataKeys.OBS]
            if obs.dim() == 3:
                obs = obs[:, -1] if kind == "last" else obs.flatten(0, 1)
        out: dict[str, Any] = dict(batch) if inplace else {}
        head = self.forward_net(0, obs)
        features = self.model.features_from_head(head)
        out[DataKeys.FEATURES] = features
        if return_actions:
            dist = self.distribution_cls(features, self.model)
            actions = dist.deterministic_sample() if deterministic else dist.sample()
            out[DataKeys.ACTIONS] = actions
            if return_logp:
                out[DataKeys.LOGP] = dist.logp(actions)
        if return_values:
            self.model._value = self.forward_net(1, obs)
            out[DataKeys.VALUES] = self.model._value
        if return_views:
            out[DataKeys.VIEWS] = {DataKeys.OBS: obs}
        return out
