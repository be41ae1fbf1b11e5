"""Policy = model + action distribution (src/rl8/policies/_feedforward.py:20-176)."""

from __future__ import annotations

from typing import Any, Literal, Mapping

import torch

from . import _lib
from .data import DataKeys, Device
from .distributions import Distribution
from .models import GenericModel, Model
from .specs import TensorSpec

ViewKind = Literal["last", "all"]



class PolicyExport:
    """``save`` / ``load`` and the pickled form shared by :class:`Policy` and ``RecurrentPolicy``."""

    # -- export (src/rl8/policies/_feedforward.py:178-190) -----------------------------------------
    def __getstate__(self) -> dict[str, Any]:
        """Pickled form: everything but the library handle and scratch memory; a default (flat-parameter) model travels
        as its class, specs, config and a CPU ``state_dict`` -- the reference's parameter names and shapes, so the same
        tensors load into the reference's ``Policy`` -- and is re-flattened on the loading side."""
        state = {k: v for k, v in self.__dict__.items() if k not in ("_lib", "_ws", "model")}
        if getattr(self, "fused", True):
            m = self.model
            state["_model_blob"] = (type(m), m.observation_spec, m.action_spec, dict(m.config),
                                    {k: v.detach().cpu() for k, v in m.state_dict().items()})
        else:
            state["model"] = self.model
        return state

    def __setstate__(self, state: dict[str, Any]) -> None:
        blob = state.pop("_model_blob", None)
        self.__dict__.update(state)
        if not torch.cuda.is_available():
            raise RuntimeError("rl8_b200 policies run on CUDA only; there is no CPU path.")
        if blob is not None:
            cls, obs_spec, act_spec, config, sd = blob
            self.model = cls(obs_spec, act_spec, **config).flatten_(self.device)
            self.model.load_state_dict(sd)
        self._lib = _lib.load()
        self._ws = None

    def save(self, path: Any, /) -> Any:
        """Cloud-pickle the policy to ``path`` (the reference's export convention, which it then wraps in an MLflow
        model -- MLflow is outside this engine's scope); ``Policy.load(path)`` restores it."""
        import cloudpickle

        with open(path, "wb") as f:
            cloudpickle.dump(self, f)
        return self

    @staticmethod
    def load(path: Any, /) -> Any:
        import pickle

        with open(path, "rb") as f:
            policy = pickle.load(f)
        if not isinstance(policy, PolicyExport):
            raise TypeError(f"{path!r} does not hold an rl8_b200 policy")
        return policy


class Policy(PolicyExport):
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
        if isinstance(model, Model):
            self.model = model.flatten_(device)
        elif isinstance(model, GenericModel):
            # a user-defined torch model: forward / backward through torch on the GPU, everything
            # around it on this library's kernels (see GenericModel)
            model.validate_view_requirements()
            self.model = model.to(device)
        else:
            raise TypeError(
                "`model` / `model_cls` must derive from rl8_b200.models.GenericModel (user-defined"
                " torch models) or be one of the default models"
            )
        #: ``True``: default model on the fully fused kernels; ``False``: user-defined torch model.
        self.fused = isinstance(model, Model)
        self.distribution_cls = distribution_cls or Distribution.default_dist_cls(action_spec)
        if not (isinstance(self.distribution_cls, type) and issubclass(self.distribution_cls, Distribution)):
            raise NotImplementedError("custom distributions must subclass rl8_b200 distributions")
        self.device = torch.device(device)
        #: GEMM precision of the kernels (``_lib.PREC_FP32_TC`` | ``_lib.PREC_BF16`` | ``_lib.PREC_FP32``).
        self.precision = _lib.precision_for(False)
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
        if not self.fused:
            return self._sample_generic(
                batch, kind=kind, deterministic=deterministic, inplace=inplace, requires_grad=requires_grad,
                return_actions=return_actions, return_logp=return_logp, return_values=return_values,
                return_views=return_views,
            )
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

    def _sample_generic(
        self, batch: Mapping[str, Any], /, *, kind: ViewKind, deterministic: bool, inplace: bool,
        requires_grad: bool, return_actions: bool, return_logp: bool, return_values: bool,
        return_views: bool,
    ) -> dict[str, Any]:
        """``sample`` for a user-defined torch model (src/rl8/policies/_feedforward.py:66-176): views ->
        ``model(views)`` (with autograd when ``requires_grad``) -> this library's sampling / log-prob
        kernels on the returned features."""
        if DataKeys.VIEWS in batch:
            views = batch[DataKeys.VIEWS]
        else:
            views = self.model.apply_view_requirements(batch, kind=kind)
        self.model.train(requires_grad)
        with torch.set_grad_enabled(requires_grad):
            features = self.model(views)
            out: dict[str, Any] = dict(batch) if inplace else {}
            out[DataKeys.FEATURES] = features
            if return_values:
                out[DataKeys.VALUES] = self.model.value_function()
        if return_actions:
            dist = self.distribution_cls({k: v.detach() for k, v in features.items()}, self.model)
            actions = dist.deterministic_sample() if deterministic else dist.sample()
            out[DataKeys.ACTIONS] = actions
            if return_logp:
                out[DataKeys.LOGP] = dist.logp(actions)
        if return_views:
            out[DataKeys.VIEWS] = views
        return out
