"""Horizon-major rollout buffer with env-major views.

The reference's buffer is a TensorDict of batch size ``[N, T+1]`` whose fields are
env-major ``[N, T+1, F]`` tensors with a row pitch of ``T+1`` elements
(src/rl8/algorithms/_feedforward.py:239-256).  Here every field lives in ONE device
allocation in horizon-major order -- ``[T+1][N]`` floats, observations ``[T+1][D][N]`` --
so a rollout step writes one contiguous slab with coalesced 128-bit stores and GAE / the
update read unit-stride columns.  ``buffer["obs"]`` etc. return ``[N, T+1, F]`` strided
*views* of that memory, so reference-style indexing (``buffer["obs"][:, 0] = obs``) works
unchanged.  Fields start 16-byte aligned; slabs stay aligned (and the kernels use their
128-bit paths) when ``num_envs`` is a multiple of 4.
"""

from __future__ import annotations

from typing import Any, Iterator

import torch

from .data import DataKeys
from .specs import Categorical, Composite


class RolloutBuffer:
    """Dict-like rollout storage (``obs, rewards, actions, logp, values, advantages, returns``
    and, when rewards are normalised, ``reversed_discounted_returns``)."""

    def __init__(self, spec: Composite, num_envs: int, horizon: int, device: torch.device | str):
        self.spec = spec
        self.num_envs, self.horizon = num_envs, horizon
        self.device = torch.device(device)
        N, T1 = num_envs, horizon + 1
        self.obs_dim = int(spec[DataKeys.OBS].shape[0])
        self.discrete = isinstance(spec[DataKeys.ACTIONS], Categorical)
        # element counts (in 4-byte words) of each horizon-major field
        words = {
            DataKeys.OBS: T1 * self.obs_dim * N,
            DataKeys.ACTIONS: T1 * N * (2 if self.discrete else 1),
        }
        self.state_dim = 0
        for k in spec.keys():
            if k == DataKeys.STATES:
                # LSTM states of the recurrent algorithm: [T+1][N][H] per state tensor, so the
                # slab of one step is the row-major [N, H] operand of that step's GEMM.
                for sk, sspec in spec[k].items():
                    self.state_dim = int(sspec.shape[-1])
                    words[sk] = T1 * N * self.state_dim
            elif k not in words:
                words[k] = T1 * N
        # every field starts on a 256-byte boundary (the kernels use 128- and 256-bit accesses on field bases)
        self._raw = torch.zeros(sum((w + 63) // 64 * 64 for w in words.values()), device=device)
        self.hm: dict[str, torch.Tensor] = {}
        off = 0
        for k, w in words.items():
            seg = self._raw[off : off + w]
            if k == DataKeys.OBS:
                self.hm[k] = seg.view(T1, self.obs_dim, N)
            elif k == DataKeys.ACTIONS and self.discrete:
                self.hm[k] = seg.view(torch.int64).view(T1, N)
            elif k in (DataKeys.HIDDEN_STATES, DataKeys.CELL_STATES):
                self.hm[k] = seg.view(T1, N, self.state_dim)
            else:
                self.hm[k] = seg.view(T1, N)
            off += (w + 63) // 64 * 64
        self._views: dict[str, Any] = {
            k: self._env_major(k)
            for k in self.hm
            if k not in (DataKeys.HIDDEN_STATES, DataKeys.CELL_STATES)
        }
        if self.state_dim:
            # buffer["states"]["hidden_states"]: [N, T+1, num_layers=1, H] like the reference
            self._views[DataKeys.STATES] = {
                k: self.hm[k].permute(1, 0, 2).unsqueeze(2)
                for k in (DataKeys.HIDDEN_STATES, DataKeys.CELL_STATES)
            }
        self.batch_size = torch.Size([N, T1])

    def _env_major(self, key: str) -> torch.Tensor:
        t = self.hm[key]
        if key == DataKeys.OBS:
            return t.permute(2, 0, 1)  # [N, T+1, D]
        return t.permute(1, 0).unsqueeze(-1)  # [N, T+1, 1]

    # -- mapping protocol (env-major views) -------------------------------------------------
    def __getitem__(self, key: str) -> Any:
        return self._views[key]

    def __contains__(self, key: str) -> bool:
        return key in self._views

    def __iter__(self) -> Iterator[str]:
        return iter(self._views)

    def keys(self):  # noqa: ANN201
        return self._views.keys()

    def items(self):  # noqa: ANN201
        return self._views.items()

    def zero_(self) -> None:
        self._raw.zero_()

    def nbytes(self) -> int:
        return self._raw.numel() * 4
