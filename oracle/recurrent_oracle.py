"""CPU oracle for rl8's RECURRENT PPO path (LSTM policy).  TEST INFRASTRUCTURE ONLY.

Companion of ``oracle/ppo_oracle.py`` (same rules: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it) restating

* ``DefaultDiscreteRecurrentModel`` / ``DefaultContinuousRecurrentModel``
  (src/rl8/models/_recurrent.py:169-341): ONE ``nn.LSTM(D, 256, batch_first=True)`` whose
  latents feed both the policy head(s) and the value head,
* ``RecurrentAlgorithm.collect`` (src/rl8/algorithms/_recurrent.py:325-479) and
* ``RecurrentAlgorithm.step`` (481-652): GAE, then truncated back-propagation through time
  over ``seq_len`` chunks replayed from the stored chunk-start states.

The LSTM cell is written out gate by gate (torch's documented ``i, f, g, o`` packing,
``c' = sig(f)*c + sig(i)*tanh(g)``, ``h' = sig(o)*tanh(c')``) instead of calling
``nn.LSTM``: on CPU ``nn.LSTM`` dispatches to oneDNN, whose internal summation order
differs from plain ``addmm`` by about one ulp per step, so this restatement is pinned
against the unmodified reference by ``tests/golden/generate_golden_recurrent.py`` at
1e-5 relative / 2e-6 absolute on fp32 values (discrete actions and all counters
bit-exact) rather than bit-for-bit.  Parity status: **pinned** (``tests/golden/rec_*.npz``).

All citations are ``path:line`` relative to the upstream repository root.
"""

from __future__ import annotations

from typing import Any, Callable

import torch
import torch.nn.functional as F

from .ppo_oracle import Dist, OracleEnv, Params, _RunningMean, approx_kl, gae, ppo_losses

# --------------------------------------------------------------------------------------
# Default recurrent models — src/rl8/models/_recurrent.py:169-341
# --------------------------------------------------------------------------------------


def init_recurrent_params(
    obs_dim: int, action_kind: str, num_actions: int, hidden: int = 256
) -> Params:
    """Parameters with the reference's initialisation, drawn from torch's global CPU
    generator in the reference's construction order (``nn.LSTM`` first, then the small
    U(+-1e-3) heads with zero bias, then the value head with ``nn.Linear`` defaults;
    src/rl8/models/_recurrent.py:210-223 continuous, 297-309 discrete)."""
    p: Params = {}
    lstm = torch.nn.LSTM(obs_dim, hidden, num_layers=1, bias=True, batch_first=True)
    for k, v in lstm.state_dict().items():
        p[f"lstm.{k}"] = v.detach().clone()

    def small_head(out_f: int) -> tuple[torch.Tensor, torch.Tensor]:
        lin = torch.nn.Linear(hidden, out_f)
        torch.nn.init.uniform_(lin.weight, a=-1e-3, b=1e-3)
        torch.nn.init.zeros_(lin.bias)
        return lin.weight.detach().clone(), lin.bias.detach().clone()

    if action_kind == "discrete":
        p["feature_head.weight"], p["feature_head.bias"] = small_head(num_actions)
        vf = torch.nn.Linear(hidden, 1)
        p["vf_head.weight"], p["vf_head.bias"] = vf.weight.detach().clone(), vf.bias.detach().clone()
    else:
        p["action_mean.weight"], p["action_mean.bias"] = small_head(num_actions)
        p["action_log_std.weight"], p["action_log_std.bias"] = small_head(num_actions)
        vf = torch.nn.Linear(hidden, 1)
        p["vf_model.weight"], p["vf_model.bias"] = vf.weight.detach().clone(), vf.bias.detach().clone()
    return p


def lstm_cell(
    p: Params, x: torch.Tensor, h: torch.Tensor, c: torch.Tensor
) -> tuple[torch.Tensor, torch.Tensor]:
    """One LSTM step: ``x [B, D]``, ``h, c [B, H]`` -> ``(h', c')`` (torch gate order i,f,g,o)."""
    gates = F.linear(x, p["lstm.weight_ih_l0"], p["lstm.bias_ih_l0"]) + F.linear(
        h, p["lstm.weight_hh_l0"], p["lstm.bias_hh_l0"]
    )
    i, f, g, o = gates.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def recurrent_forward(
    p: Params, obs: torch.Tensor, h0: torch.Tensor, c0: torch.Tensor
) -> tuple[dict[str, torch.Tensor], torch.Tensor, torch.Tensor, torch.Tensor]:
    """``obs [B, L, D]``, ``h0, c0 [B, H]`` -> (features over ``B*L`` rows (row ``b*L + k``),
    values ``[B*L, 1]``, ``h_L``, ``c_L``)  (src/rl8/models/_recurrent.py:225-250, 311-334)."""
    B, L = obs.shape[:2]
    h, c = h0, c0
    lat = []
    for k in range(L):
        h, c = lstm_cell(p, obs[:, k], h, c)
        lat.append(h)
    latents = torch.stack(lat, dim=1).reshape(B * L, -1)
    if "feature_head.weight" in p:
        logits = F.linear(latents, p["feature_head.weight"], p["feature_head.bias"])
        feats = {"logits": logits.reshape(B * L, 1, -1)}
        values = F.linear(latents, p["vf_head.weight"], p["vf_head.bias"])
    else:
        mean = F.linear(latents, p["action_mean.weight"], p["action_mean.bias"])
        raw = F.linear(latents, p["action_log_std.weight"], p["action_log_std.bias"])
        feats = {"mean": mean, "log_std": torch.tanh(raw)}
        values = F.linear(latents, p["vf_model.weight"], p["vf_model.bias"])
    return feats, values, h, c


# --------------------------------------------------------------------------------------
# collect() — src/rl8/algorithms/_recurrent.py:325-479
# --------------------------------------------------------------------------------------


def new_recurrent_buffer(
    N: int, T: int, obs_dim: int, action_kind: str, hidden: int = 256
) -> dict[str, torch.Tensor]:
    """``buffer_spec.zero([N, T+1])`` with the LSTM states (:228-247)."""
    act_dtype = torch.int64 if action_kind == "discrete" else torch.float32
    z = lambda: torch.zeros(N, T + 1, 1)  # noqa: E731
    return {
        "obs": torch.zeros(N, T + 1, obs_dim),
        "hidden_states": torch.zeros(N, T + 1, 1, hidden),
        "cell_states": torch.zeros(N, T + 1, 1, hidden),
        "rewards": z(),
        "actions": torch.zeros(N, T + 1, 1, dtype=act_dtype),
        "logp": z(),
        "values": z(),
        "advantages": z(),
        "returns": z(),
        "reversed_discounted_returns": z(),
    }


def collect_recurrent(
    p: Params,
    env: OracleEnv,
    buf: dict[str, torch.Tensor],
    dist: Dist,
    noise: None | torch.Tensor,
    *,
    seqs: int,
    seq_len: int = 4,
    seqs_per_state_reset: int = 8,
    gamma: float = 0.95,
    reset: bool = True,
    reset_state: None | torch.Tensor = None,
    deterministic: bool = False,
    normalize_rewards: bool = True,
) -> tuple[dict[str, float], int]:
    """One rollout of ``T`` steps into ``buf`` (in place); returns (stats incl.
    ``"reward_scale"``, updated ``state.seqs``).

    ``buf[states][:, t]`` is the state the policy CONSUMES at step ``t``; it is zeroed when
    ``t % seq_len == 0 and seqs % seqs_per_state_reset == 0`` (:384-392), and
    ``seqs`` advances every ``seq_len`` steps (:430-431).  Statistics use
    ``rewards[:, 1:-1]`` (:449), unlike the feedforward algorithm."""
    T = buf["obs"].shape[1] - 1
    rdr = buf["reversed_discounted_returns"]
    hs, cs = buf["hidden_states"], buf["cell_states"]
    with torch.no_grad():
        if reset:
            buf["obs"][:, 0] = env.reset(reset_state)
            if normalize_rewards:
                rdr[:, 0] = 0.0
        else:
            buf["obs"][:, 0] = buf["obs"][:, -1]
            if normalize_rewards:
                rdr[:, 0] = rdr[:, -1]
        hs[:, 0] = hs[:, -1]
        cs[:, 0] = cs[:, -1]
        for t in range(T):
            if seqs and seqs_per_state_reset < 0:
                pass
            elif not (t % seq_len) and not (seqs % seqs_per_state_reset):
                hs[:, t] = 0.0
                cs[:, t] = 0.0
            feats, values, h, c = recurrent_forward(
                p, buf["obs"][:, t : t + 1], hs[:, t, 0], cs[:, t, 0]
            )
            d = dist.bind(feats)
            actions = d.mode() if deterministic else d.sample(noise[t])
            logp = d.logp(actions)
            obs, rewards = env.step(actions)
            if normalize_rewards:
                rdr[:, t + 1] = gamma * rdr[:, t] + rewards
            buf["actions"][:, t] = actions
            buf["logp"][:, t] = logp
            buf["values"][:, t] = values
            buf["rewards"][:, t] = rewards
            buf["obs"][:, t + 1] = obs
            hs[:, t + 1, 0] = h
            cs[:, t + 1, 0] = c
            if not ((t + 1) % seq_len):
                seqs += 1
        _, values, _, _ = recurrent_forward(p, buf["obs"][:, -1:], hs[:, -1, 0], cs[:, -1, 0])
        buf["values"][:, -1] = values

        rewards = buf["rewards"][:, 1:-1]
        returns = torch.sum(rewards, dim=1)
        returns_std, returns_mean = torch.std_mean(returns)
        rewards_std, rewards_mean = torch.std_mean(rewards)
        stats = {
            "returns/min": float(torch.min(returns)),
            "returns/max": float(torch.max(returns)),
            "returns/mean": float(returns_mean),
            "returns/std": float(returns_std),
            "rewards/min": float(torch.min(rewards)),
            "rewards/max": float(torch.max(rewards)),
            "rewards/mean": float(rewards_mean),
            "rewards/std": float(rewards_std),
            "reward_scale": float(torch.std(rdr[:, 1:])) if normalize_rewards else 1.0,
        }
    return stats, seqs


# --------------------------------------------------------------------------------------
# step() — src/rl8/algorithms/_recurrent.py:481-652
# --------------------------------------------------------------------------------------


def sequence_losses(
    p: Params,
    dist: Dist,
    mb: dict[str, torch.Tensor],
    *,
    clip_param: float,
    dual_clip_param: None | float,
    entropy_coeff: float,
    vf_clip_param: float,
    vf_coeff: float,
) -> tuple[dict[str, torch.Tensor], torch.Tensor]:
    """Losses of a minibatch of sequences: ``mb[k]`` is ``[M, L, ...]``; the LSTM is replayed
    from the stored chunk-start state ``states[:, 0]`` (:553-566)."""
    M, L = mb["obs"].shape[:2]
    feats, values, _, _ = recurrent_forward(
        p, mb["obs"], mb["hidden_states"][:, 0, 0], mb["cell_states"][:, 0, 0]
    )
    d = dist.bind(feats)
    actions = mb["actions"].reshape(M * L, -1)
    logp_old = mb["logp"].reshape(M * L, 1)
    logp_new = d.logp(actions)
    ent = d.entropy() if entropy_coeff != 0 else None
    losses = ppo_losses(
        logp_new,
        values,
        ent,
        logp_old,
        mb["advantages"].reshape(M * L, 1),
        mb["returns"].reshape(M * L, 1),
        clip_param=clip_param,
        dual_clip_param=dual_clip_param,
        entropy_coeff=entropy_coeff,
        vf_clip_param=vf_clip_param,
        vf_coeff=vf_coeff,
    )
    with torch.no_grad():
        kl = approx_kl(d.logp(actions), logp_old)
    return losses, kl


def step_recurrent(
    p: Params,
    buf: dict[str, torch.Tensor],
    dist: Dist,
    opt_state: dict[str, Any],
    *,
    reward_scale: float,
    seq_len: int = 4,
    gamma: float = 0.95,
    gae_lambda: float = 0.95,
    normalize_advantages: bool = True,
    sgd_minibatch_size: None | int = None,
    num_sgd_iters: int = 4,
    shuffle: bool = False,
    accumulate_grads: bool = False,
    clip_param: float = 0.2,
    dual_clip_param: None | float = None,
    entropy_coeff: float = 0.0,
    vf_clip_param: float = 5.0,
    vf_coeff: float = 1.0,
    target_kl_div: None | float = None,
    max_grad_norm: float = 5.0,
    lr: float = 1e-3,
    betas: tuple[float, float] = (0.9, 0.999),
    eps: float = 1e-8,
    perms: None | list[torch.Tensor] = None,
    grad_hook: None | Callable[[Params], None] = None,
) -> dict[str, float]:
    """GAE + PPO epochs over minibatches of ``seq_len`` sequences + clip + Adam, in place.

    The buffer ``[N, T]`` is reshaped to ``[N*T/seq_len, seq_len]`` (sequence
    ``s = n*(T/seq_len) + chunk``, :517-518) and ``sgd_minibatch_size`` counts SEQUENCES
    (default ``N * (T // seq_len)``, :293-297)."""
    N, Tp1 = buf["rewards"].shape[:2]
    T = Tp1 - 1
    L = seq_len
    r, adv, ret = gae(
        buf["rewards"],
        buf["values"],
        gamma=gamma,
        gae_lambda=gae_lambda,
        reward_scale=reward_scale,
        normalize_advantages=normalize_advantages,
    )
    buf["rewards"], buf["advantages"], buf["returns"] = r, adv, ret
    S = N * (T // L)
    keys_in = ("obs", "hidden_states", "cell_states", "actions", "logp", "advantages", "returns")
    seqs = {k: buf[k][:, :-1].reshape(S, L, *buf[k].shape[2:]) for k in keys_in}
    M = sgd_minibatch_size or S
    num_mb = S // M
    accum = num_mb if accumulate_grads else 1

    for v in p.values():
        v.requires_grad_(True)
    if "adam" not in opt_state:
        opt_state["adam"] = torch.optim.Adam(list(p.values()), lr=lr, betas=betas, eps=eps)
    adam = opt_state["adam"]
    for g in adam.param_groups:
        g["lr"] = lr

    keys = ("losses/entropy", "losses/policy", "losses/vf", "losses/total", "monitors/kl_div")
    sums = {k: 0.0 for k in keys}
    means = {k: _RunningMean() for k in keys}
    coeff_means = {"coefficients/entropy": _RunningMean(), "coefficients/vf": _RunningMean()}
    stop = False
    for epoch in range(num_sgd_iters):
        if perms is not None:
            order = perms[epoch]
        elif shuffle:
            order = torch.randperm(S)
        else:
            order = torch.arange(S)
        for i, idx in enumerate(torch.split(order, M)):
            step_now = (i + 1) % accum == 0
            mb = {k: v[idx] for k, v in seqs.items()}
            losses, kl = sequence_losses(
                p,
                dist,
                mb,
                clip_param=clip_param,
                dual_clip_param=dual_clip_param,
                entropy_coeff=entropy_coeff,
                vf_clip_param=vf_clip_param,
                vf_coeff=vf_coeff,
            )
            losses = {k: v / accum for k, v in losses.items()}
            kl_f = float(kl)
            sums["losses/entropy"] += float(losses["entropy"].detach())
            sums["losses/policy"] += float(losses["policy"].detach())
            sums["losses/vf"] += float(losses["vf"].detach())
            sums["losses/total"] += float(losses["total"].detach())
            sums["monitors/kl_div"] += kl_f / accum
            coeff_means["coefficients/entropy"].update(entropy_coeff)
            coeff_means["coefficients/vf"].update(vf_coeff)
            if step_now:
                for k in keys:
                    means[k].update(sums[k])
                    sums[k] = 0.0
            if target_kl_div is not None and kl_f > 1.5 * target_kl_div:
                stop = True
                break
            losses["total"].backward()
            if step_now:
                if grad_hook is not None:
                    grad_hook({k: v.grad.detach().clone() for k, v in p.items()})
                torch.nn.utils.clip_grad_norm_(list(p.values()), max_grad_norm)
                adam.step()
                adam.zero_grad()
        if stop:
            break
    for v in p.values():
        v.requires_grad_(False)

    # Fresh zeroed buffer keeping the final observation AND the final states (:636-646).
    keep = {k: buf[k][:, -1].clone() for k in ("obs", "hidden_states", "cell_states")}
    for k, v in buf.items():
        buf[k] = torch.zeros_like(v)
    for k, v in keep.items():
        buf[k][:, -1] = v
    out = {k: m.avg for k, m in means.items()}
    out.update({k: m.avg for k, m in coeff_means.items()})
    return out
