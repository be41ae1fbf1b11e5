"""Functional ops of the PPO hot path (src/rl8/nn/functional.py)."""

from .functional import generalized_advantage_estimate, ppo_losses

__all__ = ["generalized_advantage_estimate", "ppo_losses"]
