"""rl8_b200: B200-native rollout-and-update engine behind rl8's public API."""

from .algorithms import Algorithm, AlgorithmConfig
from .env import Env
from .recurrent import RecurrentAlgorithm, RecurrentAlgorithmConfig, RecurrentTrainer
from .trainers import Trainer

__all__ = [
    "Algorithm",
    "AlgorithmConfig",
    "Env",
    "RecurrentAlgorithm",
    "RecurrentAlgorithmConfig",
    "RecurrentTrainer",
    "Trainer",
]
__version__ = "0.1.0"
