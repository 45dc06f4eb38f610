"""bg_b200 -- B200-native batched backgammon engine: a drop-in for the env / legal-move / board /
feature-encoding / 2-ply path of Nick-qsv/MLP-PPO-2PLY-P3.  Python + PyTorch on the host,
hand-written sm_100a CUDA behind a C ABI (include/bg_b200.h).  No CPU fallback.

The directory is named after the reference (mlp-ppo-2ply-p3_b200); import it as `bg_b200`
(the repo-root shim bg_b200.py registers it under that name).
"""
from ._lib import BgError, LIB_PATH, lib  # noqa: F401
from .build import build  # noqa: F401
from .engine import (FEATURES, LD_BF16, MovegenWorkspace, as_board52, encode, from_board52, initial_board52,  # noqa: F401
                     legal_moves, to_board52)
from .sharding import shard_range  # noqa: F401
from .value_net import ValueNet  # noqa: F401
from .policy_net import PolicyValueNet  # noqa: F401
from .ppo import ManualUpdate, PPOConfig, PPOLearner, PPOTrainer, TensorCoreUpdate, evaluate_vs_random  # noqa: F401
from .twoply import TwoPlySearch, greedy_actions, segment_argmax  # noqa: F401
from .single_env import BackgammonEnv, BoardView, Player  # noqa: F401
from .vec_env import B200BackgammonVecEnv, HostStepBuffers, StepInfos, VectorizedBackgammonEnv  # noqa: F401

__version__ = "0.1.0"
