"""pcbgpu -- B200-native batched RL-PCB placement environments.

Drop-in for the reset/step hot path of PBozmarov/RL-Environment-for-Component-
Placement (environment/dummy_env_*.py + the flatten wrappers): the dynamics run
in hand-written sm_100a CUDA kernels behind the C ABI of include/pcb_env.h;
this package is the thin Python host side (ctypes, DLPack -> torch tensors).
There is no CPU implementation: importing works anywhere, but creating an
environment needs libpcbgpu.so and a CUDA device and fails loudly otherwise.
"""
from .config import EnvSpec, make_spec, KINDS  # noqa: F401
from .instances import InstanceBank, generate_bank  # noqa: F401
from .vec_env import PcbVecEnv  # noqa: F401
from .gym_compat import GymCompatEnv, create_env, init_env, init_env_config  # noqa: F401
from .configs import CONFIGS, SHIPPED_CONFIGS  # noqa: F401
from . import spaces  # noqa: F401
from .wrappers import (  # noqa: F401
    FlatteningActionMaskObservationWrapper, FlatteningActionWrapper,
    FlatteningActionMaskObservationWrapperRect, FlatteningActionWrapperRect,
    FlatteningActionMaskObservationWrapperSquare, FlatteningActionWrapperSquare)
from .rollout import (RolloutCollector, allreduce_episode_stats, allreduce_stats,  # noqa: F401
                      shard_range)
from .rllib_adapter import VectorEnvAdapter  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["PcbVecEnv", "GymCompatEnv", "create_env", "init_env", "InstanceBank", "EnvSpec",
           "make_spec", "RolloutCollector", "allreduce_episode_stats"]
