"""B200-native (sm_100a) two-series CSTR hot path behind the reference's env / VecEnv / ReplayBuffer
protocols.  See DESIGN.md and INTEGRATION.md at the repository root.

The directory name contains a hyphen, so import it with
``importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")`` or through the ``b200rl`` alias
module at the repository root.  Importing the package does not need a GPU; constructing any of its
classes does (the CUDA library has no CPU fallback).
"""
from . import _build, _lib, dist
from ._lib import CstrLibraryError
from .buffer import GpuReplayBuffer, ReplayBufferSamples, bind_replay_buffer_class
from .env import GpuCSTRVecEnv, LazyInfos, TwoSeriesCSTREnv, bind_env_class, bind_vec_env_class
from .normalize import GpuVecNormalize, bind_vec_normalize_class
from .update import FusedSACUpdate, FusedTD3Update, bind_sac_class, bind_td3_class
from .update_ext import FusedBCQUpdate, FusedMultiAgentUpdate, bind_bcq_class, bind_multiagent_class
from .rollout import ActorWeights, AgentActorWeights, EpisodeStats, FusedRollout, bind_offpolicy_rollout, fused_rollout_unsupported

__all__ = [
    "ActorWeights",
    "AgentActorWeights",
    "CstrLibraryError",
    "EpisodeStats",
    "FusedRollout",
    "GpuCSTRVecEnv",
    "GpuReplayBuffer",
    "GpuVecNormalize",
    "FusedTD3Update",
    "FusedBCQUpdate",
    "FusedMultiAgentUpdate",
    "bind_bcq_class",
    "bind_multiagent_class",
    "FusedSACUpdate",
    "bind_env_class",
    "bind_offpolicy_rollout",
    "bind_sac_class",
    "bind_td3_class",
    "bind_vec_normalize_class",
    "LazyInfos",
    "ReplayBufferSamples",
    "TwoSeriesCSTREnv",
    "bind_replay_buffer_class",
    "bind_vec_env_class",
    "build",
    "dist",
    "fused_rollout_unsupported",
]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc, sm_100a)."""
    return _build.build(force=force, verbose=verbose)
