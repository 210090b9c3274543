"""Stand-in for the third-party `gymnasium` package (TEST INFRASTRUCTURE ONLY).

The reference (`/root/reference`) imports gymnasium but neither vendors nor pins it,
and it is not installed in this image.  This stand-in provides just the surface the
reference's CSTR hot path touches: `Env`, `Wrapper`, `spaces.*`,
`utils.seeding.np_random`, `envs.registration.EnvSpec`, `error`, `logger`.
None of the CSTR arithmetic lives in gymnasium; only `Box.sample()` numerics are
outside the reference tree (warm-up action stream: parity unpinned).
"""
from gymnasium.core import Env, Wrapper, ObservationWrapper, RewardWrapper, ActionWrapper  # noqa: F401
from gymnasium import spaces, error, logger, utils, envs  # noqa: F401
from gymnasium.spaces import Space  # noqa: F401

__version__ = "0.29.1"


class GoalEnv(Env):
    pass


def make(*args, **kwargs):
    raise error.Error("gymnasium stand-in: gym.make is not available")
