from typing import Optional

import numpy as np

from gymnasium import error


def np_random(seed: Optional[int] = None):
    """gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed))), entropy."""
    if seed is not None and not (isinstance(seed, (int, np.integer)) and 0 <= seed):
        raise error.Error(f"Seed must be a non-negative integer, actual: {seed!r}")
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    return rng, np_seed
