from typing import Any, Optional, TypeVar

import numpy as np

from gymnasium.utils import seeding

ObsType = TypeVar("ObsType")
ActType = TypeVar("ActType")
WrapperObsType = TypeVar("WrapperObsType")
WrapperActType = TypeVar("WrapperActType")
RenderFrame = TypeVar("RenderFrame")


class Env:
    def __class_getitem__(cls, item):
        return cls

    metadata: dict = {"render_modes": []}
    render_mode: Optional[str] = None
    spec = None
    observation_space = None
    action_space = None
    _np_random: Optional[np.random.Generator] = None
    _np_random_seed: Optional[int] = None

    def step(self, action):
        raise NotImplementedError

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        if seed is not None:
            self._np_random, self._np_random_seed = seeding.np_random(seed)

    def render(self):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self) -> np.random.Generator:
        if self._np_random is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value: np.random.Generator):
        self._np_random = value
        self._np_random_seed = -1

    def has_wrapper_attr(self, name: str) -> bool:
        return hasattr(self, name)

    def get_wrapper_attr(self, name: str) -> Any:
        return getattr(self, name)

    def set_wrapper_attr(self, name: str, value: Any):
        setattr(self, name, value)

    def __enter__(self):
        return self

    def __exit__(self, *args):
        self.close()
        return False


class Wrapper(Env):
    def __init__(self, env: Env):
        self.env = env
        self._action_space = None
        self._observation_space = None
        self._metadata = None

    def get_wrapper_attr(self, name: str) -> Any:
        if name in self.__dict__ or hasattr(type(self), name):
            return getattr(self, name)
        return self.env.get_wrapper_attr(name)

    def has_wrapper_attr(self, name: str) -> bool:
        return name in self.__dict__ or hasattr(type(self), name) or self.env.has_wrapper_attr(name)

    @property
    def spec(self):
        return self.env.spec

    @property
    def action_space(self):
        return self.env.action_space if self._action_space is None else self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def observation_space(self):
        return self.env.observation_space if self._observation_space is None else self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    @property
    def metadata(self):
        return self.env.metadata if self._metadata is None else self._metadata

    @metadata.setter
    def metadata(self, value):
        self._metadata = value

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def np_random(self):
        return self.env.np_random

    @np_random.setter
    def np_random(self, value):
        self.env.np_random = value

    def step(self, action):
        return self.env.step(action)

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return obs, self.reward(reward), terminated, truncated, info

    def reward(self, reward):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError
