"""``VecNormalize`` on the device for the CSTR path.

Mirrors the reference wrapper (``core/common/vec_env/vec_normalize.py:19-340``): same constructor arguments and
attributes (``training, norm_obs, norm_reward, clip_obs, clip_reward, gamma, epsilon, obs_rms, ret_rms, returns``),
same ``step_wait`` order (update obs statistics -> normalise obs -> update return statistics -> normalise reward ->
zero the returns of done rows), ``normalize_obs/normalize_reward/unnormalize_*``, ``get_original_obs/reward``.
The running statistics live in one 16-double device block updated by ``cstr_norm_update``; ``GpuReplayBuffer.sample(env=this)``
applies them inside the gather kernel, so nothing goes back to the host.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from . import _lib
from .env import GpuCSTRVecEnv, LazyInfos


class _RmsView:
    """Read view of one ``RunningMeanStd`` (running_mean_std.py:4-22) inside the device statistics block."""

    def __init__(self, owner: "GpuVecNormalize", which: str):
        self._o, self._w = owner, which

    def _host(self):
        return self._o.stats.cpu().numpy()

    @property
    def mean(self):
        s = self._host()
        return s[0:4].copy() if self._w == "obs" else np.float64(s[9])

    @property
    def var(self):
        s = self._host()
        return s[4:8].copy() if self._w == "obs" else np.float64(s[10])

    @property
    def count(self) -> float:
        s = self._host()
        return float(s[8] if self._w == "obs" else s[11])


class GpuVecNormalize:
    """Device ``VecNormalize`` around a :class:`GpuCSTRVecEnv` (drop-in where the reference wraps its VecEnv)."""

    def __init__(self, venv: GpuCSTRVecEnv, training: bool = True, norm_obs: bool = True, norm_reward: bool = True, clip_obs: float = 10.0,
                 clip_reward: float = 10.0, gamma: float = 0.99, epsilon: float = 1e-8, norm_obs_keys: Optional[list] = None):
        if venv.dtype != "fp32":
            raise ValueError("GpuVecNormalize wraps the fp32 environment")
        torch = _lib.require_cuda()
        self._torch = torch
        self._libc = _lib.load()
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.render_mode = venv.render_mode
        self.device = venv.device
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = float(clip_obs), float(clip_reward), float(gamma), float(epsilon)
        self.norm_obs_keys = norm_obs_keys
        init = np.zeros(16)
        init[4:8], init[8], init[10], init[11] = 1.0, 1e-4, 1.0, 1e-4  # RunningMeanStd(): mean 0, var 1, count epsilon=1e-4
        self.stats = torch.as_tensor(init, dtype=torch.float64, device=self.device)
        self._scratch = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.returns_device = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        self.obs_rms, self.ret_rms = _RmsView(self, "obs"), _RmsView(self, "ret")
        self._norm_obs_dev = torch.empty((self.num_envs, 4), dtype=torch.float32, device=self.device)
        self._norm_rew_dev = torch.empty(self.num_envs, dtype=torch.float32, device=self.device)
        self.old_obs: Optional[np.ndarray] = None
        self.old_reward: Optional[np.ndarray] = None
        self.launches = 0

    # ---- plumbing --------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self.device).cuda_stream

    @property
    def returns(self) -> np.ndarray:
        return self.returns_device.cpu().numpy()

    def norm_params(self) -> "_lib.NormParams":
        """``cstr_norm_params`` for the fused gather (``GpuReplayBuffer.sample(env=self)``)."""
        return _lib.NormParams(stats=self.stats.data_ptr(), epsilon=self.epsilon, clip_obs=self.clip_obs, clip_reward=self.clip_reward,
                               norm_obs=int(self.norm_obs), norm_reward=int(self.norm_reward))

    def _update(self, obs_dev, rew_dev, done_dev) -> None:
        rc = self._libc.cstr_norm_update(self.num_envs, _lib.ptr(obs_dev), _lib.ptr(rew_dev), _lib.ptr(done_dev), _lib.ptr(self.returns_device),
                                         self.gamma, _lib.ptr(self.stats), _lib.ptr(self._scratch), self._stream())
        _lib.check(rc, "cstr_norm_update")
        self.launches += 2

    def _apply(self, obs_dev, rew_dev, obs_out, rew_out) -> None:
        n = obs_dev.shape[0] if obs_dev is not None else rew_dev.shape[0]
        rc = self._libc.cstr_norm_apply(n, _lib.ptr(obs_dev), _lib.ptr(rew_dev), _lib.ptr(self.stats), self.epsilon, self.clip_obs, self.clip_reward,
                                        _lib.ptr(obs_out), _lib.ptr(rew_out), self._stream())
        _lib.check(rc, "cstr_norm_apply")
        self.launches += 1

    # ---- VecEnv protocol (vec_normalize.py:174-259, 261-298) ---------------------------------------------------------
    def step_async(self, actions) -> None:
        self.venv.step_async(actions)

    def step_wait(self):
        torch = self._torch
        obs, rewards, dones, infos = self.venv.step_wait()  # raw host copies (old_obs / old_reward)
        self.old_obs, self.old_reward = obs, rewards
        v = self.venv
        with torch.cuda.device(self.device):
            if self.training and self.norm_obs:
                self._update(v.state, None, None)  # :188-193 obs_rms.update(obs) — obs after auto-reset, as in the reference
            if self.norm_obs:
                self._apply(v.state, None, self._norm_obs_dev, None)
                obs = self._norm_obs_dev.cpu().numpy()
            if self.training:
                self._update(None, v._reward, v._done)  # :197-198 + :218
            elif dones.any():
                self.returns_device.masked_fill_(v._done.bool(), 0.0)
            if self.norm_reward:
                self._apply(None, v._reward, None, self._norm_rew_dev)
                rewards = self._norm_rew_dev.cpu().numpy()
            else:
                rewards = rewards.astype(np.float32)
        if self.norm_obs and isinstance(infos, LazyInfos):  # :201-206 terminal observations are normalised too
            for _, info in infos.done_items():
                if "terminal_observation" in info:
                    info["terminal_observation"] = self.normalize_obs(info["terminal_observation"])
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def reset(self):
        obs = self.venv.reset()
        self.old_obs = obs
        self.returns_device.zero_()
        with self._torch.cuda.device(self.device):
            if self.training and self.norm_obs:
                self._update(self.venv.state, None, None)  # :296-297
        return self.normalize_obs(obs)

    # ---- value maps on arbitrary arrays -------------------------------------------------------------------------------
    def normalize_obs(self, obs):
        if not self.norm_obs:
            return np.array(obs, copy=True)
        torch = self._torch
        a = np.asarray(obs, np.float32)
        flat = torch.as_tensor(np.ascontiguousarray(a.reshape(-1, 4)), device=self.device)
        out = torch.empty_like(flat)
        with torch.cuda.device(self.device):
            self._apply(flat, None, out, None)
        return out.cpu().numpy().reshape(a.shape)

    def normalize_reward(self, reward):
        a = np.asarray(reward, np.float32)
        if not self.norm_reward:
            return a.astype(np.float32)
        torch = self._torch
        flat = torch.as_tensor(np.ascontiguousarray(a.reshape(-1)), device=self.device)
        out = torch.empty_like(flat)
        with torch.cuda.device(self.device):
            self._apply(None, flat, None, out)
        return out.cpu().numpy().reshape(a.shape)

    def unnormalize_obs(self, obs):
        if not self.norm_obs:
            return np.array(obs, copy=True)
        return (np.asarray(obs) * np.sqrt(self.obs_rms.var + self.epsilon)) + self.obs_rms.mean  # :212-219 (host: rarely used)

    def unnormalize_reward(self, reward):
        if not self.norm_reward:
            return reward
        return np.asarray(reward) * np.sqrt(self.ret_rms.var + self.epsilon)

    def get_original_obs(self):
        return np.array(self.old_obs, copy=True)

    def get_original_reward(self):
        return self.old_reward.copy()

    # ---- persistence / interchange (vec_normalize.py:128-172,310-332) ---------------------------------------------------
    def state_dict(self) -> dict:
        s = self.stats.cpu().numpy()
        return {"obs_mean": s[0:4].copy(), "obs_var": s[4:8].copy(), "obs_count": float(s[8]), "ret_mean": float(s[9]), "ret_var": float(s[10]),
                "ret_count": float(s[11]), "training": self.training, "norm_obs": self.norm_obs, "norm_reward": self.norm_reward,
                "clip_obs": self.clip_obs, "clip_reward": self.clip_reward, "gamma": self.gamma, "epsilon": self.epsilon}

    def load_state_dict(self, d: dict) -> None:
        s = np.zeros(16)
        s[0:4], s[4:8], s[8], s[9], s[10], s[11] = d["obs_mean"], d["obs_var"], d["obs_count"], d["ret_mean"], d["ret_var"], d["ret_count"]
        self.stats.copy_(self._torch.as_tensor(s))
        for k in ("training", "norm_obs", "norm_reward", "clip_obs", "clip_reward", "gamma", "epsilon"):
            if k in d:
                setattr(self, k, d[k])

    def __getstate__(self) -> dict:  # like the reference: the wrapped env is not pickled, set_venv() after loading
        return self.state_dict()

    def __reduce__(self):  # classes made by bind_vec_normalize_class are dynamic: always pickle as the plain class
        return (_unpickle, (self.state_dict(),))

    def __setstate__(self, state: dict) -> None:
        self.__dict__["_pending_state"] = state
        self.__dict__["venv"] = None

    def set_venv(self, venv: GpuCSTRVecEnv) -> None:
        if self.__dict__.get("venv") is not None:
            raise ValueError("Trying to set venv of already initialized VecNormalize wrapper.")
        state = self.__dict__.pop("_pending_state")
        GpuVecNormalize.__init__(self, venv)
        self.load_state_dict(state)

    def save(self, save_path: str) -> None:
        import pickle

        with open(save_path, "wb") as fh:
            pickle.dump(self, fh)

    @staticmethod
    def load(load_path: str, venv: GpuCSTRVecEnv) -> "GpuVecNormalize":
        import pickle

        with open(load_path, "rb") as fh:
            obj = pickle.load(fh)
        if isinstance(obj, GpuVecNormalize):
            obj.set_venv(venv)
            return obj
        return GpuVecNormalize.from_reference(obj, venv)  # a pickle written by the reference's VecNormalize.save

    @classmethod
    def from_reference(cls, ref, venv: GpuCSTRVecEnv) -> "GpuVecNormalize":
        """Adopt the statistics and settings of a reference ``VecNormalize`` (e.g. unpickled from ``VecNormalize.save``)."""
        out = cls(venv, training=ref.training, norm_obs=ref.norm_obs, norm_reward=ref.norm_reward, clip_obs=ref.clip_obs,
                  clip_reward=ref.clip_reward, gamma=ref.gamma, epsilon=ref.epsilon)
        out.load_state_dict({"obs_mean": np.asarray(ref.obs_rms.mean, np.float64), "obs_var": np.asarray(ref.obs_rms.var, np.float64),
                             "obs_count": float(ref.obs_rms.count), "ret_mean": float(ref.ret_rms.mean), "ret_var": float(ref.ret_rms.var),
                             "ret_count": float(ref.ret_rms.count)})
        return out

    def to_reference(self, vec_normalize_class: type, venv) -> Any:
        """Build the reference's ``VecNormalize`` around ``venv`` carrying these statistics."""
        d = self.state_dict()
        ref = vec_normalize_class(venv, training=d["training"], norm_obs=d["norm_obs"], norm_reward=d["norm_reward"], clip_obs=d["clip_obs"],
                                  clip_reward=d["clip_reward"], gamma=d["gamma"], epsilon=d["epsilon"])
        ref.obs_rms.mean, ref.obs_rms.var, ref.obs_rms.count = d["obs_mean"], d["obs_var"], d["obs_count"]
        ref.ret_rms.mean, ref.ret_rms.var, ref.ret_rms.count = np.float64(d["ret_mean"]), np.float64(d["ret_var"]), d["ret_count"]
        return ref

    # ---- pass-through --------------------------------------------------------------------------------------------------
    def __getattr__(self, name: str):  # attributes of the wrapped env (VecEnvWrapper.__getattr__, base_vec_env.py:425-439)
        if name.startswith("_") or name == "venv":
            raise AttributeError(name)
        venv = self.__dict__.get("venv")
        if venv is None:
            raise AttributeError(name)
        return getattr(venv, name)

    def seed(self, seed: Optional[int] = None):
        return self.venv.seed(seed)

    def set_options(self, options=None) -> None:
        return self.venv.set_options(options)

    def close(self) -> None:
        self.venv.close()

    def render(self, mode: Optional[str] = None):
        return self.venv.render(mode)

    def get_images(self):
        return self.venv.get_images()

    def get_attr(self, attr_name: str, indices=None):
        return self.venv.get_attr(attr_name, indices)

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        self.venv.set_attr(attr_name, value, indices)

    def env_method(self, method_name: str, *args, indices=None, **kwargs):
        return self.venv.env_method(method_name, *args, indices=indices, **kwargs)

    def env_is_wrapped(self, wrapper_class, indices=None):
        return self.venv.env_is_wrapped(wrapper_class, indices)

    @property
    def unwrapped(self):
        return self.venv.unwrapped


def _unpickle(state: dict) -> GpuVecNormalize:
    obj = GpuVecNormalize.__new__(GpuVecNormalize)
    obj.__setstate__(state)
    return obj


def bind_vec_normalize_class(vec_normalize_base: type) -> type:
    """Return ``class GpuVecNormalize(GpuVecNormalize, <reference VecNormalize>)`` so that the unchanged reference finds it with
    ``unwrap_vec_normalize`` (core/common/vec_env/__init__.py:36-44 -> base_class.py:195) and then uses
    ``get_original_obs/get_original_reward/unnormalize_obs`` in ``_store_transition`` (off_policy_algorithm.py:468-494) and passes it
    as ``env=`` to ``replay_buffer.sample`` (td3.py:161), which lands in the fused gather.  The reference ``__init__`` is not run."""

    class BoundGpuVecNormalize(GpuVecNormalize, vec_normalize_base):  # type: ignore[misc, valid-type]
        pass

    BoundGpuVecNormalize.__name__ = "GpuVecNormalize"
    BoundGpuVecNormalize.__qualname__ = "GpuVecNormalize"
    return BoundGpuVecNormalize
