"""GPU-resident two-series CSTR environments behind the reference's env protocols.

* :class:`GpuCSTRVecEnv` — the ``VecEnv`` surface of ``core/common/vec_env/base_vec_env.py:50-357`` with
  the step semantics of ``DummyVecEnv.step_wait`` (``dummy_vec_env.py:56-73``) for N independent
  ``TwoSeriesCSTREnv`` instances (``twoseriescstr.py``), all state in HBM, one CUDA thread per reactor.
* :class:`TwoSeriesCSTREnv` — single-reactor ``gym.Env`` façade with the reference's constructor,
  attributes and ``reset``/``step`` contract (``twoseriescstr.py:63-112,226-269,394-454``).
* :func:`bind_vec_env_class` — derives a class that ALSO inherits the reference's ``VecEnv`` so that
  ``isinstance(env, VecEnv)`` in ``core/common/base_class.py:232`` accepts it unwrapped.

Every numeric result comes from ``libcstr_b200.so`` (include/cstr_b200.h); there is no CPU path.
Host code here is plumbing: buffers, streams, H2D/D2H of the protocol's NumPy arrays, info dicts.
"""
from __future__ import annotations

import random as _py_random
import time
from collections.abc import Sequence
from copy import deepcopy
from ctypes import byref
from typing import Any, Dict, Iterable, List, Optional, Union

import numpy as np

from . import _lib
from ._spaces import action_space as _action_space
from ._spaces import observation_space as _observation_space

try:  # the façade is a real gym.Env when gymnasium is importable
    import gymnasium as _gym

    _EnvBase = _gym.Env
except ImportError:  # pragma: no cover
    _gym = None
    _EnvBase = object

# raw physical bounds, float32 as in twoseriescstr.py:56-61
RAW_STATE_LOW = np.array([0.0, 273.15, 0.0, 273.15], dtype=np.float32)
RAW_STATE_HIGH = np.array([0.7, 400.0, 0.7, 400.0], dtype=np.float32)
RAW_ACTION_LOW = np.array([30.0, 30.0], dtype=np.float32)
RAW_ACTION_HIGH = np.array([250.0, 250.0], dtype=np.float32)
STATIC_INIT_STATE = (0.45, 310.0, 0.25, 290.0)  # twoseriescstr.py:96
MAX_STEPS = 400

_MATH = {"strict": _lib.MATH_STRICT, "fast": _lib.MATH_FAST}
_INIT = {"random": _lib.INIT_RANDOM, "static": _lib.INIT_STATIC}


# --------------------------------------------------------------------------------------------------
# host-side reset draws for reset_rng="pcg64" (bit-parity with the reference's per-env PCG64 streams)
# --------------------------------------------------------------------------------------------------
def _pcg64(seed: Optional[int]) -> np.random.Generator:
    """gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))."""
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def _initial_raw_state(gen: np.random.Generator, init_mode: str, static_base: Optional[np.ndarray]) -> np.ndarray:
    """generate_initial_state / the static branch of reset (twoseriescstr.py:167-224,245-253), float64,
    same draw order as the reference so a same-seeded PCG64 stream yields the same state."""
    if init_mode == "random":
        s = np.array([gen.uniform(0.05, 0.45), gen.uniform(280, 380), gen.uniform(0.05, 0.45 * 0.8), gen.uniform(280, 380)])
        s += gen.uniform(-0.05, 0.05, size=4)
        if s[1] < s[3]:
            s[1], s[3] = s[3], s[1]
        if s[0] < s[2]:
            s[0], s[2] = s[2], s[0]
        return np.clip(s, RAW_STATE_LOW, RAW_STATE_HIGH)
    assert static_base is not None
    static_base += gen.uniform([-0.05, -10, -0.05, -10], [0.05, 10, 0.05, 10], size=4)  # in place: quirk Q2
    return static_base


def _normalize_f64(raw: np.ndarray) -> np.ndarray:
    return 2.0 * (raw - RAW_STATE_LOW) / (RAW_STATE_HIGH - RAW_STATE_LOW) - 1.0


# --------------------------------------------------------------------------------------------------
# lazy infos
# --------------------------------------------------------------------------------------------------
class LazyInfos(Sequence):
    """The ``infos`` value of ``step_wait``: behaves like ``list[dict]`` of length N but only
    materialises dicts for rows that finished an episode.

    Keys the unchanged reference reads: ``"TimeLimit.truncated"`` (buffers.py:278),
    ``"terminal_observation"`` (off_policy_algorithm.py:481,491), ``"episode"`` (base_class.py:476).
    Non-done rows share one read-only ``{"TimeLimit.truncated": False}``-like mapping.
    ``timeouts`` / ``timeouts_device`` give the whole vector without a Python loop.
    """

    def __init__(self, n: int, done_rows: Dict[int, dict], timeouts: np.ndarray, timeouts_device=None):
        self._n = n
        self._done = done_rows
        self.timeouts = timeouts
        self.timeouts_device = timeouts_device
        self._blank = _FrozenInfo()

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._done.get(i, self._blank)

    def __iter__(self):
        done, blank = self._done, self._blank
        if not done:
            return iter([blank] * self._n)
        return (done.get(i, blank) for i in range(self._n))

    def done_items(self):
        return self._done.items()

    def __deepcopy__(self, memo):
        return LazyInfos(self._n, deepcopy(self._done, memo), self.timeouts.copy(), self.timeouts_device)


class _Repeat(Sequence):
    """Length-n sequence whose items are produced on demand (stands in for ``[x for _ in range(n)]``)."""

    def __init__(self, n: int, factory):
        self._n, self._factory = n, factory

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._factory() for _ in range(*i.indices(self._n))]
        if not -self._n <= i < self._n:
            raise IndexError(i)
        return self._factory()


class _FrozenInfo(dict):
    """Shared info dict of the not-done rows: reads behave like ``{"TimeLimit.truncated": False}``;
    writes go to a private copy semantics-wise by being ignored (the reference never writes them)."""

    def __init__(self):
        super().__init__({"TimeLimit.truncated": False})

    def __setitem__(self, k, v):  # wrappers that annotate infos get a no-op on shared rows
        pass

    def __deepcopy__(self, memo):
        return self

    def copy(self):
        return dict(self)


# --------------------------------------------------------------------------------------------------
# the vectorised environment
# --------------------------------------------------------------------------------------------------
class GpuCSTRVecEnv:
    """N two-series CSTR reactors stepped by one CUDA kernel launch per ``step``.

    Parameters mirror the reference where it has them (``init_mode``, ``default_target``,
    ``render_mode``); the rest select the device behaviour:

    :param num_envs: number of independent reactors on this device (this rank's shard).
    :param device: CUDA device.
    :param dtype: ``"fp32"`` (NumPy>=2 semantics of the reference) or ``"fp64"``.
    :param math: ``"strict"`` = reference association bit for bit (tolerances in DESIGN.md) or ``"fast"``.
    :param reset_rng: ``"philox"`` = resets drawn in-kernel (counter = global env id, episode);
        ``"pcg64"`` = the reference's per-env ``Generator(PCG64(SeedSequence(seed+idx)))`` streams on
        the host, bit-identical initial states (compat mode, small N).
    :param env_offset: global id of reactor 0 (multi-GPU sharding: results do not depend on the split).
    :param info_mode: ``"lazy"`` (dicts only for done rows) or ``"full"`` (adds the reference's
        per-step keys ``reward, truncated, state, original_state, target_C2, step`` to every row).
    """

    metadata = {"render_modes": []}

    def __init__(
        self,
        num_envs: int,
        device: Union[str, Any] = "cuda",
        dtype: str = "fp32",
        math: str = "strict",
        init_mode: str = "random",
        seed: int = 0,
        reset_rng: str = "philox",
        default_target: float = 0.20,
        env_offset: int = 0,
        info_mode: str = "lazy",
        render_mode: Optional[str] = None,
        monitor: bool = True,
    ):
        if num_envs <= 0:
            raise ValueError("num_envs must be positive")
        if dtype not in ("fp32", "fp64"):
            raise ValueError("dtype must be 'fp32' or 'fp64'")
        if math not in _MATH:
            raise ValueError("math must be 'strict' or 'fast'")
        if init_mode not in _INIT:
            raise ValueError(f"init_mode={init_mode} is not supported, please choose 'random' or 'static'")
        if reset_rng not in ("philox", "pcg64"):
            raise ValueError("reset_rng must be 'philox' or 'pcg64'")
        if dtype == "fp64" and math == "fast":
            raise ValueError("math='fast' exists for fp32 only")
        torch = _lib.require_cuda()
        self._torch = torch
        self._libc = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CstrLibraryError("GpuCSTRVecEnv needs a CUDA device (no CPU fallback)")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.observation_space = _observation_space()
        self.action_space = _action_space()
        self.render_mode = render_mode
        # per-env Python lists only while they are cheap; beyond that constant-time views
        self._small = self.num_envs <= 65536
        self.reset_infos = [{} for _ in range(num_envs)] if self._small else _Repeat(num_envs, dict)
        self._seeds = [None] * num_envs if self._small else _Repeat(num_envs, lambda: None)
        self._options = [{} for _ in range(num_envs)] if self._small else _Repeat(num_envs, dict)
        self.dtype, self.math, self.init_mode, self.reset_rng = dtype, math, init_mode, reset_rng
        self.info_mode = info_mode
        self.target_C2 = float(default_target)
        self.min_concentration, self.max_concentration = 0.05, 0.45
        self.max_steps = MAX_STEPS
        self.monitor = monitor
        self._params = _lib.EnvParams(seed=int(seed) & (2**64 - 1), env_offset=int(env_offset), target_c2=self.target_C2,
                                      max_steps=MAX_STEPS, init_mode=_INIT[init_mode])
        fdt = torch.float32 if dtype == "fp32" else torch.float64
        self._fdt = fdt
        n = self.num_envs
        with torch.cuda.device(self.device):
            self.state = torch.zeros((n, 4), dtype=fdt, device=self.device)
            self.step_count = torch.zeros(n, dtype=torch.int32, device=self.device)
            self.episode = torch.zeros(n, dtype=torch.int32, device=self.device)
            self.static_base = None
            if init_mode == "static":
                self.static_base = torch.tensor(STATIC_INIT_STATE, dtype=torch.float64, device=self.device).repeat(n, 1).contiguous()
            self._actions = torch.zeros((n, 2), dtype=fdt, device=self.device)
            self._terminal = torch.zeros((n, 4), dtype=fdt, device=self.device)
            self._reward = torch.zeros(n, dtype=fdt, device=self.device)
            self._done = torch.zeros(n, dtype=torch.uint8, device=self.device)
            self._timeout = torch.zeros(n, dtype=torch.uint8, device=self.device)
            # episode statistics (Monitor semantics, monitor.py:85-111): accumulated by the step kernel
            self._ep_return = torch.zeros(n, dtype=torch.float64, device=self.device) if monitor else None
            self._ep_final_return = torch.zeros(n, dtype=torch.float64, device=self.device) if monitor else None
            self._ep_final_length = torch.zeros(n, dtype=torch.int32, device=self.device) if monitor else None
        # pinned staging for the NumPy protocol
        ndt = np.float32 if dtype == "fp32" else np.float64
        self._ndt = ndt
        self._h_actions = torch.zeros((n, 2), dtype=fdt).pin_memory()
        self._h_obs = torch.zeros((n, 4), dtype=fdt).pin_memory()
        self._h_terminal = torch.zeros((n, 4), dtype=fdt).pin_memory()
        self._h_reward = torch.zeros(n, dtype=fdt).pin_memory()
        self._h_done = torch.zeros(n, dtype=torch.uint8).pin_memory()
        self._pending = False
        self._needs_reset = True
        self._t_start = time.time()
        self._gens: Optional[List[np.random.Generator]] = None  # pcg64 mode
        self._host_static_base: Optional[np.ndarray] = None
        if reset_rng == "pcg64":
            if n > 65536:
                raise ValueError("reset_rng='pcg64' is the host compat mode; use 'philox' for large num_envs")
            self._pcg_base_seed: Optional[int] = int(seed)
            if init_mode == "static":
                self._host_static_base = np.tile(np.array(STATIC_INIT_STATE, np.float64), (n, 1))
        self.launches = 0  # kernels this object launched (bench.py's gpu_launches bookkeeping)

    # ---- helpers -------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self.device).cuda_stream

    def _sb_ptr(self):
        return _lib.ptr(self.static_base)

    def set_target(self, target: float) -> bool:
        """twoseriescstr.py:114-127."""
        if self.min_concentration <= target <= self.max_concentration:
            self.target_C2 = float(target)
            self._params.target_c2 = float(target)
            return True
        return False

    # ---- VecEnv protocol -------------------------------------------------------------------------------
    def seed(self, seed: Optional[int] = None) -> Sequence:
        """base_vec_env.py:292-309: per-env seeds ``seed + idx``, applied at the next ``reset()``.
        philox mode: the Philox key becomes ``seed`` and the env id supplies the ``+ idx``."""
        if seed is None:
            seed = int(np.random.randint(0, np.iinfo(np.uint32).max, dtype=np.uint32))
        self._pending_seed = int(seed)
        if self._small:
            self._seeds = [seed + idx for idx in range(self.num_envs)]
            return self._seeds
        return range(seed, seed + self.num_envs)

    def set_options(self, options: Optional[Union[List[dict], dict]] = None) -> None:
        if self._small:
            if options is None:
                options = {}
            self._options = deepcopy([options] * self.num_envs) if isinstance(options, dict) else deepcopy(options)

    def reset(self) -> np.ndarray:
        """base_vec_env.py:109 / dummy_vec_env.py:75-83: reset every reactor, return obs (N,4)."""
        torch = self._torch
        pending = getattr(self, "_pending_seed", None)
        with torch.cuda.device(self.device):
            if self.reset_rng == "philox":
                if pending is not None:
                    self._params.seed = pending & (2**64 - 1)
                    self.episode.zero_()
                _lib.check(self._libc.cstr_reset(byref(self._params), self.num_envs, None, _lib.ptr(self.state), int(self.dtype == "fp64"),
                                                 _lib.ptr(self.step_count), _lib.ptr(self.episode), self._sb_ptr(), self._stream()), "cstr_reset")
                self.launches += 1
            else:
                if pending is not None or self._gens is None:
                    base = pending if pending is not None else self._pcg_base_seed
                    self._gens = [_pcg64(None if base is None else base + i) for i in range(self.num_envs)]
                self._host_reset_rows(np.arange(self.num_envs))
            if self.monitor:
                self._ep_return.zero_()
        self._pending_seed = None
        if self._small:
            self._seeds = [None] * self.num_envs
            self._options = [{} for _ in range(self.num_envs)]
        self._needs_reset = False
        self._pending = False
        return self._obs_to_host()

    def _host_reset_rows(self, rows: np.ndarray) -> np.ndarray:
        """pcg64 mode: draw new initial states for ``rows`` on the host and scatter them to the device."""
        torch = self._torch
        new = np.empty((len(rows), 4), self._ndt)
        for k, i in enumerate(rows):
            sb = None if self._host_static_base is None else self._host_static_base[i]
            raw = _initial_raw_state(self._gens[i], self.init_mode, sb)
            new[k] = _normalize_f64(np.asarray(raw, np.float64)).astype(self._ndt)
        idx = torch.as_tensor(rows, dtype=torch.int64, device=self.device)
        self.state.index_copy_(0, idx, torch.as_tensor(new, device=self.device))
        self.step_count.index_fill_(0, idx, 0)
        self.episode.index_add_(0, idx, torch.ones(len(rows), dtype=torch.int32, device=self.device))
        return new

    def _obs_to_host(self) -> np.ndarray:
        self._h_obs.copy_(self.state, non_blocking=True)
        self._torch.cuda.current_stream(self.device).synchronize()
        return self._h_obs.numpy().copy()

    def step_async(self, actions: np.ndarray) -> None:
        """base_vec_env.py:123: H2D the (N,2) actions and enqueue the step kernel (returns immediately)."""
        if self._needs_reset:
            raise ValueError("Please call env.reset() to reset the env first!")  # twoseriescstr.py:402
        torch = self._torch
        if isinstance(actions, torch.Tensor):
            self._actions.copy_(actions.reshape(self.num_envs, 2), non_blocking=True)
        else:
            a = np.asarray(actions, dtype=self._ndt).reshape(self.num_envs, 2)
            self._h_actions.numpy()[...] = a
            self._actions.copy_(self._h_actions, non_blocking=True)
        self._launch_step(self._actions)
        self._pending = True

    def _launch_step(self, actions_dev) -> None:
        auto_reset = int(self.reset_rng == "philox")
        with self._torch.cuda.device(self.device):
            if self.dtype == "fp32":
                rc = self._libc.cstr_vec_step_f32(byref(self._params), self.num_envs, _MATH[self.math], auto_reset, _lib.ptr(actions_dev),
                                                  _lib.ptr(self.state), _lib.ptr(self.step_count), _lib.ptr(self.episode), self._sb_ptr(),
                                                  _lib.ptr(self._terminal), _lib.ptr(self._reward), _lib.ptr(self._done),
                                                  _lib.ptr(self._timeout), _lib.ptr(self._ep_return), _lib.ptr(self._ep_final_return),
                                                  _lib.ptr(self._ep_final_length), self._stream())
            else:
                rc = self._libc.cstr_vec_step_f64(byref(self._params), self.num_envs, auto_reset, _lib.ptr(actions_dev), _lib.ptr(self.state),
                                                  _lib.ptr(self.step_count), _lib.ptr(self.episode), self._sb_ptr(), _lib.ptr(self._terminal),
                                                  _lib.ptr(self._reward), _lib.ptr(self._done), _lib.ptr(self._timeout),
                                                  _lib.ptr(self._ep_return), _lib.ptr(self._ep_final_return),
                                                  _lib.ptr(self._ep_final_length), self._stream())
        _lib.check(rc, "cstr_vec_step")
        self.launches += 1

    def step_wait(self):
        """base_vec_env.py:135 with DummyVecEnv.step_wait semantics (dummy_vec_env.py:56-73):
        returns fresh ``(obs (N,4), rewards (N,), dones (N,) bool, infos)``; on done rows ``obs`` is the
        post-reset observation and ``infos[i]["terminal_observation"]`` the last one."""
        if not self._pending:
            raise RuntimeError("step_wait() called without step_async()")
        torch = self._torch
        self._pending = False
        self._h_obs.copy_(self.state, non_blocking=True)
        self._h_reward.copy_(self._reward, non_blocking=True)
        self._h_done.copy_(self._done, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        obs = self._h_obs.numpy().copy()
        rewards = self._h_reward.numpy().astype(np.float32)  # buf_rews is float32 (dummy_vec_env.py:46)
        dones = self._h_done.numpy().astype(bool)
        done_rows: Dict[int, dict] = {}
        if dones.any():
            rows = np.nonzero(dones)[0]
            idx = torch.as_tensor(rows, dtype=torch.int64, device=self.device)
            term = self._terminal.index_select(0, idx).cpu().numpy()
            if self.monitor:
                ep_r = self._ep_final_return.index_select(0, idx).cpu().numpy()
                ep_l = self._ep_final_length.index_select(0, idx).cpu().numpy()
            if self.reset_rng == "pcg64":
                obs[rows] = self._host_reset_rows(rows)
            for k, i in enumerate(rows):
                info = {"TimeLimit.truncated": True, "terminal_observation": term[k]}
                if self.monitor:  # monitor.py:96-109
                    info["episode"] = {"r": round(float(ep_r[k]), 6), "l": int(ep_l[k]), "t": round(time.time() - self._t_start, 6)}
                done_rows[int(i)] = info
        infos = LazyInfos(self.num_envs, done_rows, dones.copy(), self._timeout)
        if self.info_mode == "full":
            infos = self._full_infos(infos, obs, rewards, dones)
        return obs, rewards, dones, infos

    def _full_infos(self, lazy: LazyInfos, obs, rewards, dones) -> List[dict]:
        sc = self.step_count.cpu().numpy()
        out = []
        for i in range(self.num_envs):
            d = dict(lazy[i])
            st = d.get("terminal_observation", obs[i])
            d.update({"reward": rewards[i], "truncated": bool(dones[i]), "state": st,
                      "original_state": _denormalize_state(np.asarray(st, np.float32)), "target_C2": self.target_C2,
                      "step": self.max_steps if dones[i] else int(sc[i])})
            out.append(d)
        return out

    def step(self, actions: np.ndarray):
        self.step_async(actions)
        return self.step_wait()

    def close(self) -> None:
        pass

    # ---- device-resident fast path (no host sync, no copies) --------------------------------------------
    def step_tensor(self, actions):
        """Fast mode: ``actions`` is a device tensor (N,2); returns the INTERNAL device tensors
        ``(obs, reward, done_u8, terminal_obs)`` without synchronising — valid until the next step."""
        if self._needs_reset:
            raise ValueError("Please call env.reset() to reset the env first!")
        if self.reset_rng != "philox":
            raise ValueError("step_tensor needs reset_rng='philox' (device-side auto-reset)")
        a = actions if (actions.dtype == self._fdt and actions.is_contiguous()) else actions.to(self._fdt).contiguous()
        self._launch_step(a)
        return self.state, self._reward, self._done, self._terminal

    def tape(self, T: int, actions=None, t_base: int = 0, want_rewards: bool = True, want_dones: bool = True,
             want_obs: bool = False, reward_sum: bool = False):
        """T control intervals in ONE launch with the state in registers (cstr_tape_*).  ``actions`` is a
        device tensor (T,N,2) or None for in-kernel U(-1,1) Philox actions.  Returns a dict of device
        tensors ``rewards (T,N)``, ``dones (T,N) uint8``, ``obs (T,N,4)``, ``reward_sum (1,) f64``."""
        torch = self._torch
        if self.reset_rng != "philox":
            raise ValueError("tape needs reset_rng='philox'")
        if self._needs_reset:
            raise ValueError("Please call env.reset() to reset the env first!")
        n = self.num_envs
        out: Dict[str, Any] = {}
        with torch.cuda.device(self.device):
            rew = torch.empty((T, n), dtype=self._fdt, device=self.device) if want_rewards else None
            don = torch.empty((T, n), dtype=torch.uint8, device=self.device) if want_dones else None
            obs = torch.empty((T, n, 4), dtype=self._fdt, device=self.device) if want_obs else None
            rs = torch.zeros(1, dtype=torch.float64, device=self.device) if reward_sum else None
            if actions is not None:
                if tuple(actions.shape) != (T, n, 2) or actions.dtype != self._fdt or not actions.is_contiguous():
                    raise ValueError(f"actions must be a contiguous {self._fdt} tensor of shape {(T, n, 2)}")
            if self.dtype == "fp32":
                rc = self._libc.cstr_tape_f32(byref(self._params), n, T, _MATH[self.math], _lib.ptr(actions), t_base & 0xFFFFFFFF,
                                              _lib.ptr(self.state), _lib.ptr(self.step_count), _lib.ptr(self.episode), self._sb_ptr(),
                                              _lib.ptr(rew), _lib.ptr(don), _lib.ptr(obs), _lib.ptr(rs), self._stream())
            else:
                rc = self._libc.cstr_tape_f64(byref(self._params), n, T, _lib.ptr(actions), t_base & 0xFFFFFFFF, _lib.ptr(self.state),
                                              _lib.ptr(self.step_count), _lib.ptr(self.episode), self._sb_ptr(), _lib.ptr(rew),
                                              _lib.ptr(don), _lib.ptr(obs), _lib.ptr(rs), self._stream())
        _lib.check(rc, "cstr_tape")
        self.launches += 1
        out.update(rewards=rew, dones=don, obs=obs, reward_sum=rs)
        return out

    def set_state(self, state, step_count=None) -> None:
        """Inject normalised states (parity tests / restoring a snapshot)."""
        torch = self._torch
        self.state.copy_(torch.as_tensor(np.asarray(state), dtype=self._fdt).reshape(self.num_envs, 4))
        if step_count is not None:
            self.step_count.copy_(torch.as_tensor(np.asarray(step_count), dtype=torch.int32))
        self._needs_reset = False

    # ---- attribute plumbing (base_vec_env.py:166-212, dummy_vec_env.py:116-142) ---------------------------
    def _get_indices(self, indices) -> Iterable[int]:
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices

    _PER_ENV = ("current_step", "state")

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        idx = list(self._get_indices(indices))
        if attr_name == "current_step":
            sc = self.step_count.cpu().numpy()
            return [int(sc[i]) for i in idx]
        if attr_name == "state":
            st = self.state.cpu().numpy()
            return [st[i].copy() for i in idx]
        shared = {
            "render_mode": self.render_mode, "target_C2": self.target_C2, "max_steps": self.max_steps, "init_mode": self.init_mode,
            "min_concentration": self.min_concentration, "max_concentration": self.max_concentration,
            "raw_state_low": RAW_STATE_LOW, "raw_state_high": RAW_STATE_HIGH, "raw_action_low": RAW_ACTION_LOW,
            "raw_action_high": RAW_ACTION_HIGH, "observation_space": self.observation_space, "action_space": self.action_space,
            "metadata": self.metadata, "dt": 0.1, "spec": None,
        }
        if attr_name not in shared:
            raise AttributeError(f"TwoSeriesCSTREnv has no attribute {attr_name!r}")
        return [shared[attr_name] for _ in idx]

    def has_attr(self, attr_name: str) -> bool:
        try:
            self.get_attr(attr_name, indices=0)
            return True
        except AttributeError:
            return False

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        if attr_name == "target_C2":
            self.target_C2 = float(value)
            self._params.target_c2 = float(value)
        elif attr_name == "render_mode":
            self.render_mode = value
        elif attr_name == "max_steps":  # plain attribute in the reference (twoseriescstr.py:99,438)
            if int(value) <= 0:
                raise ValueError("max_steps must be positive")
            self.max_steps = int(value)
            self._params.max_steps = int(value)
        else:
            raise AttributeError(f"attribute {attr_name!r} cannot be set on the batched CSTR env")

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
        idx = list(self._get_indices(indices))
        if method_name == "set_target":
            ok = self.set_target(*method_args, **method_kwargs)
            return [ok for _ in idx]
        if method_name == "_normalize_state":
            return [_normalize_state(*method_args, **method_kwargs) for _ in idx]
        if method_name == "_denormalize_state":
            return [_denormalize_state(*method_args, **method_kwargs) for _ in idx]
        if method_name == "render":
            return [None for _ in idx]
        raise AttributeError(f"method {method_name!r} is not available on the batched CSTR env")

    def env_is_wrapped(self, wrapper_class, indices=None) -> List[bool]:
        """evaluate_policy asks ``env_is_wrapped(Monitor)[0]`` (evaluation.py:64): episode statistics are
        produced natively (``info["episode"]``) when ``monitor=True``."""
        name = getattr(wrapper_class, "__name__", "")
        return [bool(self.monitor and name == "Monitor") for _ in self._get_indices(indices)]

    def get_images(self):
        return [None for _ in range(self.num_envs)]

    def render(self, mode: Optional[str] = None):
        return None

    @property
    def unwrapped(self):
        return self

    def getattr_depth_check(self, name: str, already_found: bool) -> Optional[str]:
        if hasattr(self, name) and already_found:
            return f"{type(self).__module__}.{type(self).__name__}"
        return None


def bind_env_class(gym_env_base: type) -> type:
    """Return ``class TwoSeriesCSTREnv(TwoSeriesCSTREnv, <gymnasium.Env>)`` for a ``gymnasium`` that became importable only after this
    module was imported (a harness that puts it on ``sys.path`` late): the reference's ``_patch_env`` (core/common/vec_env/patch_gym.py:31-35)
    insists on ``isinstance(env, gymnasium.Env)``.  When gymnasium was importable at import time the façade already is one."""
    if issubclass(TwoSeriesCSTREnv, gym_env_base):
        return TwoSeriesCSTREnv

    class BoundTwoSeriesCSTREnv(TwoSeriesCSTREnv, gym_env_base):  # type: ignore[misc, valid-type]
        pass

    BoundTwoSeriesCSTREnv.__name__ = "TwoSeriesCSTREnv"
    BoundTwoSeriesCSTREnv.__qualname__ = "TwoSeriesCSTREnv"
    return BoundTwoSeriesCSTREnv


def bind_vec_env_class(vec_env_base: type) -> type:
    """Return ``class GpuCSTRVecEnv(GpuCSTRVecEnv, <reference VecEnv>)`` so the unchanged reference
    accepts the env without wrapping it (``isinstance(env, VecEnv)``, core/common/base_class.py:232).
    ``VecEnv.__init__`` is not run: it would call ``get_attr("render_mode")`` per env (base_vec_env.py:76)."""

    class BoundGpuCSTRVecEnv(GpuCSTRVecEnv, vec_env_base):  # type: ignore[misc, valid-type]
        pass

    BoundGpuCSTRVecEnv.__name__ = "GpuCSTRVecEnv"
    BoundGpuCSTRVecEnv.__qualname__ = "GpuCSTRVecEnv"
    return BoundGpuCSTRVecEnv


# --------------------------------------------------------------------------------------------------
# affine helpers with the reference's float32 semantics (twoseriescstr.py:129-150)
# --------------------------------------------------------------------------------------------------
def _normalize_state(raw_state: np.ndarray) -> np.ndarray:
    return (2.0 * (raw_state - RAW_STATE_LOW) / (RAW_STATE_HIGH - RAW_STATE_LOW) - 1.0).astype(np.float32)


def _denormalize_state(normalized_state: np.ndarray) -> np.ndarray:
    return (RAW_STATE_LOW + (normalized_state + 1.0) * (RAW_STATE_HIGH - RAW_STATE_LOW) / 2.0).astype(np.float32)


def _denormalize_action(normalized_action: np.ndarray) -> np.ndarray:
    return (RAW_ACTION_LOW + (normalized_action + 1.0) * (RAW_ACTION_HIGH - RAW_ACTION_LOW) / 2.0).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# single-reactor gym.Env façade
# --------------------------------------------------------------------------------------------------
class TwoSeriesCSTREnv(_EnvBase):  # type: ignore[misc, valid-type]
    """Drop-in for the reference's ``TwoSeriesCSTREnv`` (twoseriescstr.py:15-519): same constructor,
    spaces, attributes and ``reset``/``step`` return values; the arithmetic of ``step`` runs in the
    CUDA step kernel (strict fp32), the reset draws come from the same per-env PCG64 stream."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 4}
    raw_state_low, raw_state_high = RAW_STATE_LOW, RAW_STATE_HIGH
    raw_action_low, raw_action_high = RAW_ACTION_LOW, RAW_ACTION_HIGH
    dt = 0.1

    def __init__(self, render_mode: Optional[str] = None, default_target: float = 0.20, min_concentration: float = 0.05,
                 max_concentration: float = 0.45, init_mode: str = "random", device: Union[str, Any] = "cuda", math: str = "strict"):
        super().__init__()
        if init_mode not in _INIT:
            raise ValueError(f"init_mode={init_mode} is not supported, please choose 'random' or 'static'")
        self.render_mode = render_mode
        self.observation_space = _observation_space()
        self.action_space = _action_space()
        self.init_mode = init_mode
        self.init_state = None if init_mode == "random" else np.array(STATIC_INIT_STATE)
        self.max_steps = MAX_STEPS
        self.current_step = 0
        self.target_C2 = default_target
        self.min_concentration, self.max_concentration = min_concentration, max_concentration
        self.initial_state_info: Dict[str, Any] = {}
        self.state: Optional[np.ndarray] = None
        self._gen: Optional[np.random.Generator] = None
        self._vec = GpuCSTRVecEnv(1, device=device, math=math, init_mode="random", reset_rng="pcg64", default_target=default_target,
                                  monitor=False)

    # the reference exposes these as methods (used by evaluate_model through env.envs[0], :541)
    _normalize_state = staticmethod(_normalize_state)
    _denormalize_state = staticmethod(_denormalize_state)
    _denormalize_action = staticmethod(_denormalize_action)

    def set_target(self, target) -> bool:
        if self.min_concentration <= target <= self.max_concentration:
            self.target_C2 = target
            self._vec.set_attr("target_C2", target)
            return True
        return False

    def seed(self, seed: Optional[int] = None):
        """twoseriescstr.py:152-165 — also reseeds the GLOBAL ``random`` / ``np.random`` (quirk Q3)."""
        ss = np.random.SeedSequence(seed)
        self._gen = np.random.Generator(np.random.PCG64(ss))
        seed = ss.entropy
        _py_random.seed(seed)
        np.random.seed(seed % (2**32) if seed >= 2**32 else seed)
        return [seed]

    def reset(self, *, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        if seed is not None:
            self.seed(seed)
        if self._gen is None:
            self._gen = _pcg64(None)  # gymnasium creates an unseeded generator lazily
        raw = _initial_raw_state(self._gen, self.init_mode, self.init_state)
        self.initial_state_info = {
            "initial_concentration_1": raw[0], "initial_temperature_1": raw[1],
            "initial_concentration_2": raw[2], "initial_temperature_2": raw[3],
        }
        self.current_step = 0
        self.state = _normalize_f64(np.asarray(raw, np.float64)).astype(np.float32)
        self._vec.set_state(self.state[None, :], np.zeros(1, np.int32))
        return self.state.astype(np.float32), self.initial_state_info

    def step(self, action: np.ndarray):
        if self.state is None:
            raise ValueError("Please call env.reset() to reset the env first!")
        action = np.asarray(action, np.float32).reshape(2)
        vec = self._vec
        vec.step_async(action[None, :])
        # raw (un-reset) outputs: the façade owns episode boundaries, like the reference's gym.Env
        torch = vec._torch
        torch.cuda.current_stream(vec.device).synchronize()
        vec._pending = False
        # auto_reset is off in pcg64 mode: vec.state is the new state and vec.step_count the incremented counter
        new_state = vec.state.cpu().numpy()[0].astype(np.float32)
        reward = np.float32(vec._reward.cpu().numpy()[0])
        truncated = bool(vec._done.cpu().numpy()[0])
        self.current_step += 1
        normalized_action = np.clip(action, self.action_space.low, self.action_space.high)
        if np.isnan(action).any():  # twoseriescstr.py:415-421
            return self.state, -10.0, False, True, {"error": "检测到非法输入：状态或动作包含NaN", "raw_action": _denormalize_action(normalized_action)}
        self.state = new_state
        info = {
            "reward": reward, "raw_action": _denormalize_action(normalized_action), "truncated": truncated, "state": self.state,
            "original_state": _denormalize_state(self.state).astype(np.float64), "target_C2": self.target_C2, "step": self.current_step,
            "concentration_error": np.abs(_denormalize_state(self.state)[2] - np.float32(self.target_C2)),
        }
        return self.state, reward, False, truncated, info

    def render(self):
        if self.render_mode == "human" and self.state is not None:
            C1, T1, C2, T2 = _denormalize_state(self.state)
            print(f"Step: {self.current_step}")
            print(f"Reactor 1: C1={C1:.4f} mol/L, T1={T1:.2f} K")
            print(f"Reactor 2: C2={C2:.4f} mol/L, T2={T2:.2f} K")
            print(f"Target C2: {self.target_C2:.4f} mol/L")
            print(f"Error: {np.abs(C2 - self.target_C2):.4f} mol/L")
            print("-" * 50)

    def close(self):
        pass
