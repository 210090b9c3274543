"""GPU-resident ring replay buffer behind the reference's ``ReplayBuffer`` surface
(``core/common/buffers.py:158-340``: ctor, ``add``, ``sample``, ``size``, ``reset``, ``extend``,
``to_torch``, the six storage attributes, ``pos``/``full``, pickling).

Storage is ONE device tensor ``records (rows, n_envs, 16) float32`` — a 64-byte record per transition
(include/cstr_b200.h) — and the reference's attributes ``observations, next_observations, actions,
rewards, dones, timeouts`` are strided views of it, so code that indexes or assigns them
(``buffer.dones[pos] = True``, off_policy_algorithm.py:290) keeps working.  ``add`` and ``sample`` are
CUDA kernels (cstr_replay_add / cstr_replay_sample*), ``sample`` writes straight into the float32
tensors the critic/actor update consumes.  No CPU path.
"""
from __future__ import annotations

from typing import Any, NamedTuple, Optional, Sequence, Union

import numpy as np

from . import _lib
from ._spaces import action_space as _default_action_space
from ._spaces import observation_space as _default_observation_space


class ReplayBufferSamples(NamedTuple):
    """Same fields, order, shapes and dtype as core/common/type_aliases.py:49-54."""

    observations: Any
    actions: Any
    next_observations: Any
    dones: Any
    rewards: Any


def _get_device(device):
    import torch

    if device == "auto" or device is None:
        device = "cuda"
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.CstrLibraryError(f"GpuReplayBuffer needs a CUDA device, got {device} (no CPU fallback)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class GpuReplayBuffer:
    """Drop-in for ``ReplayBuffer(buffer_size, observation_space, action_space, device, n_envs,
    optimize_memory_usage, handle_timeout_termination)``.

    Extra keyword arguments (pass through ``replay_buffer_kwargs=``):

    :param index_mode: ``"numpy"`` — indices come from the reference's two global-RNG draws
        (``np.random.randint``, buffers.py:114,309) so ``np.random.seed(k); sample(B)`` returns the same
        rows as the reference, bit for bit;  ``"philox"`` — indices drawn inside the gather kernel.
    :param seed: Philox key of the ``"philox"`` index stream.  In a ``torch.distributed`` job the rank is folded into the key (rank 0 and
        single-process runs use ``seed`` itself), so data-parallel ranks that were all given the same seed do not draw the same
        (row, env) pairs for their shards.
    """

    def __init__(
        self,
        buffer_size: int,
        observation_space=None,
        action_space=None,
        device: Union[str, Any] = "auto",
        n_envs: int = 1,
        optimize_memory_usage: bool = False,
        handle_timeout_termination: bool = True,
        index_mode: str = "numpy",
        seed: int = 0,
    ):
        torch = _lib.require_cuda()
        self._torch = torch
        self._libc = _lib.load()
        if optimize_memory_usage:
            # reference: incompatible with handle_timeout_termination (buffers.py:206-210); the packed
            # record layout stores next_obs explicitly, so the memory-saving variant is not offered
            raise ValueError("GpuReplayBuffer does not support optimize_memory_usage=True")
        if index_mode not in ("numpy", "philox"):
            raise ValueError("index_mode must be 'numpy' or 'philox'")
        self.observation_space = observation_space if observation_space is not None else _default_observation_space()
        self.action_space = action_space if action_space is not None else _default_action_space()
        self.obs_shape = tuple(self.observation_space.shape)
        self.action_dim = int(np.prod(self.action_space.shape))
        if self.obs_shape != (4,) or self.action_dim != 2:
            raise ValueError("GpuReplayBuffer is specialised to the CSTR path: obs (4,), action (2,)")
        if buffer_size < n_envs:
            # Q11 / H6: the reference silently keeps max(buffer_size // n_envs, 1) rows
            raise ValueError(f"buffer_size ({buffer_size}) must be >= n_envs ({n_envs}): it counts transitions, not rows")
        self.buffer_size = max(buffer_size // n_envs, 1)  # buffers.py:198
        self.n_envs = int(n_envs)
        self.pos = 0
        self.full = False
        self.optimize_memory_usage = False
        self.handle_timeout_termination = handle_timeout_termination
        self.index_mode = index_mode
        self.seed = int(seed)
        self._key: Optional[int] = None  # seed with the rank folded in, fixed at the first Philox draw (see _philox_key)
        self._draw = 0
        self._device = _get_device(device)
        with torch.cuda.device(self._device):
            self.records = torch.zeros((self.buffer_size, self.n_envs, _lib.REC_FLOATS), dtype=torch.float32, device=self._device)
        self.launches = 0

    # ---- the reference's storage attributes, as views of the packed records ------------------------------
    @property
    def observations(self):
        return self.records[..., 0:4]

    @property
    def next_observations(self):
        return self.records[..., 4:8]

    @property
    def actions(self):
        return self.records[..., 8:10]

    @property
    def rewards(self):
        return self.records[..., 10]

    @property
    def dones(self):
        return self.records[..., 11]

    @property
    def timeouts(self):
        return self.records[..., 12]

    @property
    def device(self):
        return self._device

    @device.setter
    def device(self, value):  # load_replay_buffer does `buffer.device = self.device` (off_policy_algorithm.py:254)
        dev = _get_device(value)
        if dev != self._device:
            self.records = self.records.to(dev)
            self._device = dev

    # ---- bookkeeping ------------------------------------------------------------------------------------
    def size(self) -> int:
        return self.buffer_size if self.full else self.pos

    def reset(self) -> None:
        self.pos = 0
        self.full = False

    def _stream(self) -> int:
        return self._torch.cuda.current_stream(self._device).cuda_stream

    def _dev(self, x, dtype, shape):
        torch = self._torch
        if isinstance(x, torch.Tensor):
            t = x.to(device=self._device, dtype=dtype, non_blocking=True)
        else:
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(x)), device=self._device).to(dtype)
        return t.reshape(shape).contiguous()

    # ---- add (buffers.py:247-283) -------------------------------------------------------------------------
    def add(self, obs, next_obs, action, reward, done, infos: Optional[Sequence[dict]] = None, timeouts=None) -> None:
        """Store one row of ``n_envs`` transitions at ring position ``pos``.  Arguments may be NumPy
        arrays (the reference's call) or device tensors (zero-copy fast path).  ``timeouts`` (device/host
        vector) overrides the per-info ``"TimeLimit.truncated"`` scan."""
        torch = self._torch
        n = self.n_envs
        o = self._dev(obs, torch.float32, (n, 4))
        no = self._dev(next_obs, torch.float32, (n, 4))
        a = self._dev(action, torch.float32, (n, 2))  # action.reshape((n_envs, action_dim)), :263
        r = self._dev(reward, torch.float32, (n,))
        d = self._dev(done, torch.uint8, (n,))
        to = None
        if self.handle_timeout_termination:
            if timeouts is None and infos is not None:
                lazy_dev = getattr(infos, "timeouts_device", None)
                lazy_host = getattr(infos, "timeouts", None)
                if lazy_dev is not None and lazy_dev.device == self._device:
                    timeouts = lazy_dev
                elif lazy_host is not None:
                    timeouts = lazy_host
                else:
                    timeouts = np.array([info.get("TimeLimit.truncated", False) for info in infos])  # :278
            if timeouts is not None:
                to = self._dev(timeouts, torch.uint8, (n,))
        with torch.cuda.device(self._device):
            rc = self._libc.cstr_replay_add(n, self.pos, _lib.ptr(o), _lib.ptr(no), _lib.ptr(a), _lib.ptr(r), _lib.ptr(d), _lib.ptr(to),
                                            _lib.ptr(self.records), self._stream())
        _lib.check(rc, "cstr_replay_add")
        self.launches += 1
        self.advance(1)

    def advance(self, rows: int) -> None:
        """Move the ring cursor by ``rows`` (the fused rollout kernel writes rows itself)."""
        new = self.pos + rows
        if new >= self.buffer_size:
            self.full = True
        self.pos = new % self.buffer_size

    def extend(self, *args) -> None:  # buffers.py:91-97
        for data in zip(*args):
            self.add(*data)

    # ---- sample (buffers.py:285-325, 106-115) ----------------------------------------------------------------
    def _alloc_out(self, batch_size: int):
        torch = self._torch
        dev = self._device
        return (
            torch.empty((batch_size, 4), dtype=torch.float32, device=dev),
            torch.empty((batch_size, 2), dtype=torch.float32, device=dev),
            torch.empty((batch_size, 4), dtype=torch.float32, device=dev),
            torch.empty((batch_size, 1), dtype=torch.float32, device=dev),
            torch.empty((batch_size, 1), dtype=torch.float32, device=dev),
        )

    def sample(self, batch_size: int, env=None) -> ReplayBufferSamples:
        upper_bound = self.buffer_size if self.full else self.pos
        if upper_bound <= 0:
            raise ValueError("cannot sample from an empty replay buffer")
        if self.index_mode == "numpy":
            batch_inds = np.random.randint(0, upper_bound, size=batch_size)  # buffers.py:114
            return self._get_samples(batch_inds, env=env)
        torch = self._torch
        obs, act, nobs, dones, rew = self._alloc_out(batch_size)
        with torch.cuda.device(self._device):
            rc = self._libc.cstr_replay_sample_philox(self._philox_key(), self._draw, self.n_envs, upper_bound, batch_size,
                                                      _lib.ptr(self.records), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(nobs), _lib.ptr(dones),
                                                      _lib.ptr(rew), None, None, self._norm_arg(env), self._stream())
        _lib.check(rc, "cstr_replay_sample_philox")
        self._draw += 1
        self.launches += 1
        return self._finish(obs, act, nobs, dones, rew, env)

    def _philox_key(self) -> int:
        if getattr(self, "_key", None) is None:
            rank = 0
            try:
                import torch.distributed as dist

                rank = int(dist.get_rank()) if dist.is_available() and dist.is_initialized() else 0
            except Exception:
                rank = 0
            self._key = (self.seed + 0x9E3779B97F4A7C15 * rank) & (2**64 - 1)
        return self._key

    def sample_into(self, out, draw_counter, env=None) -> None:
        """Philox sample into caller-owned tensors ``out = (obs, act, next_obs, dones, rewards)`` with the draw counter read from the
        device tensor ``draw_counter`` (int64, 1 element) at execution time: nothing per-call is baked into the launch, so the call can
        sit inside a CUDA graph (``FusedTD3Update.train(..., graph=True)``).  The caller advances the counter."""
        if self.index_mode != "philox":
            raise ValueError("sample_into needs index_mode='philox' (indices drawn inside the kernel)")
        if env is not None and getattr(env, "norm_params", None) is None:
            raise ValueError("sample_into normalises inside the kernel: env must be a GpuVecNormalize (device statistics); a host-side "
                             "VecNormalize needs sample(), or GpuVecNormalize.from_reference(env)")
        upper_bound = self.buffer_size if self.full else self.pos
        if upper_bound <= 0:
            raise ValueError("cannot sample from an empty replay buffer")
        obs, act, nobs, dones, rew = out
        with self._torch.cuda.device(self._device):
            rc = self._libc.cstr_replay_sample_philox_dev(self._philox_key(), _lib.ptr(draw_counter), self.n_envs, upper_bound, obs.shape[0],
                                                          _lib.ptr(self.records), _lib.ptr(obs), _lib.ptr(act), _lib.ptr(nobs), _lib.ptr(dones),
                                                          _lib.ptr(rew), None, None, self._norm_arg(env), self._stream())
        _lib.check(rc, "cstr_replay_sample_philox_dev")
        self.launches += 1

    def _get_samples(self, batch_inds: np.ndarray, env=None) -> ReplayBufferSamples:
        env_indices = np.random.randint(0, high=self.n_envs, size=(len(batch_inds),))  # buffers.py:309
        return self.gather(batch_inds, env_indices, env=env)

    def gather(self, batch_inds, env_indices, env=None) -> ReplayBufferSamples:
        """Gather the transitions at explicit (row, env) index pairs (host arrays or device int64 tensors)."""
        torch = self._torch
        B = len(batch_inds)
        bi = self._dev(batch_inds, torch.int64, (B,))
        ei = self._dev(env_indices, torch.int64, (B,))
        obs, act, nobs, dones, rew = self._alloc_out(B)
        with torch.cuda.device(self._device):
            rc = self._libc.cstr_replay_sample(self.n_envs, B, _lib.ptr(bi), _lib.ptr(ei), _lib.ptr(self.records), _lib.ptr(obs),
                                               _lib.ptr(act), _lib.ptr(nobs), _lib.ptr(dones), _lib.ptr(rew), self._norm_arg(env), self._stream())
        _lib.check(rc, "cstr_replay_sample")
        self.launches += 1
        return self._finish(obs, act, nobs, dones, rew, env)

    def _norm_arg(self, env):
        """Device VecNormalize statistics are applied inside the gather kernel (None otherwise)."""
        params = getattr(env, "norm_params", None)
        if params is None:
            return None
        self._norm_keepalive = params()  # ctypes struct must outlive the call
        from ctypes import byref

        return byref(self._norm_keepalive)

    def _finish(self, obs, act, nobs, dones, rew, env) -> ReplayBufferSamples:
        if env is not None and getattr(env, "norm_params", None) is None:
            # VecNormalize statistics live on the host in the reference (vec_normalize.py:174-259):
            # normalise there and come back (compat path; the device version is the §8f-4 "next" row)
            torch = self._torch
            obs = torch.as_tensor(env.normalize_obs(obs.cpu().numpy()), device=self._device, dtype=torch.float32)
            nobs = torch.as_tensor(env.normalize_obs(nobs.cpu().numpy()), device=self._device, dtype=torch.float32)
            rew = torch.as_tensor(env.normalize_reward(rew.cpu().numpy()).astype(np.float32), device=self._device)
        return ReplayBufferSamples(obs, act, nobs, dones, rew)

    def to_torch(self, array, copy: bool = True):  # buffers.py:128-140
        torch = self._torch
        if isinstance(array, torch.Tensor):
            return array.to(self._device, torch.float32, copy=copy)
        return torch.tensor(array, device=self._device, dtype=torch.float32)

    # ---- interchange with the reference's pickled ReplayBuffer (save_util.py:339-373) ------------------------
    def to_numpy_arrays(self) -> dict:
        """The reference's attribute layout: contiguous NumPy arrays of the six stores."""
        rec = self.records.cpu().numpy()
        return dict(
            observations=np.ascontiguousarray(rec[..., 0:4]),
            next_observations=np.ascontiguousarray(rec[..., 4:8]),
            actions=np.ascontiguousarray(rec[..., 8:10]),
            rewards=np.ascontiguousarray(rec[..., 10]),
            dones=np.ascontiguousarray(rec[..., 11]),
            timeouts=np.ascontiguousarray(rec[..., 12]),
        )

    def load_numpy_arrays(self, arrays: dict, pos: int, full: bool) -> None:
        torch = self._torch
        T, N = self.buffer_size, self.n_envs
        rec = np.zeros((T, N, _lib.REC_FLOATS), np.float32)
        rec[..., 0:4] = np.asarray(arrays["observations"], np.float32).reshape(T, N, 4)
        rec[..., 4:8] = np.asarray(arrays["next_observations"], np.float32).reshape(T, N, 4)
        rec[..., 8:10] = np.asarray(arrays["actions"], np.float32).reshape(T, N, 2)
        rec[..., 10] = np.asarray(arrays["rewards"], np.float32).reshape(T, N)
        rec[..., 11] = np.asarray(arrays["dones"], np.float32).reshape(T, N)
        rec[..., 12] = np.asarray(arrays["timeouts"], np.float32).reshape(T, N)
        self.records.copy_(torch.as_tensor(rec))
        self.pos, self.full = int(pos), bool(full)

    @classmethod
    def from_reference(cls, ref_buffer, device="auto", **kwargs) -> "GpuReplayBuffer":
        """Build a device buffer from a reference ``ReplayBuffer`` object (e.g. an unpickled BCQ dataset,
        offline_policy_algorithm.py:196-242)."""
        buf = cls(ref_buffer.buffer_size * ref_buffer.n_envs, ref_buffer.observation_space, ref_buffer.action_space, device=device,
                  n_envs=ref_buffer.n_envs, handle_timeout_termination=getattr(ref_buffer, "handle_timeout_termination", True), **kwargs)
        buf.load_numpy_arrays({k: getattr(ref_buffer, k) for k in
                               ("observations", "next_observations", "actions", "rewards", "dones", "timeouts")}, ref_buffer.pos, ref_buffer.full)
        return buf

    def to_reference(self, ref_buffer_class):
        """Materialise a reference ``ReplayBuffer`` (host NumPy) with this buffer's content."""
        ref = ref_buffer_class(self.buffer_size * self.n_envs, self.observation_space, self.action_space, device="cpu", n_envs=self.n_envs,
                               handle_timeout_termination=self.handle_timeout_termination)
        for k, v in self.to_numpy_arrays().items():
            getattr(ref, k)[...] = v
        ref.pos, ref.full = self.pos, self.full
        return ref

    def __reduce__(self):
        # classes made by bind_replay_buffer_class are function-local: pickle names the reference base instead and the bound class is
        # rebuilt on load (so `isinstance(buffer, ReplayBuffer)` in load_replay_buffer, off_policy_algorithm.py:239, still holds)
        base = getattr(type(self), "_bound_base", None)
        return (_unpickle_buffer, (getattr(base, "__module__", None), getattr(base, "__qualname__", None), self.__getstate__()))

    def __getstate__(self):
        state = {k: v for k, v in self.__dict__.items() if k not in ("records", "_torch", "_libc", "_device", "_norm_keepalive", "_key")}
        state.update(self.to_numpy_arrays())
        state["device"] = str(self._device)
        return state

    def __setstate__(self, state):
        torch = _lib.require_cuda()
        arrays = {k: state.pop(k) for k in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts")}
        dev = state.pop("device", "cuda")
        self.__dict__.update(state)
        self._torch = torch
        self._libc = _lib.load()
        self._device = _get_device(dev if torch.cuda.device_count() > (torch.device(dev).index or 0) else "cuda")
        self.records = torch.zeros((self.buffer_size, self.n_envs, _lib.REC_FLOATS), dtype=torch.float32, device=self._device)
        self.load_numpy_arrays(arrays, self.pos, self.full)


_BOUND_CLASSES: dict = {}


def _unpickle_buffer(base_module: Optional[str], base_qualname: Optional[str], state: dict) -> GpuReplayBuffer:
    cls = GpuReplayBuffer
    if base_module and base_qualname:
        try:
            import importlib

            base = importlib.import_module(base_module)
            for part in base_qualname.split("."):
                base = getattr(base, part)
            cls = bind_replay_buffer_class(base)
        except Exception:  # the reference is not importable here: the plain class carries the same data
            cls = GpuReplayBuffer
    obj = cls.__new__(cls)
    obj.__setstate__(state)
    return obj


def bind_replay_buffer_class(replay_buffer_base: type) -> type:
    """``class GpuReplayBuffer(GpuReplayBuffer, <reference ReplayBuffer>)`` for ``isinstance`` checks in
    the unchanged reference (e.g. ``isinstance(self.replay_buffer, ReplayBuffer)`` when loading).  One class per base (cached), and
    instances pickle through ``_unpickle_buffer`` (``save_replay_buffer`` / ``load_replay_buffer``, save_util.py:339-373)."""
    if replay_buffer_base in _BOUND_CLASSES:
        return _BOUND_CLASSES[replay_buffer_base]

    class BoundGpuReplayBuffer(GpuReplayBuffer, replay_buffer_base):  # type: ignore[misc, valid-type]
        _bound_base = replay_buffer_base

    BoundGpuReplayBuffer.__name__ = "GpuReplayBuffer"
    BoundGpuReplayBuffer.__qualname__ = "GpuReplayBuffer"
    _BOUND_CLASSES[replay_buffer_base] = BoundGpuReplayBuffer
    return BoundGpuReplayBuffer
