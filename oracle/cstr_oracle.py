"""CPU ORACLE (TEST INFRASTRUCTURE) — NumPy restatement of the reference CSTR hot path.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
(``pytorch-rl-enhancedstablebaselines_b200``) never does and fails loudly without its CUDA library.

Parity status: PINNED.  ``tests/test_oracle_vs_reference.py`` runs the unmodified reference
(``/root/reference/twoseriescstr.py``, ``core/common/buffers.py``) in the build container and
requires bit-equality with this restatement; ``oracle/make_golden.py`` froze reference outputs
into ``tests/golden/*.npz`` which are re-checked wherever the reference tree is absent.

What is restated (reference file:line):
  * ``TwoSeriesCSTREnv.step``            twoseriescstr.py:394-454   -> :func:`step_f32`, :func:`step_f64`
  * ``TwoSeriesCSTREnv._dynamics``       twoseriescstr.py:456-503   -> :func:`dynamics`
  * ``TwoSeriesCSTREnv.compute_reward``  twoseriescstr.py:271-392   -> :func:`reward_terms`
  * ``_normalize_state/_denormalize_*``  twoseriescstr.py:129-150   -> :func:`normalize_state` ...
  * ``reset`` / ``generate_initial_state`` twoseriescstr.py:226-269,167-224 -> :func:`initial_state_from_uniforms`
  * ``DummyVecEnv.step_wait``            core/common/vec_env/dummy_vec_env.py:56-73 -> :class:`VecOracle`
  * ``ReplayBuffer.add/sample``          core/common/buffers.py:247-325 -> :class:`ReplayOracle`
  * ``_sample_action`` scale/noise/clip/unscale  core/common/off_policy_algorithm.py:398-406,
    core/common/policies.py:388-413     -> :func:`sample_action_maps`
  * TD3 ``Actor.forward``                core/td3/policies.py:75-78 -> :func:`actor_forward`

Arithmetic model: under NumPy >= 2 (NEP 50) the reference computes the
whole step in float32; sub-expressions made only of Python-float class constants are folded in
float64 first and rounded to float32 when they first meet a float32 operand.  ``dtype=np.float64``
reproduces what the reference's ``_dynamics`` does when fed float64 arrays (bounds are the float32
constants widened), plus the same affine maps in float64.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple, Optional, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------
# constants (twoseriescstr.py:37-61)
# ----------------------------------------------------------------------------------------------
Q, V1, V2 = 50, 100, 100
Cf, Tf, Tcf = 0.5, 320, 370
k0, E, R = 7.2e10, 8.314e4, 8.314
delta_H = -6.78e4
rou, rou_c, c_p, c_pc = 1000, 1000, 0.239, 0.239
U, A1, A2 = 6.6e5, 8.958, 8.958
DT = 0.1
MAX_STEPS = 400

RAW_STATE_LOW = np.array([0.0, 273.15, 0.0, 273.15], dtype=np.float32)
RAW_STATE_HIGH = np.array([0.7, 400.0, 0.7, 400.0], dtype=np.float32)
RAW_ACTION_LOW = np.array([30.0, 30.0], dtype=np.float32)
RAW_ACTION_HIGH = np.array([250.0, 250.0], dtype=np.float32)

DEFAULT_TARGET = 0.20
MIN_CONC, MAX_CONC = 0.05, 0.45
NAN_REWARD = -10.0  # twoseriescstr.py:418


def folded_constants(dtype) -> dict:
    """Python-float constant folding of twoseriescstr.py:479-491, then cast to the working dtype."""
    t = np.dtype(dtype).type
    return dict(
        QV1=t(Q / V1),
        QV2=t(Q / V2),
        Cf=t(Cf),
        Tf=t(Tf),
        Tcf=t(Tcf),
        k0=t(k0),
        nE=t(-E),
        R=t(R),
        HK=t(-delta_H * k0),
        RC=t(rou * c_p),
        KC1=t((rou_c * c_pc) / (rou * c_p * V1)),
        KC2=t((rou_c * c_pc) / (rou * c_p * V2)),
        nUA1=t(-(U * A1)),
        nUA2=t(-(U * A2)),
        rou_c=t(rou_c),
        c_pc=t(c_pc),
        dt=t(DT),
    )


# ----------------------------------------------------------------------------------------------
# affine maps (twoseriescstr.py:129-150).  The reference always ends them with .astype(float32);
# in float64 mode (the "same scheme in fp64" oracle of SURVEY 8c) they stay float64.
# ----------------------------------------------------------------------------------------------
def _bounds(dtype):
    return (
        RAW_STATE_LOW.astype(dtype),
        RAW_STATE_HIGH.astype(dtype),
        RAW_ACTION_LOW.astype(dtype),
        RAW_ACTION_HIGH.astype(dtype),
    )


def normalize_state(raw: np.ndarray, dtype=np.float32) -> np.ndarray:
    slo, shi, _, _ = _bounds(raw.dtype if raw.dtype == np.float64 else dtype)
    t = raw.dtype.type
    return (t(2.0) * (raw - slo) / (shi - slo) - t(1.0)).astype(dtype)


def denormalize_state(norm: np.ndarray, dtype=np.float32) -> np.ndarray:
    slo, shi, _, _ = _bounds(norm.dtype)
    t = norm.dtype.type
    return (slo + (norm + t(1.0)) * (shi - slo) / t(2.0)).astype(dtype)


def denormalize_action(norm: np.ndarray, dtype=np.float32) -> np.ndarray:
    _, _, alo, ahi = _bounds(norm.dtype)
    t = norm.dtype.type
    return (alo + (norm + t(1.0)) * (ahi - alo) / t(2.0)).astype(dtype)


# ----------------------------------------------------------------------------------------------
# _dynamics (twoseriescstr.py:456-503), vectorised over the leading axis
# ----------------------------------------------------------------------------------------------
def dynamics(state: np.ndarray, action: np.ndarray, exp=np.exp) -> np.ndarray:
    """One explicit-Euler update of the four ODEs on RAW state (N,4) / RAW action (N,2).

    Works in the dtype of ``state`` (float32 or float64), with the association of App. A.
    NaN rows are the caller's business (the reference raises, twoseriescstr.py:466-467).
    """
    dt_ = state.dtype
    t = dt_.type
    c = folded_constants(dt_)
    slo, shi, _, _ = _bounds(dt_)
    C1, T1, C2, T2 = (state[:, i] for i in range(4))
    F1, F2 = action[:, 0], action[:, 1]

    # :470-473   T = max(T, 273.15) (weak Python float -> same value as the float32 bound)
    T1 = np.maximum(T1, t(np.float32(273.15)) if dt_ == np.float32 else t(273.15))
    T2 = np.maximum(T2, t(np.float32(273.15)) if dt_ == np.float32 else t(273.15))
    F1 = np.clip(F1, t(1e-5), t(1e5))
    F2 = np.clip(F2, t(1e-5), t(1e5))

    def safe_exp(x):  # :476-477
        return exp(np.clip(x, t(-100), t(100)))

    k1 = safe_exp(c["nE"] / (c["R"] * T1))
    k2 = safe_exp(c["nE"] / (c["R"] * T2))
    c1 = safe_exp(c["nUA1"] / ((F1 * c["rou_c"]) * c["c_pc"]))
    c2 = safe_exp(c["nUA2"] / ((F2 * c["rou_c"]) * c["c_pc"]))

    dC1 = c["QV1"] * (c["Cf"] - C1) - (c["k0"] * C1) * k1
    dT1 = (c["QV1"] * (c["Tf"] - T1) + ((c["HK"] * C1) / c["RC"]) * k1) + ((c["KC1"] * F1) * (t(1) - c1)) * (
        c["Tcf"] - T1
    )
    dC2 = c["QV2"] * (C1 - C2) - (c["k0"] * C2) * k2
    dT2 = (c["QV2"] * (T1 - T2) + ((c["HK"] * C2) / c["RC"]) * k2) + ((c["KC2"] * F2) * (t(1) - c2)) * (
        c["Tcf"] - T2
    )

    new = np.stack(
        [C1 + dC1 * c["dt"], T1 + dT1 * c["dt"], C2 + dC2 * c["dt"], T2 + dT2 * c["dt"]],
        axis=1,
    )
    return np.clip(new, slo, shi)


# ----------------------------------------------------------------------------------------------
# compute_reward (twoseriescstr.py:271-392): the two terms with non-zero weight
# ----------------------------------------------------------------------------------------------
_libm_powf = None


def _powf2(n: np.ndarray) -> np.ndarray:
    """``normalized_error ** 2`` (:291) on a NumPy float32 *scalar* is a libm ``powf(x, 2)`` call.
    It is NOT always the correctly rounded square (differs from x*x by 1 ulp in ~0.03 % of inputs,
    probed) and NumPy's array pow loop is a third, less accurate, algorithm — so the restatement
    calls libm's powf element by element.  float64: pow(x, 2) is exact-rounded == x*x."""
    global _libm_powf
    if n.dtype != np.float32:
        return n * n
    if _libm_powf is None:
        import ctypes
        import ctypes.util

        lib = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        lib.powf.restype = ctypes.c_float
        lib.powf.argtypes = [ctypes.c_float, ctypes.c_float]
        _libm_powf = np.frompyfunc(lambda v: lib.powf(v, 2.0), 1, 1)
    return _libm_powf(n).astype(np.float32)


def square_mul(n: np.ndarray) -> np.ndarray:
    """x*x — the correctly rounded square the CUDA kernels (and the C oracle's ``mul`` mode) use."""
    return n * n


def reward_terms(obs: np.ndarray, target=DEFAULT_TARGET, square=_powf2) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(reward, concentration_reward, temp_penalty) from the NORMALISED new state (Q12)."""
    dt_ = obs.dtype
    t = dt_.type
    raw = denormalize_state(obs, dt_)
    T1, C2, T2 = raw[:, 1], raw[:, 2], raw[:, 3]
    err = np.abs(C2 - t(target))  # :288
    n = err / t(MAX_CONC - MIN_CONC)  # :290
    conc = t(-5.0) * square(n) - t(2.0) * n  # :291
    tp = np.zeros_like(C2)
    for T in (T1, T2):  # :333-341
        lo = T < t(280)
        hi = (~lo) & (T > t(350))
        with np.errstate(invalid="ignore"):
            tp = np.where(lo, tp - t(0.2) * ((t(280) - T) / t(280)), tp)
            tp = np.where(hi, tp - t(0.5) * ((T - t(350)) / t(350)), tp)
    reward = t(1.0) * conc + t(0.5) * tp  # :369-377 (the five 0.0-weighted terms are value-neutral)
    return reward, conc, tp


# ----------------------------------------------------------------------------------------------
# step (twoseriescstr.py:394-454)
# ----------------------------------------------------------------------------------------------
class StepOut(NamedTuple):
    obs: np.ndarray  # (N,4) new normalised state
    reward: np.ndarray  # (N,)
    truncated: np.ndarray  # (N,) bool
    step_count: np.ndarray  # (N,) int32, already incremented
    nan_row: np.ndarray  # (N,) bool — the reference's "Dynamics calculation error" path (Q4)


def step(state: np.ndarray, action: np.ndarray, step_count: np.ndarray, target=DEFAULT_TARGET, exp=np.exp,
         square=_powf2) -> StepOut:
    """One control interval for N independent reactors, in the dtype of ``state``."""
    dt_ = state.dtype
    t = dt_.type
    action = np.asarray(action, dtype=dt_)
    slo, shi, _, _ = _bounds(dt_)
    step_count = step_count.astype(np.int32) + 1  # :396
    na = np.clip(action, t(-1), t(1))  # :399
    raw_a = denormalize_action(na, dt_)  # :400
    raw_s = np.clip(denormalize_state(state, dt_), slo, shi)  # :404-410
    nan_row = np.isnan(raw_s).any(axis=1) | np.isnan(raw_a).any(axis=1)  # :466
    with np.errstate(invalid="ignore", over="ignore"):
        new_raw = dynamics(raw_s, np.where(nan_row[:, None], t(30), raw_a), exp=exp)  # :414
        new_raw = np.clip(new_raw, slo, shi)  # :424-428
        obs = normalize_state(new_raw, dt_)  # :429
        reward, _, _ = reward_terms(obs, target, square=square)  # :432
    truncated = step_count >= MAX_STEPS  # :438
    # NaN path (:415-421): state unchanged, reward -10, terminated False, truncated True
    obs = np.where(nan_row[:, None], state, obs)
    reward = np.where(nan_row, t(NAN_REWARD), reward).astype(dt_)
    truncated = np.where(nan_row, True, truncated)
    return StepOut(obs, reward, truncated, step_count, nan_row)


def step_f32(state, action, step_count, **kw) -> StepOut:
    return step(np.asarray(state, np.float32), np.asarray(action, np.float32), step_count, **kw)


def step_f64(state, action, step_count, **kw) -> StepOut:
    """fp64 oracle of SURVEY 8c: fp64 affine maps + the reference's dtype-generic ``_dynamics``.
    ``square`` defaults to x*x in fp64 (libm pow(x,2) is exact-rounded there)."""
    kw.setdefault("square", lambda n: n * n)
    return step(np.asarray(state, np.float64), np.asarray(action, np.float64), step_count, **kw)


# ----------------------------------------------------------------------------------------------
# reset (twoseriescstr.py:226-269, 167-224).  All of it is float64 until the final astype(float32).
# ----------------------------------------------------------------------------------------------
def initial_state_from_uniforms(u: np.ndarray) -> np.ndarray:
    """init_mode="random": RAW float64 initial state (N,4) from 8 unit-uniform doubles per env,
    consumed in the reference's draw order (4 scalar uniforms, then one size-4 noise draw)."""
    u = np.asarray(u, np.float64)
    lo = np.array([0.05, 280.0, 0.05, 280.0])
    hi = np.array([0.45, 380.0, 0.45 * 0.8, 380.0])
    s = lo + (hi - lo) * u[:, :4]  # Generator.uniform = low + (high-low)*next_double  (:187-199)
    s = s + (-0.05 + (0.05 - (-0.05)) * u[:, 4:8])  # :202-207
    swap_t = s[:, 1] < s[:, 3]  # :211-212
    s[swap_t, 1], s[swap_t, 3] = s[swap_t, 3], s[swap_t, 1].copy()
    swap_c = s[:, 0] < s[:, 2]  # :214-215
    s[swap_c, 0], s[swap_c, 2] = s[swap_c, 2], s[swap_c, 0].copy()
    return np.clip(s, RAW_STATE_LOW, RAW_STATE_HIGH)  # :218-222 (float64 result)


def static_state_from_uniforms(init_state: np.ndarray, u: np.ndarray) -> np.ndarray:
    """init_mode="static": ``init_state += uniform([-.05,-10,-.05,-10],[.05,10,.05,10])`` IN PLACE
    (:245-253, quirk Q2: the base state random-walks across episodes; no clip)."""
    lo = np.array([-0.05, -10.0, -0.05, -10.0])
    hi = np.array([0.05, 10.0, 0.05, 10.0])
    init_state += lo + (hi - lo) * np.asarray(u, np.float64)[:, :4]
    return init_state


def obs_from_raw_f64(raw: np.ndarray, dtype=np.float32) -> np.ndarray:
    """``_normalize_state`` on the float64 reset state (:267, :131-132)."""
    slo, shi = RAW_STATE_LOW, RAW_STATE_HIGH  # float32 arrays: (raw64 - lo32) promotes to float64
    return (2.0 * (raw - slo) / (shi - slo) - 1.0).astype(dtype)


def pcg64_reset_uniforms(seed: int, n_resets: int = 1) -> np.ndarray:
    """The 8 doubles per reset the reference draws from Generator(PCG64(SeedSequence(seed)))."""
    g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    return g.random((n_resets, 8))


# ----------------------------------------------------------------------------------------------
# Philox4x32-10 mirror (device RNG of the product; counter layout documented in DESIGN.md)
# ----------------------------------------------------------------------------------------------
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(counter: np.ndarray, key: np.ndarray, rounds: int = 10) -> np.ndarray:
    """counter (N,4) uint32, key (N,2) or (2,) uint32 -> (N,4) uint32."""
    c = np.array(counter, dtype=np.uint32, copy=True).reshape(-1, 4)
    k = np.broadcast_to(np.asarray(key, np.uint32), (c.shape[0], 2)).copy()
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(rounds):
            p0 = _PH_M0 * c[:, 0].astype(np.uint64)
            p1 = _PH_M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k[:, 0], lo1, hi0 ^ c[:, 3] ^ k[:, 1], lo0], axis=1)
            k = np.stack([k[:, 0] + _PH_W0, k[:, 1] + _PH_W1], axis=1)
    return c


def u32_to_unit_f32(x: np.ndarray) -> np.ndarray:
    """(x >> 8) * 2^-24 in [0,1) — the product's uint32 -> float32 uniform."""
    return ((x >> np.uint32(8)).astype(np.float32)) * np.float32(2.0**-24)


def u32x2_to_unit_f64(hi: np.ndarray, lo: np.ndarray) -> np.ndarray:
    """53-bit double in [0,1): ((hi >> 5) * 2^26 + (lo >> 6)) * 2^-53 (same recipe as NumPy's next_double
    on two 32-bit words)."""
    return ((hi >> np.uint32(5)).astype(np.float64) * 67108864.0 + (lo >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0


# ----------------------------------------------------------------------------------------------
# DummyVecEnv.step_wait semantics (dummy_vec_env.py:56-73) over the vectorised step
# ----------------------------------------------------------------------------------------------
@dataclass
class VecStep:
    obs: np.ndarray  # (N,4) post-reset observation on done rows
    reward: np.ndarray
    done: np.ndarray  # terminated or truncated
    timeout: np.ndarray  # info["TimeLimit.truncated"] = truncated and not terminated
    terminal_obs: np.ndarray  # (N,4) the pre-reset observation (== obs on non-done rows)


class VecOracle:
    """N reactors with auto-reset; the reset states are injected by the caller (``reset_fn``)."""

    def __init__(self, state: np.ndarray, dtype=np.float32, target=DEFAULT_TARGET):
        self.dtype = np.dtype(dtype)
        self.state = np.array(state, dtype=self.dtype)
        self.step_count = np.zeros(len(self.state), np.int32)
        self.target = target

    def step(self, actions: np.ndarray, reset_fn=None, **kw) -> VecStep:
        out = step(self.state, np.asarray(actions, self.dtype), self.step_count, self.target, **kw)
        done = out.truncated.copy()  # terminated is always False (:435)
        terminal = out.obs.copy()
        self.state = out.obs.copy()
        self.step_count = out.step_count.copy()
        if done.any():
            idx = np.nonzero(done)[0]
            if reset_fn is not None:
                self.state[idx] = reset_fn(idx).astype(self.dtype)
            self.step_count[idx] = 0
        return VecStep(self.state.copy(), out.reward, done, done.copy(), terminal)


# ----------------------------------------------------------------------------------------------
# ReplayBuffer (core/common/buffers.py:185-325)
# ----------------------------------------------------------------------------------------------
class ReplayOracle:
    def __init__(self, buffer_size: int, n_envs: int = 1, obs_dim: int = 4, action_dim: int = 2):
        self.buffer_size = max(buffer_size // n_envs, 1)  # :198
        self.n_envs = n_envs
        T, N = self.buffer_size, n_envs
        self.observations = np.zeros((T, N, obs_dim), np.float32)
        self.next_observations = np.zeros((T, N, obs_dim), np.float32)
        self.actions = np.zeros((T, N, action_dim), np.float32)
        self.rewards = np.zeros((T, N), np.float32)
        self.dones = np.zeros((T, N), np.float32)
        self.timeouts = np.zeros((T, N), np.float32)
        self.pos, self.full = 0, False

    def add(self, obs, next_obs, action, reward, done, timeout) -> None:  # :247-283
        p = self.pos
        self.observations[p] = obs
        self.next_observations[p] = next_obs
        self.actions[p] = np.asarray(action).reshape(self.n_envs, -1)
        self.rewards[p] = reward
        self.dones[p] = done
        self.timeouts[p] = timeout
        self.pos += 1
        if self.pos == self.buffer_size:
            self.full, self.pos = True, 0

    def draw_indices(self, batch_size: int) -> Tuple[np.ndarray, np.ndarray]:
        """The two GLOBAL-RNG draws of :114 and :309, in that order."""
        upper = self.buffer_size if self.full else self.pos
        batch_inds = np.random.randint(0, upper, size=batch_size)
        env_inds = np.random.randint(0, high=self.n_envs, size=(batch_size,))
        return batch_inds, env_inds

    def gather(self, b: np.ndarray, e: np.ndarray):  # :316-324
        return (
            self.observations[b, e, :],
            self.actions[b, e, :],
            self.next_observations[b, e, :],
            (self.dones[b, e] * (1 - self.timeouts[b, e])).reshape(-1, 1),
            self.rewards[b, e].reshape(-1, 1),
        )

    def sample(self, batch_size: int):
        return self.gather(*self.draw_indices(batch_size))


# ----------------------------------------------------------------------------------------------
# VecNormalize (core/common/vec_env/vec_normalize.py:174-298, core/common/running_mean_std.py:4-56)
# ----------------------------------------------------------------------------------------------
class RunningMeanStdOracle:
    """running_mean_std.py:4-56 — ``exact=False`` keeps the reference's float32 batch moments (np.mean/np.var of the
    float32 array), ``exact=True`` takes them in float64 as the CUDA reduction does."""

    def __init__(self, shape=(), epsilon: float = 1e-4, exact: bool = False):
        self.mean, self.var, self.count, self.exact = np.zeros(shape, np.float64), np.ones(shape, np.float64), epsilon, exact

    def update(self, arr: np.ndarray) -> None:  # :35-39
        a = np.asarray(arr, np.float64) if self.exact else np.asarray(arr)
        self.update_from_moments(np.mean(a, axis=0), np.var(a, axis=0), a.shape[0])

    def update_from_moments(self, bmean, bvar, bcount) -> None:  # :41-56
        delta = bmean - self.mean
        tot = self.count + bcount
        new_mean = self.mean + delta * bcount / tot
        m2 = self.var * self.count + bvar * bcount + np.square(delta) * self.count * bcount / (self.count + bcount)
        self.mean, self.var, self.count = new_mean, m2 / (self.count + bcount), bcount + self.count


class VecNormalizeOracle:
    """The statistics/normalisation half of VecNormalize, fed with the raw (obs, reward, done) of each step."""

    def __init__(self, n_envs: int, gamma=0.99, epsilon=1e-8, clip_obs=10.0, clip_reward=10.0, training=True, norm_obs=True,
                 norm_reward=True, exact: bool = False):
        self.obs_rms, self.ret_rms = RunningMeanStdOracle((4,), exact=exact), RunningMeanStdOracle((), exact=exact)
        self.returns = np.zeros(n_envs)
        self.gamma, self.epsilon, self.clip_obs, self.clip_reward = gamma, epsilon, clip_obs, clip_reward
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward

    def normalize_obs(self, obs):  # :225-246
        if not self.norm_obs:
            return np.array(obs, copy=True)
        return np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs, self.clip_obs).astype(np.float32)

    def normalize_reward(self, reward):  # :248-259
        if self.norm_reward:
            reward = np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        return np.asarray(reward).astype(np.float32)

    def reset(self, obs):  # :287-298
        self.returns = np.zeros_like(self.returns)
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        return self.normalize_obs(obs)

    def step(self, obs, rewards, dones):  # :174-209
        if self.training and self.norm_obs:
            self.obs_rms.update(obs)
        nobs = self.normalize_obs(obs)
        if self.training:
            self.returns = self.returns * self.gamma + rewards  # :221
            self.ret_rms.update(self.returns)
        nrew = self.normalize_reward(rewards)
        self.returns[np.asarray(dones, bool)] = 0
        return nobs, nrew


# ----------------------------------------------------------------------------------------------
# action plumbing and the TD3 actor (off_policy_algorithm.py:398-406, policies.py:388-413,
# td3/policies.py:75-78 + torch_layers.py:110-183)
# ----------------------------------------------------------------------------------------------
def actor_forward(obs: np.ndarray, weights) -> np.ndarray:
    """tanh(W3 relu(W2 relu(W1 x + b1) + b2) + b3); ``weights`` = [(W,b),...] in torch Linear layout
    (out,in).  Accumulates in float64 and is compared with a tolerance (the reference's sgemm
    summation order is not defined)."""
    h = np.asarray(obs, np.float64)
    for i, (W, b) in enumerate(weights):
        h = h @ np.asarray(W, np.float64).T + np.asarray(b, np.float64)
        if i < len(weights) - 1:
            h = np.maximum(h, 0.0)
    return np.tanh(h)


def sac_actor_forward(obs: np.ndarray, weights, eps: np.ndarray) -> np.ndarray:
    """SAC squashed-Gaussian actor (core/sac/policies.py:151-168, distributions.py:207-260):
    trunk = relu(W2 relu(W1 x + b1) + b2); mu, log_std heads; a = tanh(mu + exp(clamp(log_std,-20,2)) * eps).
    ``weights`` = [(W1,b1),(W2,b2),(W3,b3)] with W3 (4,H2) = rows [mu_0, mu_1, log_std_0, log_std_1]."""
    h = np.asarray(obs, np.float64)
    for W, b in weights[:2]:
        h = np.maximum(h @ np.asarray(W, np.float64).T + np.asarray(b, np.float64), 0.0)
    W3, b3 = weights[2]
    head = h @ np.asarray(W3, np.float64).T + np.asarray(b3, np.float64)
    mu, log_std = head[:, :2], np.clip(head[:, 2:], -20.0, 2.0)
    return np.tanh(mu + np.exp(log_std) * np.asarray(eps, np.float64))


def sample_action_maps(mu: np.ndarray, noise: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """float32 chain for a Box(-1,1) action space: predict's unscale, _sample_action's scale,
    + noise, clip, unscale.  Returns (env_action, buffer_action)."""
    f = np.float32
    mu = np.asarray(mu, f)
    low, high = f(-1.0), f(1.0)
    u = low + (f(0.5) * (mu + f(1.0)) * (high - low))  # policies.py:402-413 (predict :375, squash_output)
    s = f(2.0) * ((u - low) / (high - low)) - f(1.0)  # policies.py:388-400
    s = np.clip(s + np.asarray(noise, f), f(-1), f(1))  # off_policy_algorithm.py:402
    a = low + (f(0.5) * (s + f(1.0)) * (high - low))  # :405
    return a.astype(f), s.astype(f)
