"""Build + ctypes loader for the C oracle (TEST INFRASTRUCTURE; see cstr_oracle.c header)."""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "cstr_oracle.c")
OUT_DIR = os.path.join(_HERE, "_build")
OUT = os.path.join(OUT_DIR, "libcstr_oracle.so")
BASE_FLAGS = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-Wall", "-Wextra"]

EXP_LIBM, EXP_SHARED = 0, 1
SQ_POWF, SQ_MUL = 0, 1


def build(force: bool = False) -> str:
    """Compile cstr_oracle.c -> oracle/_build/libcstr_oracle.so (OpenMP when the compiler has it)."""
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    # the image's $CC wrapper (/opt/gcc/bin/gcc) lacks libgomp.spec -> prefer the system gcc
    candidates = [c for c in ("/usr/bin/gcc", shutil.which("gcc"), os.environ.get("CC")) if c]
    last = None
    for omp in (["-fopenmp"], []):
        for cc in candidates:
            cmd = [cc, *BASE_FLAGS, *omp, "-o", OUT, SRC, "-lm"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode == 0:
                return OUT
            last = res.stderr
    raise RuntimeError(f"could not build the C oracle: {last}")


_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    lib = ctypes.CDLL(build())
    c = ctypes
    P = c.c_void_p
    lib.cstr_expf_shared_array.argtypes = [P, P, c.c_int64]
    lib.cstr_powf2_array.argtypes = [P, P, c.c_int64]
    lib.cstr_philox_array.argtypes = [P, P, P, c.c_int64]
    lib.cstr_oracle_reset_uniforms.argtypes = [c.c_uint64, c.c_int64, c.c_int64, P, P]
    lib.cstr_oracle_step_f32.argtypes = [P, P, P, P, P, P, c.c_int64, c.c_float, c.c_int, c.c_int]
    lib.cstr_oracle_step_f64.argtypes = [P, P, P, P, P, P, c.c_int64, c.c_double]
    lib.cstr_oracle_reset_f32.argtypes = [P, P, P, P, P, c.c_int64, c.c_int64, c.c_uint64, c.c_int]
    lib.cstr_oracle_tape_f32.argtypes = [P, P, P, P, P, c.c_int64, c.c_int64, c.c_int64, c.c_uint64, c.c_int,
                                         c.c_float, c.c_int, c.c_int, P, P, P, P, P]
    lib.cstr_oracle_num_threads.restype = c.c_int
    lib.cstr_oracle_set_threads.argtypes = [c.c_int]
    lib.cstr_oracle_set_threads.restype = None
    for name in ("cstr_expf_shared_array", "cstr_powf2_array", "cstr_philox_array", "cstr_oracle_reset_uniforms",
                 "cstr_oracle_step_f32", "cstr_oracle_step_f64", "cstr_oracle_reset_f32", "cstr_oracle_tape_f32"):
        getattr(lib, name).restype = None
    _lib = lib
    return lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


# ---- thin numpy wrappers -------------------------------------------------------------------------
def expf_shared(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    load().cstr_expf_shared_array(_p(x), _p(y), x.size)
    return y


def powf2(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    load().cstr_powf2_array(_p(x), _p(y), x.size)
    return y


def philox(counter: np.ndarray, key) -> np.ndarray:
    ctr = np.ascontiguousarray(counter, np.uint32).reshape(-1, 4)
    k = np.ascontiguousarray(key, np.uint32).reshape(2)
    out = np.empty_like(ctr)
    load().cstr_philox_array(_p(ctr), _p(k), _p(out), ctr.shape[0])
    return out


def reset_uniforms(seed: int, env0: int, episode: np.ndarray) -> np.ndarray:
    ep = np.ascontiguousarray(episode, np.int32)
    u = np.empty((ep.size, 8), np.float64)
    load().cstr_oracle_reset_uniforms(seed, env0, ep.size, _p(ep), _p(u))
    return u


def step_f32(state, action, step_count, target=0.2, exp_mode=EXP_LIBM, sq_mode=SQ_POWF):
    """Returns (obs, reward, truncated, step_count, nan_row); inputs are not modified."""
    s = np.array(state, np.float32, order="C")
    a = np.ascontiguousarray(action, np.float32)
    sc = np.array(step_count, np.int32, order="C")
    n = s.shape[0]
    r = np.empty(n, np.float32)
    tr = np.empty(n, np.uint8)
    bad = np.empty(n, np.uint8)
    load().cstr_oracle_step_f32(_p(s), _p(a), _p(sc), _p(r), _p(tr), _p(bad), n, target, exp_mode, sq_mode)
    return s, r, tr.astype(bool), sc, bad.astype(bool)


def step_f64(state, action, step_count, target=0.2):
    s = np.array(state, np.float64, order="C")
    a = np.ascontiguousarray(action, np.float64)
    sc = np.array(step_count, np.int32, order="C")
    n = s.shape[0]
    r = np.empty(n, np.float64)
    tr = np.empty(n, np.uint8)
    bad = np.empty(n, np.uint8)
    load().cstr_oracle_step_f64(_p(s), _p(a), _p(sc), _p(r), _p(tr), _p(bad), n, target)
    return s, r, tr.astype(bool), sc, bad.astype(bool)


def reset_f32(n, env0=0, seed=0, init_mode=0, episode=None, static_base=None, mask=None, state=None, step_count=None):
    """Philox reset of the masked envs. Returns (state, step_count, episode, static_base)."""
    state = np.zeros((n, 4), np.float32) if state is None else np.array(state, np.float32, order="C")
    step_count = np.zeros(n, np.int32) if step_count is None else np.array(step_count, np.int32, order="C")
    episode = np.zeros(n, np.int32) if episode is None else np.array(episode, np.int32, order="C")
    if static_base is None:
        static_base = np.tile(np.array([0.45, 310.0, 0.25, 290.0]), (n, 1))
    static_base = np.array(static_base, np.float64, order="C")
    m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
    load().cstr_oracle_reset_f32(_p(state), _p(step_count), _p(episode), _p(static_base), _p(m), n, env0, seed, init_mode)
    return state, step_count, episode, static_base


def tape_f32(state, step_count, episode, actions, env0=0, seed=0, init_mode=0, static_base=None, target=0.2,
             exp_mode=EXP_LIBM, sq_mode=SQ_POWF, want_rewards=True, want_dones=True, want_obs=False, want_term=False):
    """T-step tape with auto-reset.  Returns dict(state, step_count, episode, rewards, dones, obs, term, reward_sum)."""
    s = np.array(state, np.float32, order="C")
    sc = np.array(step_count, np.int32, order="C")
    ep = np.array(episode, np.int32, order="C")
    a = np.ascontiguousarray(actions, np.float32)
    T, n = a.shape[0], a.shape[1]
    if static_base is None:
        static_base = np.tile(np.array([0.45, 310.0, 0.25, 290.0]), (n, 1))
    sb = np.array(static_base, np.float64, order="C")
    rewards = np.empty((T, n), np.float32) if want_rewards else None
    dones = np.empty((T, n), np.uint8) if want_dones else None
    obs = np.empty((T, n, 4), np.float32) if want_obs else None
    term = np.empty((T, n, 4), np.float32) if want_term else None
    rsum = np.zeros(1, np.float64)
    load().cstr_oracle_tape_f32(_p(s), _p(sc), _p(ep), _p(sb), _p(a), T, n, env0, seed, init_mode, target, exp_mode,
                                sq_mode, _p(rewards), _p(dones), _p(obs), _p(term), _p(rsum))
    return dict(state=s, step_count=sc, episode=ep, static_base=sb, rewards=rewards,
                dones=None if dones is None else dones.astype(bool), obs=obs, term=term, reward_sum=float(rsum[0]))


def num_threads() -> int:
    return int(load().cstr_oracle_num_threads())


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().cstr_oracle_set_threads(n)
    return num_threads()


if __name__ == "__main__":
    print(build(force=True))
