"""Helpers for the -m gpu tests: call the C ABI directly on torch device tensors."""
import importlib
from ctypes import byref

import numpy as np
import torch

PKG = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
L = PKG._lib


def params(seed=0, env_offset=0, target=0.2, init_mode=0):
    return L.EnvParams(seed=seed, env_offset=env_offset, target_c2=target, max_steps=400, init_mode=init_mode)


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x), device="cuda")
    return t if dtype is None else t.to(dtype)


def raw_step_f32(state, action, step_count, mode=0, auto_reset=0, seed=0, env_offset=0, episode=None, target=0.2, init_mode=0,
                 static_base=None):
    """One cstr_vec_step_f32 call. Returns dict of numpy arrays (state after, terminal, reward, done, step_count, episode)."""
    lib = L.load()
    n = len(state)
    s = dev(np.asarray(state, np.float32))
    a = dev(np.asarray(action, np.float32))
    sc = dev(np.asarray(step_count, np.int32))
    ep = dev(np.zeros(n, np.int32) if episode is None else np.asarray(episode, np.int32))
    sb = None if static_base is None else dev(np.asarray(static_base, np.float64))
    term = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    rew = torch.empty(n, dtype=torch.float32, device="cuda")
    done = torch.empty(n, dtype=torch.uint8, device="cuda")
    to = torch.empty(n, dtype=torch.uint8, device="cuda")
    p = params(seed, env_offset, target, init_mode)
    rc = lib.cstr_vec_step_f32(byref(p), n, mode, auto_reset, L.ptr(a), L.ptr(s), L.ptr(sc), L.ptr(ep), L.ptr(sb), L.ptr(term), L.ptr(rew),
                               L.ptr(done), L.ptr(to), None, None, None, torch.cuda.current_stream().cuda_stream)
    L.check(rc, "cstr_vec_step_f32")
    torch.cuda.synchronize()
    return dict(state=s.cpu().numpy(), terminal=term.cpu().numpy(), reward=rew.cpu().numpy(), done=done.cpu().numpy().astype(bool),
                timeout=to.cpu().numpy().astype(bool), step_count=sc.cpu().numpy(), episode=ep.cpu().numpy(),
                static_base=None if sb is None else sb.cpu().numpy())


def raw_step_f64(state, action, step_count, auto_reset=0, seed=0, env_offset=0, target=0.2):
    lib = L.load()
    n = len(state)
    s = dev(np.asarray(state, np.float64))
    a = dev(np.asarray(action, np.float64))
    sc = dev(np.asarray(step_count, np.int32))
    ep = dev(np.zeros(n, np.int32))
    term = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    rew = torch.empty(n, dtype=torch.float64, device="cuda")
    done = torch.empty(n, dtype=torch.uint8, device="cuda")
    p = params(seed, env_offset, target)
    rc = lib.cstr_vec_step_f64(byref(p), n, auto_reset, L.ptr(a), L.ptr(s), L.ptr(sc), L.ptr(ep), None, L.ptr(term), L.ptr(rew), L.ptr(done),
                               None, None, None, None, torch.cuda.current_stream().cuda_stream)
    L.check(rc, "cstr_vec_step_f64")
    torch.cuda.synchronize()
    return dict(state=s.cpu().numpy(), terminal=term.cpu().numpy(), reward=rew.cpu().numpy(), done=done.cpu().numpy().astype(bool),
                step_count=sc.cpu().numpy())
