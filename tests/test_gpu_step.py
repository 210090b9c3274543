"""-m gpu: the CUDA step / reset / tape kernels, called through the C ABI, against the oracle.

Parity bars:
  strict fp32  == C/NumPy oracle in shared-exp mode, BIT-EXACT (obs, reward, done, counters);
               vs the reference's own NumPy arithmetic: |Δobs| <= 2e-6 per step, |Δreward| <= 1e-5
               (only float32 exp and powf(n,2) differ, by ulps), <= 1e-5 over a 400-step trajectory.
  fast fp32    |Δobs| <= 2e-6 per step, |Δreward| <= 1e-5.
  fp64         rel 1e-9 per step against the reference scheme in float64.
"""
import numpy as np
import pytest
import torch

import build_oracle as B
import cstr_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_util

    return gpu_util


def _inputs(n, seed=0, nan=True):
    rng = np.random.default_rng(seed)
    st = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    st[:64] = np.sign(st[:64])
    ac = rng.uniform(-1.3, 1.3, (n, 2)).astype(np.float32)
    sc = rng.integers(0, 402, n).astype(np.int32)
    if nan:
        ac[100] = [np.nan, 0.0]
        ac[101] = [0.3, np.nan]
        ac[102] = [np.inf, -np.inf]
    return st, ac, sc


def test_strict_step_bit_exact_vs_oracle(G):
    st, ac, sc = _inputs(200_000)
    out = G.raw_step_f32(st, ac, sc, mode=0)
    s, r, tr, sc2, bad = B.step_f32(st, ac, sc, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
    assert np.array_equal(out["terminal"], s) and np.array_equal(out["state"], s)
    assert np.array_equal(out["reward"], r)
    assert np.array_equal(out["done"], tr) and np.array_equal(out["timeout"], tr)
    assert np.array_equal(out["step_count"], sc2)
    # NaN rows: state unchanged, reward -10, truncated (Q4)
    assert np.array_equal(out["state"][100], st[100]) and out["reward"][100] == -10.0 and out["done"][100]
    # and the NumPy restatement with the same exp injected agrees too (three-way)
    o = O.step_f32(st, ac, sc, square=O.square_mul, exp=B.expf_shared)
    assert np.array_equal(out["terminal"], o.obs) and np.array_equal(out["reward"], o.reward)


def test_strict_step_vs_reference_fixture(G, golden):
    g = golden("step_f32.npz")
    out = G.raw_step_f32(g["states"], g["actions"], g["step_count"], mode=0)
    np.testing.assert_allclose(out["terminal"], g["obs"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["reward"], g["reward"], rtol=0, atol=1e-5)
    assert np.array_equal(out["done"], g["truncated"])
    exact = (out["terminal"] == g["obs"]).all(axis=1).mean()
    assert exact > 0.95, exact  # the rest differ by float32-exp ulps only


def test_fast_step_tolerance(G):
    st, ac, sc = _inputs(200_000, seed=1)
    out = G.raw_step_f32(st, ac, sc, mode=1)
    o = O.step_f32(st, ac, sc)
    np.testing.assert_allclose(out["terminal"], o.obs, rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["reward"], o.reward, rtol=0, atol=1e-5)
    assert np.array_equal(out["done"], o.truncated)


def test_f64_step_rel_1e9(G):
    st, ac, sc = _inputs(100_000, seed=2)
    st64, ac64 = st.astype(np.float64), ac.astype(np.float64)
    out = G.raw_step_f64(st64, ac64, sc)
    o = O.step_f64(st64, ac64, sc)
    ok = ~o.nan_row
    np.testing.assert_allclose(out["terminal"][ok], o.obs[ok], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out["reward"][ok], o.reward[ok], rtol=1e-9, atol=1e-12)
    assert np.array_equal(out["done"], o.truncated)
    assert np.array_equal(out["state"][~ok], st64[~ok]) and (out["reward"][~ok] == -10.0).all()


def test_f64_dynamics_fixture(G, golden):
    """fp64 kernel against the reference's own _dynamics fed float64 (fixture), through the affine maps."""
    g = golden("dyn_f64.npz")
    raw, act, new = g["raw_state"], g["raw_action"], g["new_raw_state"]
    lo, hi = O.RAW_STATE_LOW.astype(np.float64), O.RAW_STATE_HIGH.astype(np.float64)
    norm = 2.0 * (raw - lo) / (hi - lo) - 1.0
    a_norm = 2.0 * (act - 30.0) / 220.0 - 1.0
    out = G.raw_step_f64(norm, a_norm, np.zeros(len(raw), np.int32))
    got_raw = lo + (out["terminal"] + 1.0) * (hi - lo) / 2.0
    np.testing.assert_allclose(got_raw, new, rtol=1e-9, atol=1e-9)


def test_reset_kernel_bit_exact(G, pkg):
    n, seed, off = 10_000, 0xABCDEF0123, 7_000_000_000
    for init_mode in ("random", "static"):
        env = pkg.GpuCSTRVecEnv(n, seed=seed, env_offset=off, init_mode=init_mode)
        obs = env.reset()
        st, sc, ep, sb = B.reset_f32(n, off, seed, 0 if init_mode == "random" else 1)
        assert np.array_equal(obs, st)
        assert np.array_equal(env.episode.cpu().numpy(), ep)
        if init_mode == "static":
            assert np.array_equal(env.static_base.cpu().numpy(), sb)
            obs2 = env.reset()  # Q2: the base drifts
            st2, _, _, sb2 = B.reset_f32(n, off, seed, 1, episode=ep, static_base=sb)
            assert np.array_equal(obs2, st2) and np.array_equal(env.static_base.cpu().numpy(), sb2)


@pytest.mark.parametrize("mode,init_mode", [("strict", "random"), ("strict", "static")])
def test_tape_bit_exact_with_autoreset(G, pkg, mode, init_mode):
    n, T, seed = 3000, 403, 99
    rng = np.random.default_rng(3)
    env = pkg.GpuCSTRVecEnv(n, seed=seed, math=mode, init_mode=init_mode)
    obs0 = env.reset()
    acts = rng.uniform(-1, 1, (T, n, 2)).astype(np.float32)
    res = env.tape(T, torch.as_tensor(acts, device="cuda"), want_obs=True, reward_sum=True)
    im = 0 if init_mode == "random" else 1
    st, sc, ep, sb = B.reset_f32(n, 0, seed, im)
    ref = B.tape_f32(st, sc, ep, acts, 0, seed, im, static_base=sb, exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL, want_obs=True)
    assert np.array_equal(obs0, st)
    assert np.array_equal(res["rewards"].cpu().numpy(), ref["rewards"])
    assert np.array_equal(res["dones"].cpu().numpy().astype(bool), ref["dones"])
    assert np.array_equal(res["obs"].cpu().numpy(), ref["obs"])
    assert np.array_equal(env.state.cpu().numpy(), ref["state"])
    assert np.array_equal(env.step_count.cpu().numpy(), ref["step_count"])
    assert np.array_equal(env.episode.cpu().numpy(), ref["episode"])
    assert ref["dones"][399].all() and ref["dones"].sum() == n
    assert abs(res["reward_sum"].item() - ref["reward_sum"]) <= 1e-9 * abs(ref["reward_sum"])


def test_tape_equals_repeated_vec_steps(G, pkg):
    n, T = 2048, 6
    rng = np.random.default_rng(5)
    acts = rng.uniform(-1, 1, (T, n, 2)).astype(np.float32)
    e1 = pkg.GpuCSTRVecEnv(n, seed=5, monitor=False)
    e2 = pkg.GpuCSTRVecEnv(n, seed=5, monitor=False)
    e1.reset(), e2.reset()
    e1.step_count.fill_(396), e2.step_count.fill_(396)  # crosses the truncation row
    res = e1.tape(T, torch.as_tensor(acts, device="cuda"), want_obs=True)
    for t in range(T):
        obs, rew, done, infos = e2.step(acts[t])
        assert np.array_equal(obs, res["obs"][t].cpu().numpy())
        assert np.array_equal(rew, res["rewards"][t].cpu().numpy())
        assert np.array_equal(done, res["dones"][t].cpu().numpy().astype(bool))
        if t == 3:
            assert done.all() and "terminal_observation" in infos[0] and infos[0]["TimeLimit.truncated"] is True


def test_philox_actions_match_mirror(G, pkg):
    n, T, seed, off, t_base = 1000, 7, 42, 123456, 5
    env = pkg.GpuCSTRVecEnv(n, seed=seed, env_offset=off, monitor=False)
    env.reset()
    st = env.state.cpu().numpy().copy()
    res = env.tape(T, None, t_base=t_base, want_obs=True)
    ids = off + np.arange(n, dtype=np.uint64)
    acts = np.zeros((T, n, 2), np.float32)
    for t in range(T):
        g = t_base + t
        ctr = np.stack([(ids & 0xFFFFFFFF).astype(np.uint32), (ids >> np.uint64(32)).astype(np.uint32),
                        np.full(n, g >> 1, np.uint32), np.full(n, (2 << 8) | 0, np.uint32)], 1)
        r = O.philox4x32(ctr, np.array([seed, 0], np.uint32))
        w = r[:, 2:4] if g & 1 else r[:, 0:2]
        acts[t] = O.u32_to_unit_f32(w) * np.float32(2) - np.float32(1)
    e2 = pkg.GpuCSTRVecEnv(n, seed=seed, env_offset=off, monitor=False)
    e2.reset()
    ref = e2.tape(T, torch.as_tensor(acts, device="cuda"), want_obs=True)
    assert np.array_equal(res["obs"].cpu().numpy(), ref["obs"].cpu().numpy())
    assert np.array_equal(res["rewards"].cpu().numpy(), ref["rewards"].cpu().numpy())
    assert np.abs(acts).max() <= 1.0 and abs(float(acts.mean())) < 0.02


def test_full_size_properties(G, pkg):
    """BASELINE config #2 size (65,536 reactors x 400 steps): size-independent properties."""
    n, T = 65_536, 400
    env = pkg.GpuCSTRVecEnv(n, seed=7, monitor=False)
    env.reset()
    r1 = env.tape(T, None, reward_sum=True)
    # (a) checksum of checksums: atomically accumulated sum == sum of the per-step rewards
    total = r1["rewards"].double().sum().item()
    assert abs(r1["reward_sum"].item() - total) <= 1e-9 * abs(total)
    # (b) exactly one truncation row, at step 400, for every reactor; counters wrap to 0; episode advanced
    d = r1["dones"]
    assert int(d.sum().item()) == n and bool(d[399].all().item())
    assert int(env.step_count.abs().sum().item()) == 0 and int((env.episode - 2).abs().sum().item()) == 0
    # (c) observations stay inside the Box, rewards finite and <= 0 (both reward terms are penalties)
    s = env.state
    assert float(s.abs().max().item()) <= 1.0 and bool(torch.isfinite(r1["rewards"]).all().item())
    assert float(r1["rewards"].max().item()) <= 0.0
    # (d) sharding invariance: two half-size shards with env_offset reproduce the single launch bit for bit
    h = n // 2
    outs = []
    for k in range(2):
        e = pkg.GpuCSTRVecEnv(h, seed=7, env_offset=k * h, monitor=False)
        e.reset()
        outs.append(e.tape(T, None)["rewards"])
    assert torch.equal(torch.cat(outs, dim=1), r1["rewards"])
    # (e) fast mode stays within the trajectory tolerance of strict mode over the whole episode
    ef = pkg.GpuCSTRVecEnv(n, seed=7, math="fast", monitor=False)
    ef.reset()
    es = pkg.GpuCSTRVecEnv(n, seed=7, monitor=False)
    es.reset()
    of = ef.tape(399, None, want_obs=True, want_rewards=False, want_dones=False)["obs"][-1]
    os_ = es.tape(399, None, want_obs=True, want_rewards=False, want_dones=False)["obs"][-1]
    assert float((of - os_).abs().max().item()) <= 2e-5  # worst of 65,536 reactors x 399 steps (measured fp32-vs-fp64: 5.4e-6)


def test_trajectory_vs_reference_fixture(G, pkg, golden):
    """DummyVecEnv semantics end to end against the reference's own 405-step run (8 envs, seeds 100..107):
    host-PCG64 resets are bit-identical, observations within the documented exp tolerance."""
    g = golden("traj_f32.npz")
    T, N = g["reward"].shape
    env = pkg.GpuCSTRVecEnv(N, reset_rng="pcg64", math="strict")
    env.seed(int(g["seed"]))
    obs = env.reset()
    assert np.array_equal(obs, g["obs0"])
    for t in range(T):
        obs, rew, done, infos = env.step(g["actions"][t])
        assert obs.dtype == np.float32 and rew.dtype == np.float32 and done.dtype == bool and len(infos) == N
        assert np.array_equal(done, g["done"][t])
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=1e-5)
        np.testing.assert_allclose(rew, g["reward"][t], rtol=0, atol=5e-5)
        for i in range(N):
            assert infos[i].get("TimeLimit.truncated", False) == bool(g["timeout"][t, i])
            if done[i]:
                np.testing.assert_allclose(infos[i]["terminal_observation"], g["terminal_obs"][t, i], rtol=0, atol=1e-5)
                assert infos[i]["episode"]["l"] == 400
        if t == 399:
            assert done.all()
            assert np.array_equal(obs, g["obs"][t])  # post-reset obs: pure PCG64 + float64 math -> bit-exact


def test_vecenv_surface(pkg):
    env = pkg.GpuCSTRVecEnv(16, seed=3)
    assert env.num_envs == 16 and env.observation_space.shape == (4,) and env.action_space.shape == (2,)
    with pytest.raises(ValueError):
        env.step_async(np.zeros((16, 2), np.float32))  # reset first (twoseriescstr.py:402)
    assert env.seed(10) == list(range(10, 26))
    obs = env.reset()
    assert obs.shape == (16, 4) and obs.dtype == np.float32
    obs[:] = 0  # returned arrays must not alias internal state
    assert float(env.state.abs().sum().item()) > 0
    assert env.get_attr("render_mode") == [None] * 16 and env.get_attr("max_steps", [0, 3]) == [400, 400]
    assert env.get_attr("current_step", 2) == [0]
    assert env.env_method("set_target", 0.3, indices=[0]) == [True] and env.target_C2 == 0.3
    assert env.env_method("set_target", 0.9) == [False] * 16
    env.set_attr("target_C2", 0.2)
    assert env.env_is_wrapped(object) == [False] * 16
    assert env.has_attr("target_C2") and not env.has_attr("nope")
    with pytest.raises(AttributeError):
        env.get_attr("nope")
    assert env.unwrapped is env
    o, r, d, infos = env.step(np.zeros((16, 2), np.float32))
    assert len(infos) == 16 and infos[0].get("TimeLimit.truncated") is False and "episode" not in infos[0]
    env.close()


def test_single_env_facade(pkg, golden):
    g = golden("reset.npz")
    env = pkg.TwoSeriesCSTREnv(init_mode="random")
    obs, info = env.reset(seed=int(g["seeds"][2]))
    assert np.array_equal(obs, g["random_obs"][2, 0]) and info["initial_temperature_1"] == g["random_raw"][2, 0, 1]
    obs2, _ = env.reset()
    assert np.array_equal(obs2, g["random_obs"][2, 1])
    st = golden("step_f32.npz")
    for i in (30, 31, 21):
        env.state = st["states"][i].copy()
        env._vec.set_state(env.state[None, :], np.array([st["step_count"][i]], np.int32))
        env.current_step = int(st["step_count"][i])
        o, r, te, tr, info = env.step(st["actions"][i])
        assert te is False and tr == bool(st["truncated"][i])
        np.testing.assert_allclose(o, st["obs"][i], rtol=0, atol=2e-6)
        np.testing.assert_allclose(r, st["reward"][i], rtol=0, atol=1e-5)
    es = pkg.TwoSeriesCSTREnv(init_mode="static")
    for e in range(3):
        o, _ = es.reset(seed=int(g["seeds"][0]) if e == 0 else None)
        assert np.array_equal(o, g["static_obs"][0, e])


def test_abi_argument_errors(G):
    lib = G.L.load()
    from ctypes import byref

    p = G.params()
    assert lib.cstr_vec_step_f32(byref(p), 4, 0, 0, None, None, None, None, None, None, None, None, None, None, None, None, None) == -1
    assert b"null" in lib.cstr_last_error()
    s = torch.zeros(9, device="cuda")
    sc = torch.zeros(2, dtype=torch.int32, device="cuda")
    assert lib.cstr_reset(byref(p), 2, None, s.data_ptr() + 4, 0, sc.data_ptr(), sc.data_ptr(), None, None) == -2  # misaligned
    p.init_mode = 1
    assert lib.cstr_reset(byref(p), 2, None, s.data_ptr(), 0, sc.data_ptr(), sc.data_ptr(), None, None) == -1  # static w/o base
    assert lib.cstr_tape_f32(byref(G.params()), 0, 5, 0, None, 0, s.data_ptr(), sc.data_ptr(), sc.data_ptr(), None, None, None, None, None, None) == 0


def test_host_tape_entry_matches_device_tape(G, pkg):
    """cstr_tape_f32_host (chunked 3-stream H2D/compute/D2H pipeline over pinned host buffers) must return
    exactly what the device-resident tape returns, for a batch that is not a multiple of the chunk size."""
    from ctypes import byref

    n, T, seed = 40_000 + 37, 50, 3
    rng = np.random.default_rng(1)
    env = pkg.GpuCSTRVecEnv(n, seed=seed, monitor=False)
    obs0 = env.reset()
    env.step_count.fill_(370)
    acts = rng.uniform(-1, 1, (T, n, 2)).astype(np.float32)
    res = env.tape(T, torch.as_tensor(acts, device="cuda"))
    h_act = torch.as_tensor(acts).pin_memory()
    h_state = torch.as_tensor(obs0).pin_memory()
    h_sc = torch.full((n,), 370, dtype=torch.int32).pin_memory()
    h_ep = torch.ones(n, dtype=torch.int32).pin_memory()
    h_rew = torch.zeros((T, n), dtype=torch.float32).pin_memory()
    h_done = torch.zeros((T, n), dtype=torch.uint8).pin_memory()
    lib = G.L.load()
    p = G.params(seed=seed)
    G.L.check(lib.cstr_tape_f32_host(byref(p), n, T, 0, h_act.data_ptr(), h_state.data_ptr(), h_sc.data_ptr(), h_ep.data_ptr(),
                                     h_rew.data_ptr(), h_done.data_ptr(), None), "cstr_tape_f32_host")
    assert np.array_equal(h_rew.numpy(), res["rewards"].cpu().numpy())
    assert np.array_equal(h_done.numpy(), res["dones"].cpu().numpy())
    assert np.array_equal(h_state.numpy(), env.state.cpu().numpy())
    assert np.array_equal(h_sc.numpy(), env.step_count.cpu().numpy()) and np.array_equal(h_ep.numpy(), env.episode.cpu().numpy())
    assert h_done.numpy()[29].all()  # 370 + 30 = 400: the truncation row


def _fast_tape_vs_oracle(pkg, n, T, seed, start_steps):
    """tape_f32_kernel<fast> fed from an HBM action tape with obs + rewards + dones all requested (the instance bench.py times),
    against the C oracle (libm expf / powf: the reference's arithmetic) on the same initial states, counters and actions."""
    B.use_all_cores()
    rng = np.random.default_rng(seed)
    env = pkg.GpuCSTRVecEnv(n, seed=seed, math="fast", monitor=False)
    st0 = env.reset().copy()
    sc0 = np.asarray(start_steps, np.int32)
    env.step_count.copy_(torch.as_tensor(sc0, device="cuda"))
    acts = rng.uniform(-1, 1, (T, n, 2)).astype(np.float32)
    ep0 = env.episode.cpu().numpy().copy()
    res = env.tape(T, torch.as_tensor(acts, device="cuda"), want_obs=True, want_rewards=True, want_dones=True)
    obs, rew, done = res["obs"].cpu().numpy(), res["rewards"].cpu().numpy(), res["dones"].cpu().numpy().astype(bool)
    # the exact template instance bench.py launches (no observation tape) returns the same rewards / dones / final state, bit for bit
    e2 = pkg.GpuCSTRVecEnv(n, seed=seed, math="fast", monitor=False)
    e2.reset()
    e2.step_count.copy_(torch.as_tensor(sc0, device="cuda"))
    r2 = e2.tape(T, torch.as_tensor(acts, device="cuda"), want_obs=False, want_rewards=True, want_dones=True)
    assert torch.equal(r2["rewards"], res["rewards"]) and torch.equal(r2["dones"], res["dones"]) and torch.equal(e2.state, env.state)
    ref = B.tape_f32(st0, sc0, ep0, acts, 0, seed, 0, want_obs=True)
    out = {"dones_equal": bool(np.array_equal(done, ref["dones"]))}
    # (i) the whole trajectory, auto-resets included (reset rows are pure Philox + float64 math: bit-equal on both sides)
    out["traj_obs"] = float(np.abs(obs - ref["obs"]).max())
    out["traj_reward"] = float(np.abs(rew - ref["rewards"]).max())
    # (ii) per step from IDENTICAL states: the oracle steps once from the kernel's own previous observation
    prev = np.concatenate([st0[None], obs[:-1]], 0).reshape(-1, 4)
    sc = (sc0[None, :] + np.arange(T, dtype=np.int32)[:, None]) % 400  # counter before each step (wraps at the truncation row)
    s1, r1, tr1, _, _ = B.step_f32(prev, acts.reshape(-1, 2), sc.reshape(-1))
    keep = ~tr1  # the truncation row returns the post-reset observation; its terminal state is covered by (i) through the reward
    out["step_obs"] = float(np.abs(obs.reshape(-1, 4)[keep] - s1[keep]).max())
    out["step_reward"] = float(np.abs(rew.reshape(-1) - r1).max())
    out["step_dones_equal"] = bool(np.array_equal(done.reshape(-1), tr1))
    return out


def test_fast_tape_hbm_actions_rewards_dones_vs_oracle(pkg):
    """The benchmarked kernel instance (bench.py: fast math, HBM action tape, rewards + dones out) has its own oracle test, at the
    benchmark's size (65,536 x 400 = BASELINE config #2) and at a ragged size with resets inside the tape.
    Bars: per step from identical states |dobs| <= 2e-6, |dreward| <= 1e-5; dones equal; over the 400-step trajectory
    |dobs| <= 1e-5 (the dynamics are contractive: rounding differences do not accumulate) and |dreward| <= 5e-5."""
    import json
    import os

    measured = {}
    for name, n, T, starts in (("config2_65536x400", 65_536, 400, np.zeros(65_536, np.int32)),
                               ("ragged_1037x403", 1037, 403, np.random.default_rng(5).integers(0, 400, 1037).astype(np.int32))):
        m = _fast_tape_vs_oracle(pkg, n, T, seed=11, start_steps=starts)
        measured[name] = m
        assert m["dones_equal"] and m["step_dones_equal"], m
        assert m["step_obs"] <= 2e-6 and m["step_reward"] <= 1e-5, m
        assert m["traj_obs"] <= 1e-5 and m["traj_reward"] <= 5e-5, m
    print("fast tape vs oracle, measured maxima:", json.dumps(measured))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "fast_tape_parity.json"), "w") as f:
            json.dump(measured, f, indent=1)
