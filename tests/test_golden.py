"""Oracle restatement vs the committed fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only.

Tolerances: everything is bit-exact except float32 ``exp`` — NumPy's float32 exp is a SIMD kernel
whose code path depends on the host CPU — so observations are allowed 1 raw-state
ulp of drift per step (|Δobs| <= 6e-7 normalised) when the host differs from the one that generated
the fixtures; on the generating host the comparison is bit-exact (asserted in
test_oracle_vs_reference.py against the live reference).
"""
import numpy as np
import pytest

import cstr_oracle as O

OBS_TOL = 6.0e-7  # 1 ulp of a raw temperature (3.05e-5 K) mapped to the normalised scale, rounded up
REW_TOL = 1.0e-5


def test_step_f32_matches_reference_fixture(golden):
    g = golden("step_f32.npz")
    out = O.step_f32(g["states"], g["actions"], g["step_count"])
    assert np.array_equal(out.truncated, g["truncated"])
    assert not g["terminated"].any()
    assert np.array_equal(out.nan_row, np.isnan(g["conc_reward"]))
    np.testing.assert_allclose(out.obs, g["obs"], rtol=0, atol=OBS_TOL)
    np.testing.assert_allclose(out.reward, g["reward"], rtol=0, atol=REW_TOL)
    # rows that did not go through exp-sensitive rounding must be identical; report the exact share
    exact = (out.obs == g["obs"]).all(axis=1).mean()
    assert exact > 0.95, exact


def test_step_f32_edge_rows(golden):
    g = golden("step_f32.npz")
    out = O.step_f32(g["states"], g["actions"], g["step_count"])
    # NaN action rows: state unchanged, reward -10, truncated (Q4)
    for i in (21, 22):
        assert out.nan_row[i]
        assert np.array_equal(out.obs[i], g["states"][i])
        assert out.reward[i] == np.float32(-10.0) and out.truncated[i]
        assert out.step_count[i] == g["step_count"][i] + 1
    # +-inf / out-of-range actions are clipped, not errors
    assert not out.nan_row[20] and not out.nan_row[23]
    assert np.isfinite(out.obs).all()


def test_trajectory_with_autoreset(golden):
    g = golden("traj_f32.npz")
    T, N = g["reward"].shape
    seed = int(g["seed"])
    gens = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed + i))) for i in range(N)]

    def fresh(idx):
        u = np.stack([gens[i].random(8) for i in idx])
        return O.obs_from_raw_f64(O.initial_state_from_uniforms(u))

    obs0 = fresh(range(N))
    assert np.array_equal(obs0, g["obs0"])
    vec = O.VecOracle(obs0)
    for t in range(T):
        st = vec.step(g["actions"][t], reset_fn=fresh)
        assert np.array_equal(st.done, g["done"][t]), t
        assert np.array_equal(st.timeout, g["timeout"][t]), t
        np.testing.assert_allclose(st.obs, g["obs"][t], rtol=0, atol=1e-5)  # <=1e-5 over 400 steps
        np.testing.assert_allclose(st.terminal_obs, g["terminal_obs"][t], rtol=0, atol=1e-5)
        np.testing.assert_allclose(st.reward, g["reward"][t], rtol=0, atol=REW_TOL * 5)
    assert g["done"][399].all() and g["done"].sum() == N  # exactly the step-400 truncation row


def test_dynamics_f64(golden):
    g = golden("dyn_f64.npz")
    out = O.dynamics(g["raw_state"], g["raw_action"])
    assert out.dtype == np.float64
    np.testing.assert_allclose(out, g["new_raw_state"], rtol=1e-12, atol=0)


def test_reset_random_and_static(golden):
    g = golden("reset.npz")
    for k, seed in enumerate(g["seeds"]):
        E = g["random_obs"].shape[1]
        u = O.pcg64_reset_uniforms(int(seed), E)
        raw = O.initial_state_from_uniforms(u)
        assert np.array_equal(raw, g["random_raw"][k])
        assert np.array_equal(O.obs_from_raw_f64(raw), g["random_obs"][k])
        gen = np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(seed))))
        base = np.array([[0.45, 310.0, 0.25, 290.0]])
        for e in range(E):
            raw_s = O.static_state_from_uniforms(base, gen.random((1, 4)))
            assert np.array_equal(O.obs_from_raw_f64(raw_s)[0], g["static_obs"][k, e])  # Q2 drift


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_replay_buffer(golden, tag):
    g = golden("replay.npz")
    size, n_envs, n_add, batch = (int(v) for v in g[f"{tag}_cfg"])
    buf = O.ReplayOracle(size, n_envs)
    for i in range(n_add):
        buf.add(g[f"{tag}_add_obs"][i], g[f"{tag}_add_next_obs"][i], g[f"{tag}_add_action"][i],
                g[f"{tag}_add_reward"][i], g[f"{tag}_add_done"][i], g[f"{tag}_add_timeout"][i])
    for name in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts"):
        assert np.array_equal(getattr(buf, name), g[f"{tag}_store_{name}"]), name
    assert buf.pos == int(g[f"{tag}_pos"]) and buf.full == bool(g[f"{tag}_full"])
    np.random.seed(11)
    bi, ei = buf.draw_indices(batch)
    assert np.array_equal(bi, g[f"{tag}_batch_inds"]) and np.array_equal(ei, g[f"{tag}_env_inds"])
    obs, act, nobs, dones, rew = buf.gather(bi, ei)
    assert np.array_equal(obs, g[f"{tag}_s_obs"]) and np.array_equal(act, g[f"{tag}_s_act"])
    assert np.array_equal(nobs, g[f"{tag}_s_next_obs"]) and np.array_equal(dones, g[f"{tag}_s_dones"])
    assert np.array_equal(rew, g[f"{tag}_s_rewards"])


def test_actor_and_action_maps(golden):
    g = golden("td3_actor.npz")
    w = [(g["W1"], g["b1"]), (g["W2"], g["b2"]), (g["W3"], g["b3"])]
    assert g["W1"].shape == (400, 4) and g["W2"].shape == (300, 400) and g["W3"].shape == (2, 300)
    mu = O.actor_forward(g["obs"], w)
    np.testing.assert_allclose(mu, g["mu"], rtol=0, atol=2e-6)  # fp32 sgemm vs fp64 accumulate
    # predict() = unscale(mu) for the squashed TD3 actor (policies.py:375)
    a, s = O.sample_action_maps(g["mu"], g["noise"])
    assert np.array_equal(a, g["env_action"]) and np.array_equal(s, g["buffer_action"])
    a0, _ = O.sample_action_maps(g["mu"], np.zeros_like(g["noise"]))
    np.testing.assert_allclose(a0, g["predict"], rtol=0, atol=1.2e-7)


def test_vecnormalize_oracle_vs_reference_fixture(golden):
    """VecNormalizeOracle replays the reference's VecNormalize(DummyVecEnv) run (vec_normalize.py:174-298)."""
    g = golden("vecnorm.npz")
    T, N = g["raw_rew"].shape
    same, exact = O.VecNormalizeOracle(N), O.VecNormalizeOracle(N, exact=True)
    assert np.array_equal(same.reset(g["raw_obs0"]), g["norm_obs0"])
    exact.reset(g["raw_obs0"])
    for t in range(T):
        nobs, nrew = same.step(g["raw_obs"][t], g["raw_rew"][t], g["done"][t])
        exact.step(g["raw_obs"][t], g["raw_rew"][t], g["done"][t])
        assert np.array_equal(nobs, g["norm_obs"][t]) and np.array_equal(nrew, g["norm_rew"][t])
        np.testing.assert_allclose(same.obs_rms.mean, g["obs_mean"][t], rtol=1e-12)
        np.testing.assert_allclose(same.obs_rms.var, g["obs_var"][t], rtol=1e-12)
        np.testing.assert_allclose(same.ret_rms.var, g["ret_var"][t], rtol=1e-12)
        np.testing.assert_allclose(same.returns, g["returns"][t], rtol=0, atol=0)
        np.testing.assert_allclose(exact.obs_rms.var, g["obs_var"][t], rtol=2e-5)  # float64 vs float32 batch moments
        np.testing.assert_allclose(exact.ret_rms.var, g["ret_var"][t], rtol=2e-5)
    assert same.obs_rms.count == pytest.approx(float(g["obs_count"][-1]))
    # the fused sample of the fixture is the normalised gather of the stored ORIGINAL transitions (buffers.py:314-323)
    b, e = g["batch_inds"], g["env_inds"]
    assert np.array_equal(same.normalize_obs(g["store_observations"][b, e]), g["s_obs"])
    assert np.array_equal(same.normalize_reward(g["store_rewards"][b, e].reshape(-1, 1)), g["s_rewards"])


def test_td3_update_oracle_vs_reference_fixture(golden):
    """TD3UpdateOracle replays 6 gradient steps of the reference's TD3.train (td3.py:154-211) on the recorded batches and noise."""
    import td3_oracle as T
    import td3_util as U

    g = golden("td3_update.npz")
    o = U.make_oracle(T, g)
    final = U.replay(o, g)
    ref = U.nets_from(g, "final")
    for name in U.NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=5e-6)  # measured 9e-7 (float32 GEMM summation order)
    assert np.mean(o.critic_losses) == pytest.approx(float(g["critic_loss_mean"]), rel=1e-5)
    assert np.mean(o.actor_losses) == pytest.approx(float(g["actor_loss_mean"]), rel=1e-5)
    assert o.critic_opt.step_count == int(g["adam_critic_step"]) == 6 and o.actor_opt.step_count == int(g["adam_actor_step"]) == 3
    for i in range(12):
        np.testing.assert_allclose(o.critic_opt.m[i], g[f"adam_critic_m_{i}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(o.critic_opt.v[i], g[f"adam_critic_v_{i}"], rtol=1e-4, atol=1e-9)


def test_sac_update_oracle_vs_reference_fixture(golden):
    """SACUpdateOracle replays 5 gradient steps of the reference's SAC.train (sac.py:199-296): actor, twin critics, targets,
    automatic entropy coefficient, on the recorded batches and the two rsample() noise draws of every step."""
    import td3_oracle as T
    import td3_util as U

    g = golden("sac_update.npz")
    o = U.make_sac_oracle(T, g)
    final = U.replay_sac(o, g)
    ref = U.sac_nets_from(g, "final")
    for name in U.SAC_NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)  # measured 2.7e-6
    np.testing.assert_allclose(o.log_ent_coef[0], g["final_log_ent_coef"], rtol=0, atol=1e-7)
    for got, key in ((o.critic_losses, "critic_loss_mean"), (o.actor_losses, "actor_loss_mean"), (o.ent_coefs, "ent_coef_mean"),
                     (o.ent_coef_losses, "ent_coef_loss_mean")):
        assert np.mean(got) == pytest.approx(float(g[key]), rel=1e-5)


def test_bcq_update_oracle_vs_reference_fixture(golden):
    """BCQUpdateOracle replays 5 gradient steps of the reference's BCQ.train (bcq.py:129-205): VAE step, candidate target through the
    refreshed target VAE + target perturbation net (with the reference's (B, 10) reshape as written), twin critics, delayed perturbation
    step, polyak — on the recorded batches and the recorded randn / randn_like draws.  This pins the oracle for the BCQ update kernels
    (cstr_bcq_update); the CUDA counterpart of this test is tests/test_gpu_bcq_ma.py."""
    import td3_oracle as T
    import td3_util as U

    g = golden("bcq_update.npz")
    o = U.make_bcq_oracle(T, g)
    final = U.replay_bcq(o, g)
    ref = U.bcq_nets_from(g, "final")
    for name in U.BCQ_NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=5e-6, err_msg=name)  # measured 2.4e-7
    for got, key in ((o.vae_losses, "vae_loss_mean"), (o.critic_losses, "critic_loss_mean"), (o.actor_losses, "actor_loss_mean")):
        assert np.mean(got) == pytest.approx(float(g[key]), rel=1e-5)
    assert o.vae_opt.step_count == o.critic_opt.step_count == 5 and o.pert_opt.step_count == 2
    # the max over "candidates" is the reference's row-major reshape of a candidate-major column: not a per-observation max
    B = g["batch_obs"].shape[1]
    col = np.arange(10 * B).reshape(10 * B, 1)
    assert not np.array_equal(col.reshape(B, 10).max(1), col.reshape(10, B).max(0))


@pytest.mark.parametrize("algo,centralised", [("maddpg", True), ("iddpg", False)])
def test_multi_agent_update_oracle_vs_reference_fixture(golden, algo, centralised):
    """MultiAgentDDPGOracle replays 4 gradient steps of the reference's MADDPG.train / IDDPG.train (two agents = the two reactors, unequal
    learning rates so the reference's actor/critic learning-rate pairing shows) on the recorded batches and target-noise draws.  Pins the
    oracle for the multi-agent update kernels (cstr_ma_update); the CUDA counterpart is tests/test_gpu_bcq_ma.py."""
    import td3_oracle as T
    import td3_util as U

    g = golden(f"{algo}_update.npz")
    o = U.make_ma_oracle(T, g, centralised)
    final = U.replay_ma(o, g)
    ref = U.ma_nets_from(g, "final")
    for name in U.MA_NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-6, err_msg=name)  # measured 6e-8
    for i in range(2):
        assert np.mean(o.critic_losses[i]) == pytest.approx(float(g[f"critic_loss_mean_{i}"]), rel=1e-5)
        assert np.mean(o.actor_losses[i]) == pytest.approx(float(g[f"actor_loss_mean_{i}"]), rel=1e-5, abs=1e-7)
    # the naive reading (agent i trains at learning_rate_list[i]) does NOT reproduce the reference
    n = U.ma_nets_from(g, "init")
    lr0, lr1 = float(g["hyper"][4]), float(g["hyper"][5])
    naive = T.MultiAgentDDPGOracle([n["actor0"], n["actor1"]], [[n["critic0_0"], n["critic0_1"]], [n["critic1_0"], n["critic1_1"]]],
                                   [[0, 1], [2, 3]], [[0], [1]], centralised, [lr0, lr1], [lr0, lr1], gamma=float(g["hyper"][0]),
                                   tau=float(g["hyper"][1]), policy_delay=int(g["hyper"][2]), target_noise_clip=float(g["hyper"][3]))
    U.replay_ma(naive, g)
    assert np.abs(naive.critics[0][0][0] - ref["critic0_0"][0]).max() > 1e-4


def test_ddpg_update_oracle_vs_reference_fixture(golden):
    """The one-critic path of TD3UpdateOracle against the reference's DDPG (core/ddpg/ddpg.py: TD3 with n_critics=1, policy_delay=1,
    target_noise_clip=0 — the smoothing noise is still drawn, then clamped to zero), 4 gradient steps."""
    import td3_oracle as T
    import td3_util as U

    g = golden("ddpg_update.npz")
    gamma, tau, delay, _sigma, clip, lr = [float(x) for x in g["hyper"]]
    assert int(g["n_critics"]) == 1 and int(delay) == 1 and clip == 0.0
    n = U.nets_from(g, "init")
    o = T.TD3UpdateOracle(n["actor"], [n["critic0"]], n["actor_target"], [n["critic0_target"]], lr=lr, gamma=gamma, tau=tau, policy_delay=1,
                          target_noise_clip=clip)
    for k in range(g["noise"].shape[0]):
        o.step(g["batch_obs"][k], g["batch_act"][k], g["batch_next_obs"][k], g["batch_dones"][k], g["batch_rewards"][k], g["noise"][k])
    ref = U.nets_from(g, "final")
    for name, got in (("actor", o.actor), ("critic0", o.critics[0]), ("actor_target", o.actor_target), ("critic0_target", o.critic_targets[0])):
        for a, b in zip(got, ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=2e-6, err_msg=name)  # measured 3e-8
    assert np.mean(o.critic_losses) == pytest.approx(float(g["critic_loss_mean"]), rel=1e-5)
    assert np.mean(o.actor_losses) == pytest.approx(float(g["actor_loss_mean"]), rel=1e-5)
