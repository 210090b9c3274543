"""Pins the oracle against the UNMODIFIED reference, live (only where /root/reference exists —
the build container; skipped on the GPU box, which re-checks the frozen fixtures instead)."""
import contextlib
import io

import numpy as np
import pytest

import cstr_oracle as O
import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not present on this host")


@pytest.fixture(scope="module")
def ref_env_module():
    return refload.load_env_module()


def _ref_steps(m, states, actions, step_count):
    env = m.TwoSeriesCSTREnv()
    env.reset(seed=0)
    N = len(states)
    obs = np.zeros((N, 4), np.float32)
    rew = np.zeros(N, np.float32)
    tr = np.zeros(N, bool)
    for i in range(N):
        env.state = states[i].copy()
        env.current_step = int(step_count[i])
        with contextlib.redirect_stdout(io.StringIO()):
            o, r, te, t, _ = env.step(actions[i].copy())
        assert te is False
        obs[i], rew[i], tr[i] = o, r, t
    return obs, rew, tr


def test_step_bit_exact(ref_env_module):
    rng = np.random.default_rng(99)
    N = 6000
    states = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    states[:64] = np.sign(states[:64])
    actions = rng.uniform(-1.5, 1.5, (N, 2)).astype(np.float32)
    actions[70] = [np.nan, 0.0]
    actions[71] = [np.inf, -np.inf]
    sc = rng.integers(0, 402, N).astype(np.int32)
    obs, rew, tr = _ref_steps(ref_env_module, states, actions, sc)
    out = O.step_f32(states, actions, sc)
    assert np.array_equal(out.obs, obs)
    assert np.array_equal(out.reward, rew)
    assert np.array_equal(out.truncated, tr)


def test_dynamics_f64_bit_exact(ref_env_module):
    rng = np.random.default_rng(5)
    env = ref_env_module.TwoSeriesCSTREnv()
    lo, hi = env.raw_state_low.astype(np.float64), env.raw_state_high.astype(np.float64)
    raw = lo + (hi - lo) * rng.random((1500, 4))
    act = 30.0 + 220.0 * rng.random((1500, 2))
    ref = np.stack([env._dynamics(state=raw[i].copy(), action=act[i].copy()) for i in range(len(raw))])
    assert np.array_equal(O.dynamics(raw, act), ref)


def test_reset_stream(ref_env_module):
    for seed in (3, 17):
        env = ref_env_module.TwoSeriesCSTREnv(init_mode="random")
        u = O.pcg64_reset_uniforms(seed, 3)
        mine = O.obs_from_raw_f64(O.initial_state_from_uniforms(u))
        for e in range(3):
            o, _ = env.reset(seed=seed if e == 0 else None)
            assert np.array_equal(o, mine[e])


def test_replay_buffer_live():
    refload.load_core()
    from core.common.buffers import ReplayBuffer
    from gymnasium import spaces

    rng = np.random.default_rng(1)
    n_envs, size = 3, 21
    ref = ReplayBuffer(size, spaces.Box(-1, 1, (4,), np.float32), spaces.Box(-1, 1, (2,), np.float32), "cpu", n_envs)
    mine = O.ReplayOracle(size, n_envs)
    for _ in range(17):
        o, no = rng.random((n_envs, 4), np.float32), rng.random((n_envs, 4), np.float32)
        a, r = rng.random((n_envs, 2), np.float32), rng.random(n_envs, np.float32)
        d = rng.random(n_envs) < 0.4
        to = d & (rng.random(n_envs) < 0.5)
        ref.add(o, no, a, r, d, [{"TimeLimit.truncated": bool(x)} for x in to])
        mine.add(o, no, a, r, d, to)
    for k in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts"):
        assert np.array_equal(getattr(ref, k), getattr(mine, k))
    assert (ref.pos, ref.full) == (mine.pos, mine.full)
    np.random.seed(4)
    s = ref.sample(50)
    np.random.seed(4)
    obs, act, nobs, dones, rew = mine.sample(50)
    assert np.array_equal(s.observations.numpy(), obs) and np.array_equal(s.actions.numpy(), act)
    assert np.array_equal(s.next_observations.numpy(), nobs) and np.array_equal(s.dones.numpy(), dones)
    assert np.array_equal(s.rewards.numpy(), rew)


def test_td3_update_default_arch_live(ref_env_module):
    """The TD3 gradient-step restatement against the reference's TD3.train at the default [400, 300] architecture."""
    import make_golden
    import td3_oracle as T
    import td3_util as U

    g = make_golden.td3_update_reference_run(ref_env_module, refload.load_core(), [400, 300], K=4, B=32)
    o = U.make_oracle(T, g)
    final = U.replay(o, g)
    ref = U.nets_from(g, "final")
    for name in U.NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)
    assert np.mean(o.critic_losses) == pytest.approx(float(g["critic_loss_mean"]), rel=1e-5)


def test_sac_update_default_arch_live(ref_env_module):
    """The SAC gradient-step restatement against the reference's SAC.train at its default [256, 256] architecture."""
    import make_golden
    import td3_oracle as T
    import td3_util as U

    g = make_golden.sac_update_reference_run(ref_env_module, refload.load_core(), [256, 256], K=3, B=32)
    o = U.make_sac_oracle(T, g)
    final = U.replay_sac(o, g)
    ref = U.sac_nets_from(g, "final")
    for name in U.SAC_NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-5)
    assert np.mean(o.actor_losses) == pytest.approx(float(g["actor_loss_mean"]), rel=1e-5)


def test_fused_update_support_check_live(ref_env_module):
    """``fused_update_unsupported`` on live reference models: the defaults qualify, anything the kernels do not implement is named —
    and a bound class then runs the reference's own torch ``train()`` (here on CPU, where the fused engine could not even be built)."""
    import importlib
    import warnings

    import torch

    core = refload.load_core()  # mirrors the tree first (the bare reference package cannot be imported from a read-only mount)
    from core.common.vec_env import DummyVecEnv

    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    upd = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200.update")
    torch.set_num_threads(1)
    venv = DummyVecEnv([lambda: ref_env_module.TwoSeriesCSTREnv(init_mode="random")])
    mk = lambda cls, **kw: cls("MlpPolicy", venv, buffer_size=500, batch_size=16, learning_starts=0, device="cpu", seed=1, **kw)  # noqa: E731
    assert upd.fused_update_unsupported(mk(core.TD3)) is None
    assert upd.fused_update_unsupported(mk(core.DDPG)) is None
    assert upd.fused_update_unsupported(mk(core.SAC)) is None
    assert "activation_fn" in upd.fused_update_unsupported(mk(core.TD3, policy_kwargs=dict(activation_fn=torch.nn.Tanh)))
    assert "net_arch" in upd.fused_update_unsupported(mk(core.TD3, policy_kwargs=dict(net_arch=[64])))
    assert "net_arch" in upd.fused_update_unsupported(mk(core.TD3, policy_kwargs=dict(net_arch=[64, 30])))
    assert "different" in upd.fused_update_unsupported(mk(core.SAC, policy_kwargs=dict(net_arch=dict(pi=[64, 64], qf=[32, 32]))))
    assert "n_critics" in upd.fused_update_unsupported(mk(core.TD3, policy_kwargs=dict(n_critics=3)))
    assert "optimizer" in upd.fused_update_unsupported(mk(core.TD3, policy_kwargs=dict(optimizer_class=torch.optim.SGD)))
    assert "optimizer" in upd.fused_update_unsupported(mk(core.TD3, policy_kwargs=dict(optimizer_kwargs=dict(weight_decay=1e-2))))
    # an unsupported model under the bound class trains through the reference's code path
    model = mk(pkg.bind_td3_class(core.TD3), policy_kwargs=dict(activation_fn=torch.nn.Tanh, net_arch=[32, 32]))
    before = [p.detach().clone() for p in model.critic.parameters()]
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        model.learn(total_timesteps=40)
    assert any("fused update not used" in str(x.message) for x in w)
    assert model._fused is None and model._n_updates > 0
    assert any(not torch.equal(a, b) for a, b in zip(before, model.critic.parameters()))


def test_bcq_update_default_arch_live(ref_env_module):
    """The BCQ gradient-step restatement against the reference's BCQ.train at BCQPolicy's default architecture
    (VAE 6-64-64-(32+32) / 36-64-64-2, perturbation 6-64-64-2, critics 6-400-300-1)."""
    import make_golden
    import td3_oracle as T
    import td3_util as U

    arch = dict(vae_latent_dim=32, vae_hidden_dim=64, perturbation_hidden_dim=64, max_perturbation=0.05)
    g = make_golden.bcq_update_reference_run(ref_env_module, refload.load_core(), K=4, B=24, arch=arch, critic_arch=(400, 300))
    o = U.make_bcq_oracle(T, g)
    final = U.replay_bcq(o, g)
    ref = U.bcq_nets_from(g, "final")
    for name in U.BCQ_NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-5, err_msg=name)
    assert np.mean(o.vae_losses) == pytest.approx(float(g["vae_loss_mean"]), rel=1e-5)


@pytest.mark.parametrize("algo,centralised", [("MADDPG", True), ("IDDPG", False)])
def test_multi_agent_update_default_arch_live(ref_env_module, algo, centralised):
    """The MADDPG / IDDPG gradient-step restatement against the reference's own train() at the default [400, 300] per-agent architecture."""
    import make_golden
    import td3_oracle as T
    import td3_util as U

    g = make_golden.multi_agent_reference_run(ref_env_module, refload.load_core(), algo, K=4, B=24, arch=(400, 300))
    o = U.make_ma_oracle(T, g, centralised)
    final = U.replay_ma(o, g)
    ref = U.ma_nets_from(g, "final")
    for name in U.MA_NETS:
        for a, b in zip(final[name], ref[name]):
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-5, err_msg=name)
