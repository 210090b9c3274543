"""-m gpu, needs the reference tree too (``oracle/refload.py``: /root/reference or the staged, git-ignored ``baseline/_ref``; skipped
otherwise): the reference's OWN ``learn()`` (core/common/off_policy_algorithm.py:309-355) driven by the fused rollout kernel through
``bind_offpolicy_rollout`` — collect_rollouts is one ``cstr_rollout_fused`` launch, train() is ``cstr_td3_update`` / ``cstr_sac_update``.
"""
import numpy as np
import pytest
import torch

import refload

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refload.available(), reason="reference tree not staged")]


@pytest.fixture(scope="module")
def ref(pkg):
    refload.install_shims()
    core = refload.load_core()
    from core.common.buffers import ReplayBuffer
    from core.common.noise import NormalActionNoise
    from core.common.vec_env import VecEnv

    return dict(core=core, VecEnv=pkg.bind_vec_env_class(VecEnv), Buffer=pkg.bind_replay_buffer_class(ReplayBuffer), Noise=NormalActionNoise,
                ReplayBuffer=ReplayBuffer)


def _td3(pkg, ref, n_envs, actor_mode="fp32", **kw):
    core = ref["core"]
    cls = pkg.bind_offpolicy_rollout(pkg.bind_td3_class(core.TD3), actor_mode=actor_mode)
    env = ref["VecEnv"](num_envs=n_envs, init_mode="random", reset_rng="philox", seed=3)
    args = dict(action_noise=ref["Noise"](mean=np.zeros(2), sigma=0.1 * np.ones(2)), device="cuda", replay_buffer_class=ref["Buffer"],
                buffer_size=64 * n_envs, replay_buffer_kwargs=dict(index_mode="philox", seed=1), learning_starts=6 * n_envs, batch_size=128,
                train_freq=(4, "step"), gradient_steps=2, seed=5, verbose=0)
    args.update(kw)
    return cls("MlpPolicy", env, **args), env


@pytest.mark.parametrize("actor_mode", ["fp32", "tc"])
def test_td3_learn_runs_on_the_fused_rollout(pkg, ref, actor_mode):
    core = ref["core"]
    n = 512
    model, env = _td3(pkg, ref, n, actor_mode)
    assert isinstance(model, core.TD3)
    dumps, seen = [], []

    class Cb(__import__("core.common.callbacks", fromlist=["BaseCallback"]).BaseCallback):
        def _on_training_start(self) -> None:  # the reference's Logger clears name_to_value inside dump(): look at it just before
            logger, orig = self.model.logger, self.model.logger.dump

            def spy(step=0):
                dumps.append(dict(logger.name_to_value))
                orig(step)

            logger.dump = spy

        def _on_step(self) -> bool:
            seen.append((self.model.num_timesteps, self.locals["num_collected_steps"]))
            return True

    launches_before = env.launches
    total = n * 4 * 110  # 110 launches of 4 steps: 440 env steps -> every reactor finishes one 400-step episode
    model.learn(total_timesteps=total, callback=Cb(), log_interval=100)
    assert model.num_timesteps == total and model.fused_rollout_launches == 110 + 1  # the launch that straddles learning_starts is split in two
    assert env.launches - launches_before <= 1  # nothing but the reset went through the per-step VecEnv path
    assert model.replay_buffer.launches <= model._n_updates  # ... and the buffer only served samples (fewer once graphs replay): no add() from Python
    assert model._n_updates == 2 * 109 and model._fused is not None and model._fused.n_updates == model._n_updates
    # learning_starts = 6 steps: the first launch (4 steps) and half of the second are warm-up (uniform actions), the rest the actor's
    rec = model.replay_buffer.records
    assert model.replay_buffer.pos == (4 * 110) % 64 and model.replay_buffer.full
    # callbacks: once per launch, with num_timesteps advanced by the whole launch
    assert len(seen) == 110 and seen[0] == (4 * n, 4) and seen[-1][0] == total
    # Monitor semantics from the device: every reactor finished exactly one episode of length 400
    assert model._episode_num == n and len(model.ep_info_buffer) == model.ep_info_buffer.maxlen
    assert all(e["l"] == 400 and -2000 < e["r"] < 0 for e in model.ep_info_buffer)
    assert dumps and "rollout/ep_rew_mean" in dumps[-1] and "rollout/ep_len_mean" in dumps[-1] and "time/fps" in dumps[-1]
    assert dumps[-1]["rollout/ep_len_mean"] == 400 and "train/critic_loss" in dumps[-1]
    assert isinstance(model._last_obs, np.ndarray) and model._last_obs.shape == (n, 4)
    # the stored transitions are the strict step of the stored action (spot check on the newest row)
    import build_oracle as B

    row = (model.replay_buffer.pos - 1) % 64
    r = rec[row].cpu().numpy()
    nx, rw, _, _, _ = B.step_f32(r[:, 0:4], r[:, 8:10], np.full(n, 10, np.int32), exp_mode=B.EXP_SHARED, sq_mode=B.SQ_MUL)
    assert np.array_equal(r[:, 4:8], nx) and np.array_equal(r[:, 10], rw)
    assert np.abs(r[:, 8:10]).max() <= 1.0


def test_sac_learn_runs_on_the_fused_rollout(pkg, ref):
    core = ref["core"]
    n = 256
    cls = pkg.bind_offpolicy_rollout(pkg.bind_sac_class(core.SAC))
    env = ref["VecEnv"](num_envs=n, init_mode="random", reset_rng="philox", seed=4)
    model = cls("MlpPolicy", env, device="cuda", replay_buffer_class=ref["Buffer"], buffer_size=64 * n,
                replay_buffer_kwargs=dict(index_mode="philox", seed=2), learning_starts=8 * n, batch_size=128, train_freq=(8, "step"), gradient_steps=4,
                seed=1, verbose=0)
    before = [p.detach().clone() for p in model.actor.parameters()]
    model.learn(total_timesteps=n * 8 * 12, log_interval=None)
    assert model.fused_rollout_launches == 12 and model._fused is not None and model._n_updates == 4 * 11
    assert any(not torch.equal(a, b) for a, b in zip(before, model.actor.parameters()))
    rec = model.replay_buffer.records[(model.replay_buffer.pos - 1) % 64].cpu().numpy()
    assert np.abs(rec[:, 8:10]).max() < 1.0  # tanh-squashed Gaussian actions, no extra noise


def test_unsupported_models_keep_the_reference_rollout(pkg, ref):
    core = ref["core"]
    n = 8
    cls = pkg.bind_offpolicy_rollout(core.TD3)
    env = ref["VecEnv"](num_envs=n, init_mode="random", reset_rng="pcg64", seed=3)  # host-RNG resets: not the kernel's stream
    model = cls("MlpPolicy", env, device="cuda", replay_buffer_class=ref["Buffer"], buffer_size=64 * n, learning_starts=2 * n, batch_size=16,
                train_freq=(1, "step"), gradient_steps=1, seed=0, verbose=0)
    with pytest.warns(RuntimeWarning, match="fused rollout not used"):
        model.learn(total_timesteps=6 * n, log_interval=None)
    assert model.fused_rollout_launches == 0 and model.num_timesteps == 6 * n and model.replay_buffer.pos == 6


@pytest.mark.parametrize("algo", ["MADDPG", "IDDPG"])
def test_multi_agent_learn_runs_on_the_fused_rollout_and_update(pkg, ref, algo):
    """BASELINE config #5 in small: the reference's MADDPG / IDDPG ``learn()`` with ``collect_rollouts`` = cstr_rollout_fused_multi and
    ``train()`` = cstr_ma_update (bind_offpolicy_rollout(bind_multiagent_class(...)))."""
    core = ref["core"]
    n = 256
    cls = pkg.bind_offpolicy_rollout(pkg.bind_multiagent_class(getattr(core, algo)))
    env = ref["VecEnv"](num_envs=n, init_mode="random", reset_rng="philox", seed=6)
    model = cls(policy="MlpPolicy", env=env, n_agents=2, observation_splits=[[0, 1], [2, 3]], action_splits=[[0], [1]], learning_rate_list=[1e-3, 5e-4],
                device="cuda", replay_buffer_class=ref["Buffer"], buffer_size=64 * n, replay_buffer_kwargs=dict(index_mode="philox", seed=2),
                learning_starts=8 * n, batch_size=128, train_freq=(8, "step"), gradient_steps=4, seed=1, verbose=0)
    assert isinstance(model, getattr(core, algo))
    before = [p.detach().clone() for p in model.actor.mu_list[1].parameters()]
    launches_before = env.launches
    model.learn(total_timesteps=n * 8 * 12, log_interval=None)
    assert model.fused_rollout_launches == 12 and env.launches - launches_before <= 1
    assert model._fused is not None and model._n_updates == 4 * 11 == model._fused.n_updates
    assert model._fused.centralised == (algo == "MADDPG")
    # the reference's learning-rate pairing: every actor at learning_rate_list[0], every critic at learning_rate_list[1]
    assert model._fused.actor_lrs == [1e-3, 1e-3] and model._fused.critic_lrs == [5e-4, 5e-4]
    assert any(not torch.equal(a, b) for a, b in zip(before, model.actor.mu_list[1].parameters()))
    assert model.actor.mu_list[0][0].weight.data_ptr() == model._fused.views("params")["actor0"][0].data_ptr()
    rec = model.replay_buffer.records[(model.replay_buffer.pos - 1) % 64].cpu().numpy()
    assert np.abs(rec[:, 8:10]).max() <= 1.0


def test_bcq_learn_runs_on_the_fused_update(pkg, ref, tmp_path):
    """BASELINE config #4 in small: a CSTR dataset produced by the tape kernel, handed to the reference's ``BCQ`` through its pickle path
    (offline_policy_algorithm.py:196-242, quirk Q7), adopted into a GpuReplayBuffer (Q8), and trained by ``BCQ.learn()`` whose ``train()`` is
    cstr_bcq_update (bind_bcq_class)."""
    import pickle

    core, ReplayBuffer = ref["core"], ref["ReplayBuffer"]
    from core.common.vec_env import DummyVecEnv

    n, T = 64, 100
    env = pkg.GpuCSTRVecEnv(n, seed=3, monitor=False)
    env.reset()
    obs0 = env.state.clone()
    acts = torch.rand((T, n, 2), device="cuda") * 2 - 1
    res = env.tape(T, acts, want_obs=True)
    obs = torch.cat([obs0[None], res["obs"][:-1]], 0)
    ref_buf = ReplayBuffer(n * T, env.observation_space, env.action_space, device="cpu", n_envs=1)
    ref_buf.observations[:, 0] = obs.reshape(-1, 4).cpu().numpy()
    ref_buf.next_observations[:, 0] = res["obs"].reshape(-1, 4).cpu().numpy()
    ref_buf.actions[:, 0] = acts.reshape(-1, 2).cpu().numpy()
    ref_buf.rewards[:, 0] = res["rewards"].reshape(-1).cpu().numpy()
    ref_buf.full, ref_buf.pos = True, 0
    path = tmp_path / "cstr_dataset.pkl"
    with open(path, "wb") as fh:
        pickle.dump(ref_buf, fh)
    import gymnasium

    Facade = pkg.bind_env_class(gymnasium.Env)  # the harness put gymnasium on sys.path after the package was imported
    single = DummyVecEnv([lambda: Facade(init_mode="static")])
    FusedBCQ = pkg.bind_bcq_class(core.BCQ)
    model = FusedBCQ("MlpPolicy", single, dataset=str(path), batch_size=128, device="cuda", seed=0)
    model.replay_buffer = ref["Buffer"].from_reference(model.replay_buffer, index_mode="philox", seed=1)
    vae_before = model.actor.vae.decoder[0].weight.detach().clone()
    model.learn(total_timesteps=60)
    assert isinstance(model, core.BCQ) and model._fused is not None and model._fused.n_updates == model._n_updates > 0
    assert model._fused.actor_step == model._n_updates // model.actor_delay
    assert not torch.equal(vae_before, model.actor.vae.decoder[0].weight)
    # shared storage: policy modules are views of the flat block, and the target VAE is the VAE
    assert model.actor.vae.decoder[0].weight.data_ptr() == model._fused.views("params")["vae_dec"][0].data_ptr()
    assert model.actor_target.vae.mean.weight.data_ptr() == model.actor.vae.mean.weight.data_ptr()
    a, _ = model.predict(np.zeros((1, 4), np.float32), deterministic=True)
    assert a.shape == (1, 2) and np.abs(a).max() <= 1.0
