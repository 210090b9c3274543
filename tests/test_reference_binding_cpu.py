"""CPU (no GPU needed), with the unmodified reference loaded by oracle/refload.py (skipped where its tree is absent): the host-side logic of
the binding layer — what decides whether the fused kernels may stand in for the reference's collect_rollouts / train(), and the fallback to
the reference's own code when they may not.  The kernels themselves are exercised by the -m gpu tests."""
import importlib
import pickle
import warnings

import numpy as np
import pytest

import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not available")
PKG = "pytorch-rl-enhancedstablebaselines_b200"


@pytest.fixture(scope="module")
def ref():
    refload.install_shims()
    core = refload.load_core()
    env_mod = refload.load_env_module()
    from core.common.noise import NormalActionNoise, OrnsteinUhlenbeckActionNoise, VectorizedActionNoise
    from core.common.vec_env import DummyVecEnv

    return dict(core=core, env_mod=env_mod, Normal=NormalActionNoise, OU=OrnsteinUhlenbeckActionNoise, Vec=VectorizedActionNoise, DummyVecEnv=DummyVecEnv)


def _venv(ref, n=2):
    return ref["DummyVecEnv"]([lambda: ref["env_mod"].TwoSeriesCSTREnv() for _ in range(n)])


def test_noise_sigma_reads_what_the_kernel_can_draw(ref):
    r = importlib.import_module(PKG + ".rollout")
    assert r._noise_sigma(None) == 0.0
    assert r._noise_sigma(ref["Normal"](np.zeros(2), 0.1 * np.ones(2))) == pytest.approx(0.1)
    assert r._noise_sigma(ref["Vec"](ref["Normal"](np.zeros(2), 0.3 * np.ones(2)), 4)) == pytest.approx(0.3)
    assert r._noise_sigma(ref["Normal"](np.array([0.1, 0.0]), 0.1 * np.ones(2))) is None      # non-zero mean
    assert r._noise_sigma(ref["Normal"](np.zeros(2), np.array([0.1, 0.2]))) is None           # per-dimension sigma
    assert r._noise_sigma(ref["OU"](np.zeros(2), 0.1 * np.ones(2))) is None


def test_rollout_binding_falls_back_to_the_reference_on_a_foreign_env(ref):
    """bind_offpolicy_rollout on a model whose env is the reference's own DummyVecEnv: the reference's collect_rollouts runs (one warning), the
    bookkeeping of learn() is the reference's."""
    pkg = importlib.import_module(PKG)
    core = ref["core"]
    cls = pkg.bind_offpolicy_rollout(core.TD3)
    assert cls.__name__ == "TD3" and issubclass(cls, core.TD3)
    model = cls("MlpPolicy", _venv(ref, 2), action_noise=ref["Normal"](np.zeros(2), 0.1 * np.ones(2)), learning_starts=8, batch_size=8, buffer_size=1000,
                train_freq=(1, "step"), gradient_steps=1, device="cpu", seed=0, verbose=0)
    with pytest.warns(RuntimeWarning, match="fused rollout not used.*DummyVecEnv"):
        model.learn(total_timesteps=24, log_interval=None)
    assert model.num_timesteps == 24 and model.replay_buffer.pos == 12 and model.fused_rollout_launches == 0 and model._n_updates > 0
    with warnings.catch_warnings():
        warnings.simplefilter("error")  # warned once per model
        model.learn(total_timesteps=4, reset_num_timesteps=False, log_interval=None)
    with pytest.raises(TypeError):
        pkg.bind_offpolicy_rollout(dict)
    with pytest.raises(ValueError):
        pkg.bind_offpolicy_rollout(core.TD3, actor_mode="fp64")


def test_multi_agent_architecture_is_read_off_the_modules(ref):
    ue = importlib.import_module(PKG + ".update_ext")
    core = ref["core"]
    ma = dict(n_agents=2, observation_splits=[[0, 1], [2, 3]], action_splits=[[0], [1]], learning_rate_list=[1e-3, 5e-4])
    for algo, width in (("MADDPG", 6), ("IDDPG", 3)):
        model = getattr(core, algo)(policy="MlpPolicy", env=_venv(ref, 1), device="cpu", buffer_size=100, **ma)
        assert ue._ma_arch(model.policy) == [400, 300]
        assert int(next(model.critic.q_networks_list[0][0].parameters()).shape[1]) == width
        assert ue.multiagent_update_unsupported(model) is None
        small = getattr(core, algo)(policy="MlpPolicy", env=_venv(ref, 1), device="cpu", buffer_size=100, policy_kwargs=dict(net_arch=[[30, 20], [30, 20]]), **ma)
        assert ue._ma_arch(small.policy) is None and "multiples of 4" in ue.multiagent_update_unsupported(small)  # 30 is not a multiple of 4
        bound = importlib.import_module(PKG).bind_multiagent_class(getattr(core, algo))
        assert bound.__name__ == algo and issubclass(bound, getattr(core, algo))
    r = importlib.import_module(PKG + ".rollout")
    model = core.MADDPG(policy="MlpPolicy", env=_venv(ref, 1), device="cpu", buffer_size=100, **ma)
    # on a foreign env the multi-agent rollout is refused for the env, not for the policy
    assert "DummyVecEnv" in r.fused_rollout_unsupported(model, model.env, model.replay_buffer, model.train_freq, None)


def test_bcq_fit_check_reads_the_policy(ref, tmp_path):
    ue = importlib.import_module(PKG + ".update_ext")
    core = ref["core"]
    from core.common.buffers import ReplayBuffer

    venv = _venv(ref, 1)
    buf = ReplayBuffer(64, venv.observation_space, venv.action_space, device="cpu", n_envs=1)
    rng = np.random.default_rng(0)
    for _ in range(64):
        buf.add(rng.uniform(-1, 1, (1, 4)).astype(np.float32), rng.uniform(-1, 1, (1, 4)).astype(np.float32), rng.uniform(-1, 1, (1, 2)).astype(np.float32),
                np.zeros(1, np.float32), np.zeros(1, bool), [{}])
    path = tmp_path / "d.pkl"
    with open(path, "wb") as fh:
        pickle.dump(buf, fh)
    model = core.BCQ("MlpPolicy", venv, dataset=str(path), batch_size=16, device="cpu", seed=0)
    assert ue.bcq_update_unsupported(model) is None
    assert model.policy.actor_arch["vae_latent_dim"] == 32 and list(model.policy.critic_arch) == [400, 300]
    odd = core.BCQ("MlpPolicy", venv, dataset=str(path), batch_size=16, device="cpu", seed=0,
                   policy_kwargs=dict(actor_net_arch=dict(vae_latent_dim=6, vae_hidden_dim=64, perturbation_hidden_dim=64, max_perturbation=0.05)))
    assert "latent" in ue.bcq_update_unsupported(odd)
    bound = importlib.import_module(PKG).bind_bcq_class(core.BCQ)
    assert bound.__name__ == "BCQ" and issubclass(bound, core.BCQ)
