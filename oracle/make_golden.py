"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (TEST INFRASTRUCTURE).

Run in the build container (needs /root/reference):   python oracle/make_golden.py
The reference ships no tests or golden vectors of its own, so every fixture here is
the output of the reference's own code on seeded inputs:

  step_f32.npz   TwoSeriesCSTREnv.step            twoseriescstr.py:394-454  (4096 single steps incl. edge rows)
  traj_f32.npz   DummyVecEnv over 8 envs, 405 steps (auto-reset at 400)  dummy_vec_env.py:56-73
  dyn_f64.npz    TwoSeriesCSTREnv._dynamics fed float64                   twoseriescstr.py:456-503
  reset.npz      reset()/generate_initial_state, random + static modes    twoseriescstr.py:167-269
  replay.npz     ReplayBuffer.add/sample under np.random.seed             core/common/buffers.py:247-325
  td3_actor.npz  TD3Policy actor forward / predict / _sample_action maps  core/td3/policies.py:75-78,
                                                                           core/common/off_policy_algorithm.py:364-411
  vecnorm.npz    VecNormalize over DummyVecEnv (statistics, normalised obs/reward, normalised replay sample)
                                                                           core/common/vec_env/vec_normalize.py:174-298
  td3_update.npz TD3.train for 6 gradient steps (weights, Adam state, losses) on CPU torch   core/td3/td3.py:154-211
  sac_update.npz SAC.train for 5 gradient steps (actor, critics, entropy coefficient) on CPU torch  core/sac/sac.py:199-296
Host note: NumPy's float32 exp is a SIMD kernel whose code path depends on the CPU, so the fp32
fixtures are bit-stable only on hosts taking the same path; tests re-check them with the ulp-level
tolerance stated in tests/test_golden.py and bit-exactly in tests/test_oracle_vs_reference.py (live).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import refload  # noqa: E402

OUT = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def edge_rows(states: np.ndarray, actions: np.ndarray) -> None:
    """Overwrite the first rows with the edge cases the reference handles specially."""
    states[:16] = np.sign(states[:16])  # on the clip bounds
    states[16] = [1.0, 1.0, 1.0, 1.0]
    states[17] = [-1.0, -1.0, -1.0, -1.0]
    states[18] = [1.5, -1.5, 2.0, -2.0]  # outside the Box: clipped by :406-410
    actions[20] = [np.inf, -np.inf]
    actions[21] = [np.nan, 0.1]  # NaN path :415-421
    actions[22] = [0.2, np.nan]
    actions[23] = [5.0, -5.0]
    actions[24] = [1.0, -1.0]
    actions[25] = [-1.0, 1.0]


def gen_step(m) -> None:
    rng = np.random.default_rng(20260101)
    N = 4096
    states = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
    actions = rng.uniform(-1.25, 1.25, (N, 2)).astype(np.float32)
    step_count = rng.integers(0, 402, N).astype(np.int32)
    step_count[:8] = [398, 399, 400, 0, 1, 397, 399, 399]
    edge_rows(states, actions)
    env = m.TwoSeriesCSTREnv()
    env.reset(seed=0)
    obs = np.zeros((N, 4), np.float32)
    rew = np.zeros(N, np.float32)
    trunc = np.zeros(N, bool)
    term = np.zeros(N, bool)
    conc = np.zeros(N, np.float32)
    tpen = np.zeros(N, np.float32)
    for i in range(N):
        env.state = states[i].copy()
        env.current_step = int(step_count[i])
        with contextlib.redirect_stdout(io.StringIO()):
            o, r, te, tr, info = env.step(actions[i].copy())
        obs[i], rew[i], term[i], trunc[i] = o, r, te, tr
        conc[i] = info.get("concentration_reward", np.nan)
        tpen[i] = info.get("temp_penalty", np.nan)
    np.savez_compressed(os.path.join(OUT, "step_f32.npz"), states=states, actions=actions, step_count=step_count,
                        obs=obs, reward=rew, terminated=term, truncated=trunc, conc_reward=conc, temp_penalty=tpen)


def gen_traj(m, core) -> None:
    from core.common.vec_env import DummyVecEnv

    N, T, seed = 8, 405, 100
    venv = DummyVecEnv([(lambda: m.TwoSeriesCSTREnv(init_mode="random")) for _ in range(N)])
    venv.seed(seed)
    obs0 = venv.reset()
    rng = np.random.default_rng(7)
    actions = rng.uniform(-1, 1, (T, N, 2)).astype(np.float32)
    actions[:, 0] = [1.0, -1.0]  # T2 pins at 400 K under this tape
    actions[:, 1] = [-1.0, 1.0]
    obs = np.zeros((T, N, 4), np.float32)
    rew = np.zeros((T, N), np.float32)
    done = np.zeros((T, N), bool)
    timeout = np.zeros((T, N), bool)
    term_obs = np.zeros((T, N, 4), np.float32)
    for t in range(T):
        o, r, d, infos = venv.step(actions[t])
        obs[t], rew[t], done[t] = o, r, d
        for i, info in enumerate(infos):
            timeout[t, i] = info.get("TimeLimit.truncated", False)
            term_obs[t, i] = info["terminal_observation"] if d[i] else o[i]
    np.savez_compressed(os.path.join(OUT, "traj_f32.npz"), seed=seed, obs0=obs0, actions=actions, obs=obs, reward=rew,
                        done=done, timeout=timeout, terminal_obs=term_obs)


def gen_dyn64(m) -> None:
    rng = np.random.default_rng(5)
    N = 2048
    env = m.TwoSeriesCSTREnv()
    lo, hi = env.raw_state_low.astype(np.float64), env.raw_state_high.astype(np.float64)
    raw = lo + (hi - lo) * rng.random((N, 4))
    raw[:8] = np.where(rng.random((8, 4)) < 0.5, lo, hi)
    act = 30.0 + 220.0 * rng.random((N, 2))
    out = np.zeros((N, 4), np.float64)
    for i in range(N):
        res = env._dynamics(state=raw[i].copy(), action=act[i].copy())
        assert res.dtype == np.float64
        out[i] = res
    np.savez_compressed(os.path.join(OUT, "dyn_f64.npz"), raw_state=raw, raw_action=act, new_raw_state=out)


def gen_reset(m) -> None:
    seeds = np.array([0, 1, 7, 42, 123, 2**31 - 1, 99991, 5], np.int64)
    E = 5
    rnd = np.zeros((len(seeds), E, 4), np.float32)
    rnd_raw = np.zeros((len(seeds), E, 4), np.float64)
    sta = np.zeros((len(seeds), E, 4), np.float32)
    for k, s in enumerate(seeds):
        env = m.TwoSeriesCSTREnv(init_mode="random")
        for e in range(E):
            o, info = env.reset(seed=int(s) if e == 0 else None)
            rnd[k, e] = o
            rnd_raw[k, e] = [info["initial_concentration_1"], info["initial_temperature_1"],
                             info["initial_concentration_2"], info["initial_temperature_2"]]
        env = m.TwoSeriesCSTREnv(init_mode="static")
        for e in range(E):
            o, _ = env.reset(seed=int(s) if e == 0 else None)
            sta[k, e] = o
    np.savez_compressed(os.path.join(OUT, "reset.npz"), seeds=seeds, random_obs=rnd, random_raw=rnd_raw, static_obs=sta)


def gen_replay(core) -> None:
    from core.common.buffers import ReplayBuffer
    from gymnasium import spaces

    ospace = spaces.Box(-1, 1, (4,), np.float32)
    aspace = spaces.Box(-1, 1, (2,), np.float32)
    out = {}
    for tag, (size, n_envs, n_add, batch) in {"a": (40, 4, 25, 32), "b": (50, 1, 30, 16), "c": (64, 8, 5, 64)}.items():
        rng = np.random.default_rng(ord(tag) + 3)
        buf = ReplayBuffer(size, ospace, aspace, device="cpu", n_envs=n_envs)
        O_, NO, A, R_, D, TO = [], [], [], [], [], []
        for _ in range(n_add):
            o = rng.uniform(-1, 1, (n_envs, 4)).astype(np.float32)
            no = rng.uniform(-1, 1, (n_envs, 4)).astype(np.float32)
            a = rng.uniform(-1, 1, (n_envs, 2)).astype(np.float32)
            r = rng.normal(size=n_envs).astype(np.float32)
            d = rng.random(n_envs) < 0.3
            to = d & (rng.random(n_envs) < 0.5)
            buf.add(o, no, a, r, d, [{"TimeLimit.truncated": bool(x)} for x in to])
            for lst, v in zip((O_, NO, A, R_, D, TO), (o, no, a, r, d, to)):
                lst.append(v)
        np.random.seed(11)
        s = buf.sample(batch)
        np.random.seed(11)
        upper = buf.buffer_size if buf.full else buf.pos
        bi = np.random.randint(0, upper, size=batch)
        ei = np.random.randint(0, high=n_envs, size=(batch,))
        out.update({
            f"{tag}_cfg": np.array([size, n_envs, n_add, batch]),
            f"{tag}_add_obs": np.stack(O_), f"{tag}_add_next_obs": np.stack(NO), f"{tag}_add_action": np.stack(A),
            f"{tag}_add_reward": np.stack(R_), f"{tag}_add_done": np.stack(D), f"{tag}_add_timeout": np.stack(TO),
            f"{tag}_store_observations": buf.observations, f"{tag}_store_next_observations": buf.next_observations,
            f"{tag}_store_actions": buf.actions, f"{tag}_store_rewards": buf.rewards, f"{tag}_store_dones": buf.dones,
            f"{tag}_store_timeouts": buf.timeouts, f"{tag}_pos": np.array(buf.pos), f"{tag}_full": np.array(buf.full),
            f"{tag}_batch_inds": bi, f"{tag}_env_inds": ei,
            f"{tag}_s_obs": s.observations.numpy(), f"{tag}_s_act": s.actions.numpy(),
            f"{tag}_s_next_obs": s.next_observations.numpy(), f"{tag}_s_dones": s.dones.numpy(),
            f"{tag}_s_rewards": s.rewards.numpy(),
        })
    np.savez_compressed(os.path.join(OUT, "replay.npz"), **out)


def gen_actor(m, core) -> None:
    import torch as th
    from core.common.vec_env import DummyVecEnv

    th.set_num_threads(1)
    venv = DummyVecEnv([lambda: m.TwoSeriesCSTREnv(init_mode="random")])
    model = core.TD3("MlpPolicy", venv, seed=0, device="cpu", buffer_size=1000, learning_starts=0)
    actor = model.policy.actor
    lin = [mod for mod in actor.mu if isinstance(mod, th.nn.Linear)]
    W = {f"W{i + 1}": l.weight.detach().numpy().copy() for i, l in enumerate(lin)}
    Bz = {f"b{i + 1}": l.bias.detach().numpy().copy() for i, l in enumerate(lin)}
    rng = np.random.default_rng(3)
    obs = rng.uniform(-1, 1, (256, 4)).astype(np.float32)
    with th.no_grad():
        mu = actor(th.as_tensor(obs)).numpy()
    pred, _ = model.predict(obs, deterministic=False)
    # _sample_action's maps with an injected noise callable (off_policy_algorithm.py:398-406)
    noise = (0.1 * rng.standard_normal((256, 2))).astype(np.float32)
    noise[:4] = [[3.0, -3.0], [0.0, 0.0], [-0.5, 0.5], [1.0, 1.0]]
    model._last_obs = obs
    model.num_timesteps = 10
    action, buffer_action = model._sample_action(0, action_noise=lambda: noise, n_envs=256)
    np.savez_compressed(os.path.join(OUT, "td3_actor.npz"), obs=obs, mu=mu, predict=pred, noise=noise,
                        env_action=action, buffer_action=buffer_action, **W, **Bz)


def gen_vecnorm(m, core) -> None:
    from core.common.buffers import ReplayBuffer
    from core.common.vec_env import DummyVecEnv, VecNormalize

    N, T, seed = 16, 120, 5
    def make():
        e = m.TwoSeriesCSTREnv(init_mode="random")
        e.max_steps = 50  # plain attribute in the reference (twoseriescstr.py:99): short episodes -> several auto-resets in T steps
        return e

    venv = DummyVecEnv([make for _ in range(N)])
    venv.seed(seed)
    vn = VecNormalize(venv, gamma=0.99)
    buf = ReplayBuffer(128 * N, venv.observation_space, venv.action_space, device="cpu", n_envs=N)
    nobs0 = vn.reset()
    raw0 = vn.get_original_obs()
    rng = np.random.default_rng(23)
    actions = rng.uniform(-1, 1, (T, N, 2)).astype(np.float32)
    keys = ("raw_obs", "raw_rew", "done", "norm_obs", "norm_rew", "obs_mean", "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count", "returns")
    rec = {k: [] for k in keys}
    last_raw = raw0
    for t in range(T):
        o, r, d, infos = vn.step(actions[t])
        raw_o, raw_r = vn.get_original_obs(), vn.get_original_reward()
        nxt = raw_o.copy()
        for i, info in enumerate(infos):
            if d[i]:
                nxt[i] = vn.unnormalize_obs(info["terminal_observation"])  # off_policy_algorithm.py:468-481 stores the ORIGINAL obs
        buf.add(last_raw, nxt, actions[t], raw_r, d, infos)
        last_raw = raw_o
        for k, v in zip(keys, (raw_o, raw_r, d, o, r, vn.obs_rms.mean, vn.obs_rms.var, vn.obs_rms.count, vn.ret_rms.mean, vn.ret_rms.var,
                               vn.ret_rms.count, vn.returns)):
            rec[k].append(np.array(v, copy=True))
    np.random.seed(3)
    s = buf.sample(256, env=vn)
    np.random.seed(3)
    bi = np.random.randint(0, buf.size(), size=256)
    ei = np.random.randint(0, high=N, size=(256,))
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(seed=seed, raw_obs0=raw0, norm_obs0=nobs0, actions=actions, batch_inds=bi, env_inds=ei,
               store_observations=buf.observations[:T], store_next_observations=buf.next_observations[:T], store_actions=buf.actions[:T],
               store_rewards=buf.rewards[:T], store_dones=buf.dones[:T], store_timeouts=buf.timeouts[:T],
               s_obs=s.observations.numpy(), s_act=s.actions.numpy(), s_next_obs=s.next_observations.numpy(), s_dones=s.dones.numpy(),
               s_rewards=s.rewards.numpy())
    np.savez_compressed(os.path.join(OUT, "vecnorm.npz"), **out)


def td3_update_reference_run(m, core, net_arch, K=6, B=64, algo="TD3") -> dict:
    """Run the reference's TD3.train (``algo="DDPG"``: the subclass with one critic, delay 1, no target noise, core/ddpg/ddpg.py:100-109) for K
    gradient steps on CPU torch and return everything needed to replay it."""
    import torch
    from types import SimpleNamespace
    from core.common.vec_env import DummyVecEnv

    torch.set_num_threads(1)
    venv = DummyVecEnv([(lambda: m.TwoSeriesCSTREnv(init_mode="random")) for _ in range(4)])
    model = getattr(core, algo)("MlpPolicy", venv, buffer_size=4000, batch_size=B, learning_starts=0, device="cpu", seed=5,
                                policy_kwargs=dict(net_arch=list(net_arch)))
    logged = {}
    model._logger = SimpleNamespace(record=lambda k, v, **kw: logged.__setitem__(k, v))
    rng = np.random.default_rng(31)
    buf = model.replay_buffer
    for _ in range(200):  # synthetic transitions: the update only sees what sample() returns
        o = rng.uniform(-1, 1, (4, 4)).astype(np.float32)
        no = rng.uniform(-1, 1, (4, 4)).astype(np.float32)
        a = rng.uniform(-1, 1, (4, 2)).astype(np.float32)
        r = rng.normal(-1, 1, 4).astype(np.float32)
        d = rng.random(4) < 0.05
        buf.add(o, no, a, r, d, [{"TimeLimit.truncated": bool(x and rng.random() < 0.5)} for x in d])

    def nets():
        pol = model.policy
        get = lambda seq: [t.detach().numpy().copy() for t in seq.parameters()]  # noqa: E731
        single = len(pol.critic.q_networks) == 1  # DDPG: the second critic slot of the fixture repeats the first
        return {"actor": get(pol.actor.mu), "actor_target": get(pol.actor_target.mu), "critic0": get(pol.critic.q_networks[0]),
                "critic1": get(pol.critic.q_networks[0 if single else 1]), "critic0_target": get(pol.critic_target.q_networks[0]),
                "critic1_target": get(pol.critic_target.q_networks[0 if single else 1])}

    out = {}
    for name, ps in nets().items():
        for i, t in enumerate(ps):
            out[f"init_{name}_{i}"] = t
    np.random.seed(17)
    batches = [buf.sample(B) for _ in range(K)]
    torch.manual_seed(23)
    noise = [torch.empty(B, 2).normal_(0, model.target_policy_noise).numpy().copy() for _ in range(K)]
    np.random.seed(17)
    torch.manual_seed(23)
    model.train(gradient_steps=K, batch_size=B)
    for name, ps in nets().items():
        for i, t in enumerate(ps):
            out[f"final_{name}_{i}"] = t
    for tag, opt in (("actor", model.actor.optimizer), ("critic", model.critic.optimizer)):
        for i, prm in enumerate(opt.param_groups[0]["params"]):
            st = opt.state[prm]
            out[f"adam_{tag}_m_{i}"], out[f"adam_{tag}_v_{i}"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
            out[f"adam_{tag}_step"] = np.array(float(st["step"]))
    for k, f in zip(("obs", "act", "next_obs", "dones", "rewards"), ("observations", "actions", "next_observations", "dones", "rewards")):
        out["batch_" + k] = np.stack([getattr(b, f).numpy() for b in batches])
    out["noise"] = np.stack(noise)
    out["critic_loss_mean"], out["actor_loss_mean"] = np.array(logged["train/critic_loss"]), np.array(logged["train/actor_loss"])
    out["hyper"] = np.array([model.gamma, model.tau, model.policy_delay, model.target_policy_noise, model.target_noise_clip, model.lr_schedule(1)])
    return out


def gen_ddpg_update(m, core) -> None:
    g = td3_update_reference_run(m, core, [64, 48], K=4, algo="DDPG")
    g["n_critics"] = np.array(1)
    np.savez_compressed(os.path.join(OUT, "ddpg_update.npz"), **g)


def gen_td3_update(m, core) -> None:
    # net_arch [64, 48] keeps the fixture small; the restatement is architecture-agnostic (the default [400, 300] is compared live in
    # tests/test_oracle_vs_reference.py when the reference tree is present)
    np.savez_compressed(os.path.join(OUT, "td3_update.npz"), **td3_update_reference_run(m, core, [64, 48]))


def sac_update_reference_run(m, core, net_arch, K=5, B=64) -> dict:
    """Run the reference's SAC.train for K gradient steps on CPU torch and return everything needed to replay it."""
    import torch
    from types import SimpleNamespace
    from core.common.vec_env import DummyVecEnv

    torch.set_num_threads(1)
    venv = DummyVecEnv([(lambda: m.TwoSeriesCSTREnv(init_mode="random")) for _ in range(4)])
    model = core.SAC("MlpPolicy", venv, buffer_size=4000, batch_size=B, learning_starts=0, device="cpu", seed=7, policy_kwargs=dict(net_arch=list(net_arch)))
    logged = {}
    model._logger = SimpleNamespace(record=lambda k, v, **kw: logged.__setitem__(k, v))
    rng = np.random.default_rng(41)
    buf = model.replay_buffer
    for _ in range(200):
        o = rng.uniform(-1, 1, (4, 4)).astype(np.float32)
        no = rng.uniform(-1, 1, (4, 4)).astype(np.float32)
        a = rng.uniform(-1, 1, (4, 2)).astype(np.float32)
        r = rng.normal(-1, 1, 4).astype(np.float32)
        d = rng.random(4) < 0.05
        buf.add(o, no, a, r, d, [{"TimeLimit.truncated": False} for _ in d])

    def nets():
        pol = model.policy
        lat = [t.detach().numpy().copy() for t in pol.actor.latent_pi.parameters()]
        head_w = np.concatenate([pol.actor.mu.weight.detach().numpy(), pol.actor.log_std.weight.detach().numpy()], 0)
        head_b = np.concatenate([pol.actor.mu.bias.detach().numpy(), pol.actor.log_std.bias.detach().numpy()], 0)
        get = lambda seq: [t.detach().numpy().copy() for t in seq.parameters()]  # noqa: E731
        return {"actor": lat + [head_w, head_b], "critic0": get(pol.critic.q_networks[0]), "critic1": get(pol.critic.q_networks[1]),
                "critic0_target": get(pol.critic_target.q_networks[0]), "critic1_target": get(pol.critic_target.q_networks[1])}

    out = {}
    for name, ps in nets().items():
        for i, t in enumerate(ps):
            out[f"init_{name}_{i}"] = t
    out["init_log_ent_coef"] = model.log_ent_coef.detach().numpy().copy()
    np.random.seed(19)
    batches = [buf.sample(B) for _ in range(K)]
    torch.manual_seed(29)
    eps = [[torch.empty(B, 2).normal_().numpy().copy() for _ in range(2)] for _ in range(K)]  # Normal.rsample: one standard-normal draw each
    np.random.seed(19)
    torch.manual_seed(29)
    model.train(gradient_steps=K, batch_size=B)
    for name, ps in nets().items():
        for i, t in enumerate(ps):
            out[f"final_{name}_{i}"] = t
    out["final_log_ent_coef"] = model.log_ent_coef.detach().numpy().copy()
    for k, f in zip(("obs", "act", "next_obs", "dones", "rewards"), ("observations", "actions", "next_observations", "dones", "rewards")):
        out["batch_" + k] = np.stack([getattr(b, f).numpy() for b in batches])
    out["eps_pi"], out["eps_next"] = np.stack([e[0] for e in eps]), np.stack([e[1] for e in eps])
    for k in ("critic_loss", "actor_loss", "ent_coef", "ent_coef_loss"):
        out[k + "_mean"] = np.array(logged["train/" + k])
    out["hyper"] = np.array([model.gamma, model.tau, model.target_entropy, model.lr_schedule(1), model.target_update_interval])
    return out


def gen_sac_update(m, core) -> None:
    np.savez_compressed(os.path.join(OUT, "sac_update.npz"), **sac_update_reference_run(m, core, [64, 48]))


def bcq_update_reference_run(m, core, K=5, B=32, arch=None, critic_arch=(36, 20)) -> dict:
    """Run the reference's BCQ.train for K gradient steps on CPU torch, recording every random draw, and return what replays it."""
    import pickle
    import tempfile

    import torch
    from types import SimpleNamespace
    from core.common.buffers import ReplayBuffer
    from core.common.vec_env import DummyVecEnv

    torch.set_num_threads(1)
    arch = dict(arch or dict(vae_latent_dim=8, vae_hidden_dim=48, perturbation_hidden_dim=40, max_perturbation=0.05))
    venv = DummyVecEnv([lambda: m.TwoSeriesCSTREnv(init_mode="random")])
    rng = np.random.default_rng(51)
    n = 600
    data = ReplayBuffer(n, venv.observation_space, venv.action_space, device="cpu", n_envs=1)
    data.observations[:, 0] = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    data.next_observations[:, 0] = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
    data.actions[:, 0] = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    data.rewards[:, 0] = rng.normal(-1, 1, n).astype(np.float32)
    data.dones[:, 0] = (rng.random(n) < 0.05).astype(np.float32)
    data.full, data.pos = True, 0
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "dataset.pkl")
        with open(path, "wb") as fh:
            pickle.dump(data, fh)
        model = core.BCQ("MlpPolicy", venv, dataset=path, batch_size=B, device="cpu", seed=9,
                         policy_kwargs=dict(actor_net_arch=arch, critic_net_arch=list(critic_arch)))
    logged = {}
    model._logger = SimpleNamespace(record=lambda k, v, **kw: logged.__setitem__(k, v))

    def nets():
        pol = model.policy
        get = lambda seq: [t.detach().numpy().copy() for t in seq.parameters()]  # noqa: E731

        def vae(v):
            enc = get(v.encoder)
            head_w = np.concatenate([v.mean.weight.detach().numpy(), v.log_std.weight.detach().numpy()], 0)
            head_b = np.concatenate([v.mean.bias.detach().numpy(), v.log_std.bias.detach().numpy()], 0)
            return enc + [head_w, head_b], get(v.decoder)

        enc, dec = vae(pol.actor.vae)
        enc_t, dec_t = vae(pol.actor_target.vae)
        return {"vae_enc": enc, "vae_dec": dec, "pert": get(pol.actor.perturbation.model), "critic0": get(pol.critic.q_networks[0]),
                "critic1": get(pol.critic.q_networks[1]), "vae_enc_target": enc_t, "vae_dec_target": dec_t,
                "pert_target": get(pol.actor_target.perturbation.model), "critic0_target": get(pol.critic_target.q_networks[0]),
                "critic1_target": get(pol.critic_target.q_networks[1])}

    out = {}
    for name, ps in nets().items():
        for i, t in enumerate(ps):
            out[f"init_{name}_{i}"] = t
    np.random.seed(23)
    batches = [model.replay_buffer.sample(B) for _ in range(K)]
    draws = []
    orig_randn, orig_like = torch.randn, torch.randn_like

    def rec_randn(*a, **k):
        t = orig_randn(*a, **k)
        draws.append(t.numpy().copy())
        return t

    def rec_like(x, **k):
        t = orig_like(x, **k)
        draws.append(t.numpy().copy())
        return t

    np.random.seed(23)
    torch.manual_seed(33)
    torch.randn, torch.randn_like = rec_randn, rec_like
    try:
        model.train(gradient_steps=K, batch_size=B)
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like
    for name, ps in nets().items():
        for i, t in enumerate(ps):
            out[f"final_{name}_{i}"] = t
    for k, f in zip(("obs", "act", "next_obs", "dones", "rewards"), ("observations", "actions", "next_observations", "dones", "rewards")):
        out["batch_" + k] = np.stack([getattr(b, f).numpy() for b in batches])
    # draw order per gradient step: randn_like (B, L); randn (10 B, L); on actor steps randn (B, L)
    it = iter(draws)
    L = arch["vae_latent_dim"]
    eps_vae, z_next, z_actor = [], [], []
    for k in range(1, K + 1):
        eps_vae.append(next(it))
        z_next.append(next(it))
        z_actor.append(next(it) if k % model.actor_delay == 0 else np.zeros((B, L), np.float32))
    assert next(it, None) is None and eps_vae[0].shape == (B, L) and z_next[0].shape == (10 * B, L)
    out["eps_vae"], out["z_next"], out["z_actor"] = np.stack(eps_vae), np.stack(z_next), np.stack(z_actor)
    for k in ("critic_loss", "actor_loss", "vae_loss"):
        out[k + "_mean"] = np.array(logged["train/" + k])
    out["hyper"] = np.array([model.gamma, model.tau, arch["max_perturbation"], model.lr_schedule(1), model.actor_delay])
    return out


def gen_bcq_update(m, core) -> None:
    np.savez_compressed(os.path.join(OUT, "bcq_update.npz"), **bcq_update_reference_run(m, core))


def multi_agent_reference_run(m, core, algo="MADDPG", K=4, B=32, arch=(24, 16)) -> dict:
    """Run the reference's MADDPG.train / IDDPG.train (two agents = the two reactors) for K gradient steps on CPU torch."""
    import torch
    from types import SimpleNamespace
    from core.common.vec_env import DummyVecEnv

    torch.set_num_threads(1)
    venv = DummyVecEnv([(lambda: m.TwoSeriesCSTREnv(init_mode="random")) for _ in range(4)])
    splits = dict(n_agents=2, observation_splits=[[0, 1], [2, 3]], action_splits=[[0], [1]], learning_rate_list=[1e-3, 5e-4])
    model = getattr(core, algo)(policy="MlpPolicy", env=venv, buffer_size=4000, batch_size=B, learning_starts=0, device="cpu", seed=13,
                                policy_kwargs=dict(net_arch=[list(arch), list(arch)]), **splits)
    logged = {}
    model._logger = SimpleNamespace(record=lambda k, v, **kw: logged.__setitem__(k, v))
    rng = np.random.default_rng(61)
    buf = model.replay_buffer
    for _ in range(200):
        d = rng.random(4) < 0.05
        buf.add(rng.uniform(-1, 1, (4, 4)).astype(np.float32), rng.uniform(-1, 1, (4, 4)).astype(np.float32), rng.uniform(-1, 1, (4, 2)).astype(np.float32),
                rng.normal(-1, 1, 4).astype(np.float32), d, [{"TimeLimit.truncated": False} for _ in d])

    def nets():
        pol = model.policy
        get = lambda seq: [t.detach().numpy().copy() for t in seq.parameters()]  # noqa: E731
        out = {}
        for i in range(2):
            out[f"actor{i}"], out[f"actor{i}_target"] = get(pol.actor.mu_list[i]), get(pol.actor_target.mu_list[i])
            for k in range(2):
                out[f"critic{i}_{k}"] = get(pol.critic.q_networks_list[i][k])
                out[f"critic{i}_{k}_target"] = get(pol.critic_target.q_networks_list[i][k])
        return out

    out = {}
    for name, ps in nets().items():
        for j, t in enumerate(ps):
            out[f"init_{name}_{j}"] = t
    np.random.seed(27)
    batches = [buf.sample(B) for _ in range(K)]
    torch.manual_seed(37)
    noise = [[torch.empty(B, 1).normal_(0, model.target_policy_noise).numpy().copy() for _ in range(2)] for _ in range(K)]  # :139, agent order
    np.random.seed(27)
    torch.manual_seed(37)
    model.train(gradient_steps=K, batch_size=B)
    for name, ps in nets().items():
        for j, t in enumerate(ps):
            out[f"final_{name}_{j}"] = t
    for k, f in zip(("obs", "act", "next_obs", "dones", "rewards"), ("observations", "actions", "next_observations", "dones", "rewards")):
        out["batch_" + k] = np.stack([getattr(b, f).numpy() for b in batches])
    out["noise"] = np.asarray(noise)  # (K, agent, B, 1)
    for i in range(2):
        out[f"critic_loss_mean_{i}"] = np.array(logged[f"train/agent_{i}_critic_loss"])
        out[f"actor_loss_mean_{i}"] = np.array(logged[f"train/agent_{i}_actor_loss"])
    out["hyper"] = np.array([model.gamma, model.tau, model.policy_delay, model.target_noise_clip, 1e-3, 5e-4])
    return out


def gen_multi_agent_update(m, core) -> None:
    np.savez_compressed(os.path.join(OUT, "maddpg_update.npz"), **multi_agent_reference_run(m, core, "MADDPG"))
    np.savez_compressed(os.path.join(OUT, "iddpg_update.npz"), **multi_agent_reference_run(m, core, "IDDPG"))


def main() -> None:
    if not refload.available():
        raise SystemExit("reference tree not found; fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    m = refload.load_env_module()
    core = refload.load_core()
    gen_step(m)
    gen_traj(m, core)
    gen_dyn64(m)
    gen_reset(m)
    gen_replay(core)
    gen_actor(m, core)
    gen_vecnorm(m, core)
    gen_td3_update(m, core)
    gen_ddpg_update(m, core)
    gen_sac_update(m, core)
    gen_bcq_update(m, core)
    gen_multi_agent_update(m, core)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
