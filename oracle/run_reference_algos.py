"""Integration run (TEST INFRASTRUCTURE): the UNMODIFIED reference algorithms on top of the GPU classes.

Needs a CUDA device AND the reference tree (``CSTR_REFERENCE_ROOT``, default /root/reference).  Those two never
coexist in the build setup (the build container has no GPU, the GPU box has no reference), so this is not a pytest:
it is run by hand through ``gpurun`` with the reference staged in the git-ignored ``baseline/_ref`` for the duration
of the call, and its log is committed as ``profiles/rNN_reference_algos.log``.

What it shows: ``core.TD3 / SAC / IDDPG / MADDPG .learn()`` and ``core.BCQ.learn()`` run with
``env=GpuCSTRVecEnv`` (bound to the reference's ``VecEnv``) and ``replay_buffer_class=GpuReplayBuffer`` — i.e.
``collect_rollouts``, ``_store_transition``, ``_update_info_buffer``, ``train()`` and ``evaluate_policy`` are the
reference's own code, and every env step / buffer add / buffer sample underneath is a CUDA kernel.
"""
from __future__ import annotations

import importlib
import os
import pickle
import sys
import tempfile
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
for p in (ROOT, _HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

import refload  # noqa: E402


def main() -> None:
    import torch

    if not refload.available():
        raise SystemExit("reference tree not found (set CSTR_REFERENCE_ROOT)")
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    refload.install_shims()
    core = refload.load_core()
    ref_env_mod = refload.load_env_module()
    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    from core.common.buffers import ReplayBuffer
    from core.common.evaluation import evaluate_policy
    from core.common.noise import NormalActionNoise
    from core.common.vec_env import DummyVecEnv, VecEnv

    GpuVecEnv = pkg.bind_vec_env_class(VecEnv)
    GpuBuffer = pkg.bind_replay_buffer_class(ReplayBuffer)
    torch.set_num_threads(4)
    launches = {}

    def noise():
        return NormalActionNoise(mean=np.zeros(2), sigma=0.1 * np.ones(2))

    def run(name, make, steps, **learn_kw):
        env = GpuVecEnv(num_envs=16, init_mode="static", reset_rng="pcg64", seed=42)
        assert isinstance(env, VecEnv)
        t0 = time.time()
        model = make(env)
        model.learn(total_timesteps=steps, **learn_kw)
        dt = time.time() - t0
        buf = model.replay_buffer
        assert isinstance(buf, pkg.GpuReplayBuffer) and isinstance(buf, ReplayBuffer), type(buf)
        ep = [e["r"] for e in model.ep_info_buffer]
        eval_env = GpuVecEnv(num_envs=1, init_mode="static", reset_rng="pcg64", seed=7)
        mean_r, std_r = evaluate_policy(model, eval_env, n_eval_episodes=2, warn=False)
        launches[name] = getattr(env, "launches", 0) + buf.launches
        print(f"[{name}] {steps} timesteps in {dt:.1f} s | buffer size {buf.size()} rows x {buf.n_envs} envs, pos {buf.pos} | "
              f"episodes logged {len(ep)} (mean return {np.mean(ep) if ep else float('nan'):.1f}) | eval return {mean_r:.1f} +- {std_r:.1f} | "
              f"CUDA launches: env {env.unwrapped.launches if hasattr(env, 'unwrapped') else env.launches}, buffer {buf.launches}", flush=True)
        return model

    common = dict(replay_buffer_class=GpuBuffer, buffer_size=64_000, learning_starts=800, batch_size=256, device="cuda", seed=42)
    run("TD3", lambda env: core.TD3("MlpPolicy", env, action_noise=noise(), train_freq=(1, "step"), gradient_steps=4, **common), 8000)
    run("SAC", lambda env: core.SAC("MlpPolicy", env, train_freq=(1, "step"), gradient_steps=4, **common), 6400)
    ma = dict(n_agents=2, observation_splits=[[0, 1], [2, 3]], action_splits=[[0], [1]], learning_rate_list=[1e-3, 1e-3])
    run("MADDPG", lambda env: core.MADDPG(policy="MlpPolicy", env=env, train_freq=(1, "step"), gradient_steps=2, **ma, **common), 4800)
    run("IDDPG", lambda env: core.IDDPG(policy="MlpPolicy", env=env, train_freq=(1, "step"), gradient_steps=2, **ma, **common), 4800)

    # ---- TD3 under VecNormalize: statistics, normalisation and the normalised sample all on the device ---------------
    from core.common.vec_env import VecNormalize, unwrap_vec_normalize

    GpuNorm = pkg.bind_vec_normalize_class(VecNormalize)

    def make_norm(env):
        wrapped = GpuNorm(env, gamma=0.99)
        assert isinstance(wrapped, VecNormalize) and unwrap_vec_normalize(wrapped) is wrapped
        model = core.TD3("MlpPolicy", wrapped, action_noise=noise(), train_freq=(1, "step"), gradient_steps=4, **common)
        assert model.get_vec_normalize_env() is wrapped
        return model

    m = run("TD3+VecNormalize", make_norm, 8000)
    vn = m.get_vec_normalize_env()
    print(f"[TD3+VecNormalize] obs_rms.count {vn.obs_rms.count:.1f} mean {np.round(vn.obs_rms.mean, 3)} var {np.round(vn.obs_rms.var, 4)} | "
          f"ret_rms.var {float(vn.ret_rms.var):.3f} | normaliser launches {vn.launches}", flush=True)
    with tempfile.TemporaryDirectory() as tmp:  # VecNormalize.save/load interchange with the reference class
        vn.save(os.path.join(tmp, "vn.pkl"))
        back = pkg.GpuVecNormalize.load(os.path.join(tmp, "vn.pkl"), GpuVecEnv(num_envs=16, init_mode="static", reset_rng="pcg64", seed=1))
        assert np.array_equal(back.obs_rms.var, vn.obs_rms.var)
        ref_vn = vn.to_reference(VecNormalize, DummyVecEnv([lambda: ref_env_mod.TwoSeriesCSTREnv(init_mode="static")]))
        probe = np.random.default_rng(0).uniform(-1, 1, (5, 4)).astype(np.float32)
        assert np.array_equal(ref_vn.normalize_obs(probe), vn.normalize_obs(probe))
        ref_vn.save(os.path.join(tmp, "ref.pkl"))
        adopted = pkg.GpuVecNormalize.load(os.path.join(tmp, "ref.pkl"), GpuVecEnv(num_envs=16, init_mode="static", reset_rng="pcg64", seed=1))
        assert np.array_equal(adopted.normalize_obs(probe), vn.normalize_obs(probe))
        print("[TD3+VecNormalize] save/load + to_reference/from_reference: normalised values identical on both sides", flush=True)

    # ---- TD3 with the gradient steps on the device too (bind_td3_class): learn() is the reference's, train() is cstr_td3_update ----
    FusedTD3 = pkg.bind_td3_class(core.TD3)
    m = run("TD3 fused update", lambda env: FusedTD3("MlpPolicy", env, action_noise=noise(), train_freq=(1, "step"), gradient_steps=4, **common), 8000)
    assert isinstance(m, core.TD3) and m._fused is not None and m._n_updates == m._fused.n_updates
    assert m.policy.actor.mu[0].weight.data_ptr() == m._fused.views("params")["actor"][0].data_ptr()
    with tempfile.TemporaryDirectory() as tmp:  # save/load through the reference's zip format, Adam state included
        m.save(os.path.join(tmp, "td3"))
        back = core.TD3.load(os.path.join(tmp, "td3"), device="cuda")
        assert torch.equal(back.policy.actor.mu[2].weight, m.policy.actor.mu[2].weight)
        st = back.critic.optimizer.state_dict()["state"]
        assert len(st) == 12 and float(st[0]["step"]) == m._fused.critic_step
    print(f"[TD3 fused update] n_updates {m._n_updates}, kernel launches {m._fused.launches}; model.save -> TD3.load round trip ok", flush=True)

    # ---- DDPG (TD3 with one critic, core/ddpg/ddpg.py) through the same fused update ----
    FusedDDPG = pkg.bind_td3_class(core.DDPG)
    m = run("DDPG fused update", lambda env: FusedDDPG("MlpPolicy", env, action_noise=noise(), train_freq=(1, "step"), gradient_steps=4, **common), 6400)
    assert isinstance(m, core.DDPG) and m._fused is not None and m._fused.n_critics == 1 and m._fused.policy_delay == 1

    # ---- SAC with the gradient steps on the device (bind_sac_class) ----
    FusedSAC = pkg.bind_sac_class(core.SAC)
    m = run("SAC fused update", lambda env: FusedSAC("MlpPolicy", env, train_freq=(1, "step"), gradient_steps=4, **common), 6400)
    assert isinstance(m, core.SAC) and m._fused is not None and m._n_updates == m._fused.n_updates
    assert m.policy.actor.mu.weight.data_ptr() == m._fused.views("params")["actor"][4].data_ptr() and m.log_ent_coef.data_ptr() == m._fused.log_ent_coef.data_ptr()
    with tempfile.TemporaryDirectory() as tmp:  # save -> SAC.load (plain reference class): weights, log_ent_coef and all three Adam states
        m.save(os.path.join(tmp, "sac"))
        back = core.SAC.load(os.path.join(tmp, "sac"), device="cuda")
        assert torch.equal(back.policy.actor.log_std.weight, m.policy.actor.log_std.weight) and torch.equal(back.log_ent_coef, m.log_ent_coef)
        for opt, count in ((back.actor.optimizer, 8), (back.critic.optimizer, 12), (back.ent_coef_optimizer, 1)):
            st = opt.state_dict()["state"]
            assert len(st) == count and float(st[0]["step"]) == m._fused.critic_step == m._n_updates
        assert torch.equal(back.actor.optimizer.state_dict()["state"][5]["exp_avg"], m._fused.views("adam_m")["actor"][5][0:2])  # mu.bias moments
        # ... and back into a fused model: the engine takes the moments and the step count over and keeps going
        again = FusedSAC.load(os.path.join(tmp, "sac"), env=m.get_env(), device="cuda")
        again.learn(total_timesteps=64, reset_num_timesteps=False)
        assert again._fused is not None and again._fused.critic_step == again._n_updates > m._n_updates
    print(f"[SAC fused update] n_updates {m._n_updates}, ent_coef {float(m.log_ent_coef.detach().exp()):.4f}, kernel launches {m._fused.launches}; "
          f"model.save -> SAC.load -> fused continue ok", flush=True)

    # ---- BCQ offline: dataset generated by the GPU tape kernel, handed over in the reference's pickle format -------
    n, T = 2500, 400
    env = pkg.GpuCSTRVecEnv(n, seed=3, monitor=False)
    env.reset()
    obs0 = env.state.clone()
    res = env.tape(T, None, want_obs=True)
    obs = torch.cat([obs0[None], res["obs"][:-1]], 0)
    # next_obs of the truncation row is the terminal observation, not the post-reset one: drop that row for simplicity
    keep = torch.arange(T, device=obs.device) != 399
    import cstr_oracle as O

    ids = np.arange(n, dtype=np.uint64)
    acts = np.zeros((T, n, 2), np.float32)
    for t in range(T):
        ctr = np.stack([ids.astype(np.uint32), np.zeros(n, np.uint32), np.full(n, t >> 1, np.uint32), np.full(n, 2 << 8, np.uint32)], 1)
        r = O.philox4x32(ctr, np.array([3, 0], np.uint32))
        acts[t] = O.u32_to_unit_f32(r[:, 2:4] if t & 1 else r[:, 0:2]) * np.float32(2) - np.float32(1)
    size = int(keep.sum().item()) * n
    ref_buf = ReplayBuffer(size, env.observation_space, env.action_space, device="cpu", n_envs=1)
    o_np = obs[keep].reshape(-1, 4).cpu().numpy()
    no_np = res["obs"][keep].reshape(-1, 4).cpu().numpy()
    ref_buf.observations[:, 0] = o_np
    ref_buf.next_observations[:, 0] = no_np
    ref_buf.actions[:, 0] = acts[keep.cpu().numpy()].reshape(-1, 2)
    ref_buf.rewards[:, 0] = res["rewards"][keep].reshape(-1).cpu().numpy()
    ref_buf.full, ref_buf.pos = True, 0
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "cstr_dataset.pkl")
        with open(path, "wb") as fh:
            pickle.dump(ref_buf, fh)
        t0 = time.time()
        single = DummyVecEnv([lambda: pkg.TwoSeriesCSTREnv(init_mode="static")])  # the single-reactor gym.Env façade
        model = core.BCQ("MlpPolicy", single, dataset=path, batch_size=256, device="cuda", seed=0)
        model.replay_buffer = GpuBuffer.from_reference(model.replay_buffer)  # Q8: the loaded buffer is adopted by reference
        model.learn(total_timesteps=300)
        s = model.replay_buffer.sample(4)
        mean_r, std_r = evaluate_policy(model, single, n_eval_episodes=1, warn=False)
        print(f"[BCQ] dataset {size} transitions generated by cstr_tape_f32 -> reference pickle -> GpuReplayBuffer.from_reference; "
              f"300 updates in {time.time() - t0:.1f} s, buffer launches {model.replay_buffer.launches}, sample on {s.observations.device}, "
              f"eval return on the gym.Env façade {mean_r:.1f}", flush=True)
    print("reference algorithms ran unchanged on the GPU env + GPU replay buffer: OK")


if __name__ == "__main__":
    main()
