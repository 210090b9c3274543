"""Integration runs (TEST INFRASTRUCTURE): the reference's own ``learn()`` on the GPU path, beside the unmodified reference.

Needs a CUDA device AND the reference tree (``oracle/refload.py``: ``CSTR_REFERENCE_ROOT``, /root/reference, or the git-ignored staging
copy ``baseline/_ref`` that travels to the GPU box).  Run through ``gpurun``; the log is committed under ``profiles/``.

  config1   BASELINE.json configs[0] exactly as ``experiments/basic_test/TwoSeriesCSTR_TD3.py:57-75`` writes it (n_envs=1, lr 3e-4,
            buffer 1e5, learning_starts 5000, batch 256, train_freq (1, "step"), gradient_steps 1, NormalActionNoise sigma 0.1,
            policy_delay 2, seed 42) with total_timesteps=10,000: wall time and the 10-episode evaluation return of
              (a) the unmodified reference on the host CPU (DummyVecEnv of its own TwoSeriesCSTREnv, its own ReplayBuffer, torch CPU),
              (b) the reference's learn()/collect_rollouts/train() on GpuCSTRVecEnv + GpuReplayBuffer (NumPy protocol per step),
              (c) the same learn() with the fused update AND the fused rollout bound in (no per-env Python, nothing leaves the GPU).
  seeds     eval return of the fused update (bind_td3_class) against the reference's torch train() over 5 seeds, same config, same
            env/buffer classes underneath — the two differ only in the source of the random draws (Philox vs torch/NumPy global RNG).
  config3   core.TD3.learn() through the fused rollout at 131,072 reactors (one GPU's share of config #3) with the tcgen05 actor.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
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

    ap = argparse.ArgumentParser()
    ap.add_argument("--what", nargs="+", default=["config1", "seeds", "config3"])
    ap.add_argument("--seeds", type=int, default=5)
    ap.add_argument("--timesteps", type=int, default=10_000)
    args = ap.parse_args()
    if not refload.available():
        raise SystemExit("reference tree not found")
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
    FusedTD3 = pkg.bind_td3_class(core.TD3)
    FullyFusedTD3 = pkg.bind_offpolicy_rollout(FusedTD3)
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))

    def noise():
        return NormalActionNoise(mean=np.zeros(2), sigma=0.1 * np.ones(2))

    script = dict(learning_rate=3e-4, buffer_size=int(1e5), learning_starts=5000, batch_size=256, tau=0.005, gamma=0.99, train_freq=(1, "step"),
                  gradient_steps=1, policy_delay=2, target_policy_noise=0.2, target_noise_clip=0.5, verbose=0, seed=42)

    def evaluate(model, episodes=10, seed=7):
        env = GpuVecEnv(num_envs=1, init_mode="static", reset_rng="pcg64", seed=seed)
        mean_r, std_r = evaluate_policy(model, env, n_eval_episodes=episodes, warn=False)
        return float(mean_r), float(std_r)

    if "config1" in args.what:
        rows = []
        # (a) the unmodified reference, CPU
        t0 = time.time()
        env = DummyVecEnv([lambda: ref_env_mod.TwoSeriesCSTREnv(init_mode="static")])
        m = core.TD3("MlpPolicy", env, action_noise=noise(), device="cpu", **script)
        m.learn(total_timesteps=args.timesteps)
        rows.append(("reference (CPU: DummyVecEnv + ReplayBuffer + torch cpu)", time.time() - t0, *evaluate(m), m._n_updates, None))
        # (b) reference learn()/collect_rollouts/train() over the GPU env + buffer, NumPy protocol
        t0 = time.time()
        env = GpuVecEnv(num_envs=1, init_mode="static", reset_rng="pcg64", seed=42)
        m = core.TD3("MlpPolicy", env, action_noise=noise(), device="cuda", replay_buffer_class=GpuBuffer, **script)
        m.learn(total_timesteps=args.timesteps)
        rows.append(("reference learn() on GpuCSTRVecEnv + GpuReplayBuffer (per-step NumPy protocol, torch cuda train)", time.time() - t0, *evaluate(m),
                     m._n_updates, env.launches + m.replay_buffer.launches))
        # (c) fused update + fused rollout under the same learn()
        t0 = time.time()
        env = GpuVecEnv(num_envs=1, init_mode="static", reset_rng="philox", seed=42)
        m = FullyFusedTD3("MlpPolicy", env, action_noise=noise(), device="cuda", replay_buffer_class=GpuBuffer,
                          replay_buffer_kwargs=dict(index_mode="philox", seed=42), **script)
        m.learn(total_timesteps=args.timesteps)
        rows.append(("reference learn() with bind_offpolicy_rollout(bind_td3_class(TD3)): fused rollout + fused update", time.time() - t0, *evaluate(m),
                     m._n_updates, m.fused_rollout_launches + m._fused.launches))
        assert m._fused is not None and m.fused_rollout_launches == args.timesteps and m.num_timesteps == args.timesteps
        print(f"== config #1: TD3 MlpPolicy, single reactor, {args.timesteps} timesteps (TwoSeriesCSTR_TD3.py:57-75) ==")
        for name, dt, mr, sr, nu, launches in rows:
            print(f"  {name}\n      wall {dt:7.1f} s | eval return (10 episodes) {mr:8.1f} +- {sr:5.1f} | n_updates {nu} | kernel launches {launches}", flush=True)

    if "seeds" in args.what:
        common = dict(replay_buffer_class=GpuBuffer, buffer_size=64_000, learning_starts=800, batch_size=256, device="cuda", train_freq=(1, "step"),
                      gradient_steps=4, verbose=0)
        out = {"reference_train": [], "fused_update": [], "fused_update_and_rollout": []}
        for seed in range(args.seeds):
            for name, cls, rng_mode, extra in (("reference_train", core.TD3, "pcg64", {}), ("fused_update", FusedTD3, "pcg64", {}),
                                               ("fused_update_and_rollout", FullyFusedTD3, "philox",
                                                dict(replay_buffer_kwargs=dict(index_mode="philox", seed=seed)))):
                env = GpuVecEnv(num_envs=16, init_mode="static", reset_rng=rng_mode, seed=seed)
                t0 = time.time()
                m = cls("MlpPolicy", env, action_noise=noise(), seed=seed, **common, **extra)
                m.learn(total_timesteps=8000)
                mr, sr = evaluate(m, episodes=5, seed=1000 + seed)
                ep = [e["r"] for e in m.ep_info_buffer]
                out[name].append(dict(seed=seed, eval_return=mr, train_ep_mean=float(np.mean(ep)) if ep else None, wall_s=time.time() - t0))
                print(f"  seed {seed} {name:26s} eval {mr:8.1f} | last training episodes {np.mean(ep) if ep else float('nan'):8.1f} | {time.time() - t0:5.1f} s",
                      flush=True)
        print("== TD3, 16 reactors, 8000 timesteps, learning_starts 800, gradient_steps 4: evaluation return over seeds ==")
        for name, rows in out.items():
            r = np.array([x["eval_return"] for x in rows])
            print(f"  {name:26s} mean {r.mean():8.1f}  median {np.median(r):8.1f}  min {r.min():8.1f}  max {r.max():8.1f}  ({', '.join(f'{v:.0f}' for v in r)})")
        print(json.dumps(out))

    if "config3" in args.what:
        n = 131_072
        TcTD3 = pkg.bind_offpolicy_rollout(FusedTD3, actor_mode="tc")
        env = GpuVecEnv(num_envs=n, init_mode="random", reset_rng="philox", seed=0, monitor=False)
        rows_ring = 64
        m = TcTD3("MlpPolicy", env, action_noise=noise(), device="cuda", replay_buffer_class=GpuBuffer, buffer_size=rows_ring * n,
                  replay_buffer_kwargs=dict(index_mode="philox", seed=0), learning_starts=16 * n, batch_size=4096, train_freq=(16, "step"),
                  gradient_steps=16, learning_rate=1e-3, verbose=0, seed=0)
        total = 16 * n * 60
        t0 = time.time()
        m.learn(total_timesteps=total, log_interval=None)
        torch.cuda.synchronize()
        dt = time.time() - t0
        ep = np.array([e["r"] for e in m.ep_info_buffer])
        print(f"== config #3 share of one GPU through core.TD3.learn(): {n} reactors, tcgen05 fused rollout (16 steps / launch) + fused update "
              f"(16 x batch 4096 per launch, CUDA graph) ==\n  {total} timesteps in {dt:.2f} s = {total / dt:.3e} transitions/s incl. updates | "
              f"rollout launches {m.fused_rollout_launches} | n_updates {m._n_updates} | episodes {m._episode_num} | "
              f"last-100 episode return {ep.mean() if len(ep) else float('nan'):.1f}", flush=True)
        assert m.num_timesteps == total and m._fused is not None and m.fused_rollout_launches >= 60


if __name__ == "__main__":
    main()
