#!/usr/bin/env python
"""SAC end to end on the device (BASELINE.json configs[2], the SAC half): squashed-Gaussian actor inference fused with the CSTR step
and the replay write (``FusedRollout``, actor kind "gaussian", tcgen05 hidden layer), Philox replay sampling, and the SAC gradient
step (``FusedSACUpdate`` = cstr_sac_update: entropy coefficient, soft target, critics, actor, polyak) — no torch autograd anywhere.

    python examples/sac_fused_rollout.py                      # 16,384 reactors, 30 episodes, 48,000 updates: learns in ~10 s
    python examples/sac_fused_rollout.py --n-envs 131072 --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096   # throughput shape
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--steps-per-iter", type=int, default=4)
    ap.add_argument("--updates-per-iter", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--actor-mode", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(args.seed)
    n = args.n_envs
    env = pkg.GpuCSTRVecEnv(n, device=dev, seed=args.seed, monitor=False)
    buf = pkg.GpuReplayBuffer(args.rows * n, device=dev, n_envs=n, index_mode="philox", seed=args.seed)
    eng = pkg.FusedSACUpdate([256, 256], args.batch, device=dev, seed=args.seed)

    def mlp(i, o):  # torch nn.Linear default init
        out = []
        for fi, fo in ((i, 256), (256, 256), (256, o)):
            b = 1.0 / np.sqrt(fi)
            out += [rng.uniform(-b, b, (fo, fi)).astype(np.float32), rng.uniform(-b, b, fo).astype(np.float32)]
        return out

    eng.load_nets({"actor": mlp(4, 4), "critic0": mlp(6, 1), "critic1": mlp(6, 1)})
    actor_views = eng.views("params")["actor"]
    weights = pkg.ActorWeights(*actor_views, device=dev, kind="gaussian")
    roll = pkg.FusedRollout(env, buf, weights, sigma=0.0, actor_mode=args.actor_mode)
    env.reset()
    roll.collect(args.steps_per_iter, warmup=True)
    rsum = torch.zeros(1, dtype=torch.float64, device=dev)
    log = []
    torch.cuda.synchronize()
    t0 = time.time()
    window = max(1, 400 // args.steps_per_iter)  # iterations per 400-step episode: report whole episodes
    for it in range(args.iters):
        roll.collect(args.steps_per_iter, reward_sum=rsum)
        eng.train(args.updates_per_iter, buf, args.batch, graph=True)  # one CUDA-graph replay per update once the ring is full
        weights.refresh_from_tensors(actor_views)  # device-to-device; repacks the bf16 UMMA image of W2
        if (it + 1) % window == 0 or it == args.iters - 1:
            mean_r = float(rsum.item()) / (n * args.steps_per_iter * ((it % window) + 1))
            rsum.zero_()
            critic_loss, actor_loss, _, ent_coef = eng.pop_losses()
            log.append(mean_r)
            print(f"episode {len(log):3d} (iter {it:4d})  mean reward/step {mean_r:8.4f}  episode return {400 * mean_r:8.1f}  critic loss {critic_loss:.4f}  "
                  f"actor loss {actor_loss:.4f}  ent_coef {ent_coef:.4f}", flush=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    transitions = args.iters * args.steps_per_iter * n
    print(json.dumps({"n_envs": n, "transitions": transitions, "seconds": dt, "transitions_per_s_incl_updates": transitions / dt,
                      "updates": eng.n_updates, "mean_reward_first": log[0], "mean_reward_last": log[-1], "actor_mode": args.actor_mode}))


if __name__ == "__main__":
    main()
