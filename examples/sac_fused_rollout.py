#!/usr/bin/env python
"""SAC end to end on the device (BASELINE.json configs[2], the SAC half): squashed-Gaussian actor inference fused with the CSTR step
and the replay write (``FusedRollout``, actor kind "gaussian", tcgen05 hidden layer), Philox replay sampling, and the SAC gradient
step (``FusedSACUpdate`` = cstr_sac_update: entropy coefficient, soft target, critics, actor, polyak) — no torch autograd anywhere.

    python examples/sac_fused_rollout.py                      # 16,384 reactors, 30 episodes, 48,000 updates: learns in ~10 s
    python examples/sac_fused_rollout.py --n-envs 131072 --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096   # throughput shape
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/sac_fused_rollout.py --n-envs 1048576
        # reactors sharded over the GPUs; every rank updates on its own batch and the flat gradient ranges are all-reduced (NCCL) between
        # backward and Adam, so all ranks hold the same weights (--n-envs and the batch are then totals / per rank respectively)
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
    ap.add_argument("--n-envs", type=int, default=16384, help="total reactors over all ranks")
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--steps-per-iter", type=int, default=4)
    ap.add_argument("--updates-per-iter", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1024, help="per-rank batch")
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--actor-mode", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(args.seed)  # same initial weights on every rank
    env = pkg.dist.make_sharded_env(args.n_envs, rank, world, device=dev, seed=args.seed, monitor=False)
    n = env.num_envs
    buf = pkg.GpuReplayBuffer(args.rows * n, device=dev, n_envs=n, index_mode="philox", seed=args.seed * 1000 + rank)
    eng = pkg.FusedSACUpdate([256, 256], args.batch, device=dev, seed=args.seed * 7919, dp_rank=rank)

    def mlp(i, o):  # torch nn.Linear default init
        out = []
        for fi, fo in ((i, 256), (256, 256), (256, o)):
            b = 1.0 / np.sqrt(fi)
            out += [rng.uniform(-b, b, (fo, fi)).astype(np.float32), rng.uniform(-b, b, fo).astype(np.float32)]
        return out

    eng.load_nets({"actor": mlp(4, 4), "critic0": mlp(6, 1), "critic1": mlp(6, 1)})
    if world > 1:  # the gradient mean over the ranks happens inside the Adam kernels (NVLink peer memory): no collective launch
        eng.enable_peer_allreduce()
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
            mean_r = pkg.dist.global_sum(float(rsum.item()), device=dev) / (args.n_envs * args.steps_per_iter * ((it % window) + 1))
            rsum.zero_()
            critic_loss, actor_loss, _, ent_coef = eng.pop_losses()
            log.append(mean_r)
            if rank == 0:
                print(f"episode {len(log):3d} (iter {it:4d})  mean reward/step {mean_r:8.4f}  episode return {400 * mean_r:8.1f}  critic loss {critic_loss:.4f}  "
                      f"actor loss {actor_loss:.4f}  ent_coef {ent_coef:.4f}", flush=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    transitions = args.iters * args.steps_per_iter * args.n_envs
    if rank == 0:
        print(json.dumps({"world_size": world, "n_envs": args.n_envs, "transitions": transitions, "seconds": dt,
                          "transitions_per_s_incl_updates": transitions / dt, "updates": eng.n_updates, "mean_reward_first": log[0],
                          "mean_reward_last": log[-1], "actor_mode": args.actor_mode, "steps_per_iter": args.steps_per_iter,
                          "global_batch": args.batch * world, "peer_error": eng.peer_error()}))
    eng.close_peer_allreduce()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
