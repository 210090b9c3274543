#!/usr/bin/env python
"""MADDPG / IDDPG with the two reactors as two agents, entirely on the device (BASELINE.json configs[4]).

    python examples/maddpg_two_agents.py                                  # 32,768 reactor pairs on one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/maddpg_two_agents.py --n-envs 262144                     # 262,144 env copies sharded over 8 GPUs

Agent 0 sees (C1, T1) and sets the coolant flow of reactor 1, agent 1 sees (C2, T2) and sets that of reactor 2 — the reference's
``observation_splits=[[0, 1], [2, 3]]``, ``action_splits=[[0], [1]]``.  Per iteration, on every rank:

1. ``--steps-per-iter`` control intervals in ONE launch of ``cstr_rollout_fused_multi``: both per-agent actors (2 -> 400 -> 300 -> 1) on
   their observation slices, the CSTR step, reward / truncation / device-side auto-reset and the 64-byte replay record — with the
   reference's multi-agent ``_sample_action`` semantics (no exploration noise or rescaling is ever applied there: quirk Q5);
2. ``--updates-per-iter`` gradient steps of ``cstr_ma_update`` (``FusedMultiAgentUpdate``): Philox sample + the loop body of
   ``core/maddpg/maddpg.py:127-185`` (``--iddpg``: ``core/iddpg/iddpg.py``) replayed from one CUDA graph per ``policy_delay`` cycle.
   Multi-GPU: the gradient mean over the ranks is taken inside the Adam kernels over NVLink peer memory (``enable_peer_allreduce``), so
   every rank holds the same agents and no collective is launched.

Nothing leaves the device between iterations except the episode statistics.  As in the reference's own ``MADDPG.learn`` on this task
(``profiles/r01_reference_algos.log``: eval return -678) the return does not improve much: without exploration noise (Q5) and with agent i's
observation fed to every actor (maddpg.py:169-171) there is little to learn from; this example is about the data path.
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

import torch
import torch.nn as nn


def mlp(i, o, squash):
    layers = [nn.Linear(i, 400), nn.ReLU(), nn.Linear(400, 300), nn.ReLU(), nn.Linear(300, o)]
    return nn.Sequential(*(layers + ([nn.Tanh()] if squash else [])))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=32768, help="total env copies over all ranks")
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--steps-per-iter", type=int, default=4)
    ap.add_argument("--updates-per-iter", type=int, default=4)
    ap.add_argument("--batch", type=int, default=1024, help="per-rank batch")
    ap.add_argument("--rows", type=int, default=64, help="ring rows per rank")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--iddpg", action="store_true", help="independent critics over each agent's own slices (core/iddpg) instead of centralised ones")
    args = ap.parse_args(argv)

    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)  # the same initial agents on every rank
    env = pkg.dist.make_sharded_env(args.n_envs, rank, world, device=dev, seed=args.seed, monitor=False)
    n = env.num_envs
    buf = pkg.GpuReplayBuffer(args.rows * n, device=dev, n_envs=n, index_mode="philox", seed=args.seed * 1000 + rank)
    actors = [mlp(2, 1, True).to(dev) for _ in range(2)]
    ci = 3 if args.iddpg else 6
    nets = {f"actor{i}": [p.detach() for p in actors[i].parameters()] for i in range(2)}
    nets.update({f"critic{i}_{k}": [p.detach() for p in mlp(ci, 1, False).to(dev).parameters()] for i in range(2) for k in range(2)})
    eng = pkg.FusedMultiAgentUpdate([400, 300], args.batch, not args.iddpg, device=dev, actor_lrs=[args.lr] * 2, critic_lrs=[args.lr] * 2,
                                    seed=args.seed * 7919, dp_rank=rank)
    eng.load_nets(nets)
    if world > 1:
        torch.distributed.broadcast(eng.params, 0)
        eng.targets.copy_(eng.params)
        eng.enable_peer_allreduce()
    views = eng.views("params")
    for i in range(2):  # the rollout's actor modules share the engine's parameter block
        for p, v in zip(actors[i].parameters(), views[f"actor{i}"]):
            p.data = v
    weights = pkg.AgentActorWeights(actors, device=dev)
    roll = pkg.FusedRollout(env, buf, weights)
    stats = pkg.EpisodeStats(n, device=dev)
    env.reset()
    log = []
    torch.cuda.synchronize()
    t0 = time.time()
    for it in range(args.iters):
        warm = not buf.full and buf.pos < 8
        if not warm:
            weights.refresh_from_modules(actors)
        roll.collect(args.steps_per_iter, warmup=warm, stats=stats)
        if buf.full or buf.pos >= 8:
            eng.train(args.updates_per_iter, buf, args.batch, graph=True)
        if (it + 1) % max(1, 400 // args.steps_per_iter) == 0 or it == args.iters - 1:
            done = stats.pop()
            tot, cnt = pkg.dist.global_sum(float(done[:, 0].sum()), device=dev), pkg.dist.global_sum(float(len(done)), device=dev)
            losses = eng.pop_losses()
            if cnt:
                log.append(tot / cnt)
                if rank == 0:
                    print(f"iter {it:4d}  episodes {int(cnt):8d}  mean episode return {tot / cnt:8.1f}  critic losses {losses[0][0]:.4f} {losses[1][0]:.4f}",
                          flush=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    out = None
    if rank == 0:
        transitions = args.iters * args.steps_per_iter * args.n_envs
        out = {"world_size": world, "n_envs": args.n_envs, "algo": "IDDPG" if args.iddpg else "MADDPG", "transitions": transitions, "seconds": dt,
               "transitions_per_s_incl_updates": transitions / dt, "updates": eng.n_updates, "global_batch": args.batch * world,
               "rollout_launches": roll.launches, "env_step_launches": env.launches, "buffer_launches": buf.launches, "peer_error": eng.peer_error(),
               "episode_return_first": log[0] if log else None, "episode_return_last": log[-1] if log else None}
        print(json.dumps(out))
    eng.close_peer_allreduce()
    if world > 1:
        torch.distributed.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
