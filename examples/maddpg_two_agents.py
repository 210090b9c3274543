#!/usr/bin/env python
"""MADDPG with the two reactors as two agents on the GPU-resident step / replay path (BASELINE.json configs[4]).

    python examples/maddpg_two_agents.py                                  # 32,768 reactor pairs on one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/maddpg_two_agents.py --n-envs 262144                     # 262,144 env copies sharded over 8 GPUs

Agent 0 sees (C1, T1) and sets the coolant flow of reactor 1, agent 1 sees (C2, T2) and sets that of reactor 2 — the reference's
``observation_splits=[[0, 1], [2, 3]]``, ``action_splits=[[0], [1]]`` (experiments with ``core/maddpg``).  Per iteration, on every rank:

1. ``--steps-per-iter`` control intervals: both per-agent actors (2 -> 400 -> 300 -> 1, torch) on their observation slices, Gaussian
   exploration noise, then ``GpuCSTRVecEnv.step_tensor`` (the CSTR step kernel: dynamics, reward, truncation, device-side auto-reset)
   and ``GpuReplayBuffer.add`` (the coalesced 64-byte-record write) on device tensors — no host copy, no per-env Python;
2. ``--updates-per-iter`` gradient steps: ``GpuReplayBuffer.sample`` (Philox gather kernel) straight into the float32 tensors of the
   update, which follows the loop body of ``core/maddpg/maddpg.py:117-191`` in torch: per agent twin centralised critics Q_i(s, a_1, a_2)
   with a clipped-noise target from the target actors and the twin minimum, delayed actor step through Q_i's first critic, polyak.
   Multi-GPU: gradients are all-reduced (one flat NCCL bucket per optimiser) so every rank holds the same agents.

The multi-agent update itself stays torch (the fused update engine is specialised to the single-agent 4 -> 2 shapes; its oracle,
``oracle/td3_oracle.py::MultiAgentDDPGOracle``, is pinned against the reference in ``tests/golden/maddpg_update.npz``); the hot
path exercised here is the env step, the buffer write and the sample.  What to expect (``profiles/r01_maddpg_two_agents.log``): 5.1e6
transitions/s on one B200 and 9.0e6 on two, bounded by the torch updates; the return does NOT improve — neither here nor in the
reference's own ``MADDPG.learn`` on this task (``profiles/r01_reference_algos.log``: eval return -678); ``--own-observations`` reaches
-140 after two episodes and then diverges as well.  This example is about the data path, not about tuning MADDPG.
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
import torch.nn.functional as F

OBS_SPLITS, ACT_SPLITS = [[0, 1], [2, 3]], [[0], [1]]


def mlp(i, o, squash):
    layers = [nn.Linear(i, 400), nn.ReLU(), nn.Linear(400, 300), nn.ReLU(), nn.Linear(300, o)]
    return nn.Sequential(*(layers + ([nn.Tanh()] if squash else [])))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=32768, help="total env copies over all ranks")
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--steps-per-iter", type=int, default=4)
    ap.add_argument("--updates-per-iter", type=int, default=4)
    ap.add_argument("--batch", type=int, default=1024, help="per-rank batch")
    ap.add_argument("--rows", type=int, default=64, help="ring rows per rank")
    ap.add_argument("--sigma", type=float, default=0.2)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--own-observations", action="store_true",
                    help="actor step: feed every actor ITS OWN observation slice (textbook MADDPG) instead of the reference's maddpg.py:169-171, "
                         "which feeds the slice of the agent being updated to all actors")
    args = ap.parse_args()

    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)
    n_agents = len(OBS_SPLITS)
    env = pkg.dist.make_sharded_env(args.n_envs, rank, world, device=dev, seed=args.seed, monitor=False)
    n = env.num_envs
    buf = pkg.GpuReplayBuffer(args.rows * n, device=dev, n_envs=n, index_mode="philox", seed=args.seed * 1000 + rank)
    actors = nn.ModuleList([mlp(len(o), len(a), True) for o, a in zip(OBS_SPLITS, ACT_SPLITS)]).to(dev)
    actors_t = nn.ModuleList([mlp(len(o), len(a), True) for o, a in zip(OBS_SPLITS, ACT_SPLITS)]).to(dev)
    critics = nn.ModuleList([nn.ModuleList([mlp(6, 1, False) for _ in range(2)]) for _ in range(n_agents)]).to(dev)  # Q_i(s, a): all obs, all actions
    critics_t = nn.ModuleList([nn.ModuleList([mlp(6, 1, False) for _ in range(2)]) for _ in range(n_agents)]).to(dev)
    pkg.dist.broadcast_parameters(list(actors.parameters()) + list(critics.parameters()))
    actors_t.load_state_dict(actors.state_dict())
    critics_t.load_state_dict(critics.state_dict())
    opt_a = [torch.optim.Adam(actors[i].parameters(), lr=args.lr) for i in range(n_agents)]
    opt_c = [torch.optim.Adam(critics[i].parameters(), lr=args.lr) for i in range(n_agents)]
    buckets = {}
    gamma, tau, tnoise, tclip, delay = 0.99, 0.005, 0.2, 0.5, 2

    def reduce_grads(key, params):
        if world > 1:
            buckets[key] = pkg.dist.allreduce_gradients(list(params), bucket=buckets.get(key))

    def polyak(src, dst):
        with torch.no_grad():
            for p, t in zip(src.parameters(), dst.parameters()):
                t.mul_(1 - tau).add_(p, alpha=tau)

    @torch.no_grad()
    def act(obs, explore):
        a = torch.cat([actors[i](obs[:, OBS_SPLITS[i]]) for i in range(n_agents)], 1)
        if explore:
            a = (a + args.sigma * torch.randn_like(a)).clamp(-1, 1)
        return a

    env.reset()
    obs = env.state
    rsum = torch.zeros((), dtype=torch.float64, device=dev)
    log, n_updates = [], 0
    window = max(1, 400 // args.steps_per_iter)  # iterations per 400-interval episode: report whole episodes
    torch.cuda.synchronize()
    t0 = time.time()
    for it in range(args.iters):
        for _ in range(args.steps_per_iter):
            prev = obs.clone()  # the step kernel advances env.state in place
            a = torch.rand(n, 2, device=dev) * 2 - 1 if not buf.full and buf.pos < 8 else act(prev, True)
            obs, rew, done, terminal = env.step_tensor(a)
            nxt = torch.where(done.bool()[:, None], terminal, obs)  # the transition ends in the terminal observation, not the reset one
            buf.add(prev, nxt, a, rew, done, timeouts=done)  # every done of this env is a time-limit truncation
            rsum += rew.sum(dtype=torch.float64)
        for _ in range(args.updates_per_iter):
            n_updates += 1
            b = buf.sample(args.batch)
            with torch.no_grad():  # maddpg.py:132-144 target actions of every agent, clipped smoothing noise
                nxt_a = torch.cat([(actors_t[i](b.next_observations[:, OBS_SPLITS[i]]) +
                                    (torch.randn(args.batch, len(ACT_SPLITS[i]), device=dev) * tnoise).clamp(-tclip, tclip)).clamp(-1, 1)
                                   for i in range(n_agents)], 1)
                nxt_in = torch.cat([b.next_observations, nxt_a], 1)
            cur_in = torch.cat([b.observations, b.actions], 1)
            for i in range(n_agents):
                with torch.no_grad():  # :147-152
                    target = b.rewards + (1 - b.dones) * gamma * torch.min(*[q(nxt_in) for q in critics_t[i]])
                loss_c = sum(F.mse_loss(q(cur_in), target) for q in critics[i])  # :155-158
                opt_c[i].zero_grad(set_to_none=True)
                loss_c.backward()
                reduce_grads(("c", i), critics[i].parameters())
                opt_c[i].step()
                if n_updates % delay == 0:  # :166-185 (the reference feeds agent i's observation slice to every actor here; kept by default)
                    joint = torch.cat([actors[j](b.observations[:, OBS_SPLITS[j if args.own_observations else i]]) for j in range(n_agents)], 1)
                    loss_a = -critics[i][0](torch.cat([b.observations, joint], 1)).mean()
                    opt_a[i].zero_grad(set_to_none=True)
                    loss_a.backward()
                    reduce_grads(("a", i), actors[i].parameters())
                    opt_a[i].step()
                    for j in range(n_agents):  # the other actors only lent their forward pass
                        if j != i:
                            actors[j].zero_grad(set_to_none=True)
                    polyak(critics, critics_t)
                    polyak(actors, actors_t)
        if (it + 1) % window == 0 or it == args.iters - 1:
            iters_in = (it % window) + 1
            mean_r = pkg.dist.global_sum(float(rsum.item()), device=dev) / (args.n_envs * args.steps_per_iter * iters_in)
            rsum.zero_()
            log.append(mean_r)
            if rank == 0:
                print(f"episode {len(log):3d} (iter {it:4d})  mean reward/step {mean_r:8.4f}  episode return {400 * mean_r:8.1f}  "
                      f"critic loss {float(loss_c.detach()):.4f}", flush=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        transitions = args.iters * args.steps_per_iter * args.n_envs
        print(json.dumps({"world_size": world, "n_envs": args.n_envs, "agents": n_agents, "transitions": transitions, "seconds": dt,
                          "transitions_per_s_incl_updates": transitions / dt, "updates": n_updates, "env_launches": env.launches,
                          "buffer_launches": buf.launches, "mean_reward_first": log[0], "mean_reward_last": log[-1],
                          "actor_step": "own observations" if args.own_observations else "reference (maddpg.py:169-171)"}))
    if world > 1:
        torch.distributed.destroy_process_group()
    return log


if __name__ == "__main__":
    main()
