#!/usr/bin/env python
"""TD3 end to end on the GPU-resident rollout path (config #3 shape, any number of GPUs).

    python examples/td3_fused_rollout.py        # 16,384 reactors, 30 episodes, 48,000 updates: episode return -281 -> -33 in ~11 s
    python examples/td3_fused_rollout.py --n-envs 131072 --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096   # throughput shape
    torchrun --nproc-per-node 8 examples/td3_fused_rollout.py --n-envs 1048576 --iters 200 --steps-per-iter 8 --updates-per-iter 8 --batch 4096

Per iteration, on every rank:
  1. ``FusedRollout.collect(K)``   — actor inference (tcgen05) + noise + bounds + CSTR step + reward/done + replay
                                     records for this rank's reactor shard, ONE kernel launch, nothing leaves the GPU;
  2. ``GpuReplayBuffer.sample(B)`` — Philox-index gather straight into the update's float32 tensors;
  3. TD3 update with the reference's semantics (``core/td3/td3.py:154-211``: target policy smoothing, twin-min
     target, delayed actor, polyak): ``--update fused`` (default) = ``FusedTD3Update`` (cstr_td3_update: hand-written
     forward/backward/Adam/polyak kernels, the flat gradient block all-reduced in place with ONE NCCL call per phase);
     ``--update torch`` = the same update in eager torch (autograd + cuBLAS) for comparison;
  4. the fresh actor weights are handed back to the rollout kernel device-to-device (the nn.Module parameters are
     views of the fused engine's flat parameter block, so no copy is involved on the update side).
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


def mlp(i, hs, o, squash=False):
    layers, last = [], i
    for h in hs:
        layers += [nn.Linear(last, h), nn.ReLU()]
        last = h
    layers.append(nn.Linear(last, o))
    if squash:
        layers.append(nn.Tanh())
    return nn.Sequential(*layers)


def polyak(src, dst, tau):
    with torch.no_grad():
        ps, pt = list(src.parameters()), list(dst.parameters())
        torch._foreach_mul_(pt, 1.0 - tau)
        torch._foreach_add_(pt, ps, alpha=tau)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=16384, help="total reactors over all ranks")
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--steps-per-iter", type=int, default=4, help="env steps per fused launch")
    ap.add_argument("--updates-per-iter", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1024, help="per-rank batch")
    ap.add_argument("--rows", type=int, default=64, help="ring rows per rank")
    ap.add_argument("--actor-mode", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--lr", type=float, default=1e-3, help="Adam learning rate (the reference's TD3 default)")
    ap.add_argument("--sigma", type=float, default=0.1, help="exploration noise of the rollout (NormalActionNoise)")
    ap.add_argument("--update", default="fused", choices=["fused", "torch"])
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tensor", "bf16"], help="hidden-layer GEMMs of the fused update")
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl-eager"],
                    help="multi-GPU gradient mean: peer = inside the Adam kernels over NVLink peer memory (the whole update in a CUDA graph), "
                         "nccl-eager = launch by launch with ncclAllReduce between the phases (round 1)")
    args = ap.parse_args()

    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)  # same initial weights on every rank (+ broadcast below)

    env = pkg.dist.make_sharded_env(args.n_envs, rank, world, device=dev, seed=args.seed, monitor=False)
    n = env.num_envs
    buf = pkg.GpuReplayBuffer(args.rows * n, device=dev, n_envs=n, index_mode="philox", seed=args.seed * 1000 + rank)
    actor, actor_t = mlp(4, [400, 300], 2, squash=True).to(dev), mlp(4, [400, 300], 2, squash=True).to(dev)
    critics = nn.ModuleList([mlp(6, [400, 300], 1) for _ in range(2)]).to(dev)
    critics_t = nn.ModuleList([mlp(6, [400, 300], 1) for _ in range(2)]).to(dev)
    pkg.dist.broadcast_parameters(list(actor.parameters()) + list(critics.parameters()))
    actor_t.load_state_dict(actor.state_dict())
    critics_t.load_state_dict(critics.state_dict())
    opt_a = torch.optim.Adam(actor.parameters(), lr=args.lr)
    opt_c = torch.optim.Adam(critics.parameters(), lr=args.lr)
    weights = pkg.ActorWeights.from_module(actor, device=dev)
    roll = pkg.FusedRollout(env, buf, weights, sigma=args.sigma, actor_mode=args.actor_mode)
    gamma, tau, tnoise, tclip, delay = 0.99, 0.005, 0.2, 0.5, 2
    fused = None
    if args.update == "fused":
        fused = pkg.FusedTD3Update([400, 300], args.batch, device=dev, gamma=gamma, tau=tau, learning_rate=args.lr, policy_delay=delay,
                                   target_policy_noise=tnoise, target_noise_clip=tclip, seed=args.seed * 7919, dp_rank=rank, gemm=args.gemm)
        fused.adopt_modules(actor, list(critics), actor_t, list(critics_t))  # module parameters become views of the flat blocks
        hook = pkg.dist.allreduce_flat if world > 1 and args.dp != "peer" else None
        if world > 1 and args.dp == "peer":
            fused.enable_peer_allreduce()

    env.reset()
    roll.collect(args.steps_per_iter, warmup=True)  # learning_starts phase: uniform random actions
    rsum = torch.zeros(1, dtype=torch.float64, device=dev)
    bucket_c = bucket_a = None
    n_updates, log = 0, []
    torch.cuda.synchronize()
    t0 = time.time()
    window = max(1, 400 // args.steps_per_iter)  # iterations per 400-step episode: rewards depend on the episode phase, so report whole episodes
    for it in range(args.iters):
        roll.collect(args.steps_per_iter, reward_sum=rsum)
        if fused is not None:  # sample + update on the device: whole policy_delay cycles (incl. the gradient mean over ranks) replay from one CUDA graph
            fused.train(args.updates_per_iter, buf, args.batch, allreduce=hook, graph=(args.dp != "nccl-eager"))
            n_updates += args.updates_per_iter
        for _ in range(args.updates_per_iter if fused is None else 0):
            b = buf.sample(args.batch)
            with torch.no_grad():
                noise = (torch.randn_like(b.actions) * tnoise).clamp(-tclip, tclip)
                na = (actor_t(b.next_observations) + noise).clamp(-1, 1)
                q_next = torch.min(*[c(torch.cat([b.next_observations, na], 1)) for c in critics_t])
                target = b.rewards + (1 - b.dones) * gamma * q_next
            qs = [c(torch.cat([b.observations, b.actions], 1)) for c in critics]
            loss_c = sum(F.mse_loss(q, target) for q in qs)
            opt_c.zero_grad(set_to_none=False)
            loss_c.backward()
            bucket_c = pkg.dist.allreduce_gradients(list(critics.parameters()), bucket=bucket_c)
            opt_c.step()
            n_updates += 1
            if n_updates % delay == 0:
                loss_a = -critics[0](torch.cat([b.observations, actor(b.observations)], 1)).mean()
                opt_a.zero_grad(set_to_none=False)
                loss_a.backward()
                bucket_a = pkg.dist.allreduce_gradients(list(actor.parameters()), bucket=bucket_a)
                opt_a.step()
                polyak(actor, actor_t, tau)
                polyak(critics, critics_t, tau)
        weights.refresh_from_module(actor)  # device-to-device; repacks the bf16 UMMA image of W2
        if (it + 1) % window == 0 or it == args.iters - 1:
            iters_in = (it % window) + 1
            mean_r = pkg.dist.global_sum(float(rsum.item()), device=dev) / (args.n_envs * args.steps_per_iter * iters_in)
            rsum.zero_()
            log.append(mean_r)
            if rank == 0:
                closs = fused.pop_losses()[0] if fused is not None else float(loss_c.detach())
                print(f"episode {len(log):3d} (iter {it:4d})  mean reward/step {mean_r:8.4f}  episode return {400 * mean_r:8.1f}  critic loss {closs:.4f}", flush=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        transitions = args.iters * args.steps_per_iter * args.n_envs
        print(json.dumps({"world_size": world, "n_envs": args.n_envs, "transitions": transitions, "seconds": dt,
                          "transitions_per_s_incl_updates": transitions / dt, "updates": n_updates,
                          "mean_reward_first": log[0], "mean_reward_last": log[-1], "actor_mode": args.actor_mode, "update": args.update,
                          "gemm": args.gemm, "dp": args.dp if world > 1 else None, "global_batch": args.batch * world,
                          "peer_error": fused.peer_error() if fused is not None else 0}))
    if fused is not None:
        fused.close_peer_allreduce()
    if world > 1:
        torch.distributed.destroy_process_group()
    return log


if __name__ == "__main__":
    main()
