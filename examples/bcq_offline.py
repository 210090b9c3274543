#!/usr/bin/env python
"""BCQ offline training on a synthetic 10M-transition CSTR dataset held in the GPU-resident replay buffer
(BASELINE.json configs[3]).

    python examples/bcq_offline.py [--transitions 10240000] [--updates 300] [--batch 4096]

1. The dataset is produced ON the device: ``FusedRollout.collect(400, warmup=True)`` drives 25,600 reactors for one full episode
   with uniform random actions (the behaviour policy of an offline dataset) and writes the 64-byte records straight into the
   ring — no pickle, no host copy.  (``GpuReplayBuffer.from_reference`` is the path for a dataset that already exists as a
   reference pickle; ``oracle/run_reference_algos.py`` runs the reference's own ``BCQ.learn`` on it.)
2. Every gradient step samples with the Philox gather kernel (``GpuReplayBuffer.sample``) straight into the float32 tensors the
   update consumes.
3. The update follows ``core/bcq/bcq.py:129-213`` (VAE reconstruction + 0.5 KL, 10 candidate actions from the target VAE +
   perturbation net, twin-critic min then the max over the reference's (B, 10) reshape, delayed perturbation-actor step, polyak) in
   plain torch — the BCQ update kernels are not built (their oracle is: ``oracle/td3_oracle.py::BCQUpdateOracle``,
   pinned against the reference in ``tests/golden/bcq_update.npz``); the hot path here is 1-2.  Network sizes default to
   ``BCQPolicy``'s (core/bcq/policies.py:305-307).
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


def mlp(i, hs, o):
    layers, last = [], i
    for h in hs:
        layers += [nn.Linear(last, h), nn.ReLU()]
        last = h
    return nn.Sequential(*layers, nn.Linear(last, o))


class VAE(nn.Module):  # core/bcq/policies.py: encoder (s,a) -> (mean, log_std), decoder (s,z) -> a
    def __init__(self, latent=32, hidden=64):
        super().__init__()
        self.enc, self.dec, self.latent = mlp(6, [hidden, hidden], 2 * latent), mlp(4 + latent, [hidden, hidden], 2), latent

    def forward(self, s, a):
        mean, log_std = self.enc(torch.cat([s, a], 1)).chunk(2, 1)
        std = log_std.clamp(-4, 15).exp()
        return self.decode(s, mean + std * torch.randn_like(std)), mean, std

    def decode(self, s, z=None):
        if z is None:
            z = torch.randn(s.shape[0], self.latent, device=s.device).clamp(-0.5, 0.5)
        return torch.tanh(self.dec(torch.cat([s, z], 1)))


class Actor(nn.Module):  # VAE proposal + perturbation net: a + phi * tanh(xi(s, a)), clipped to the action box
    def __init__(self, phi=0.05, latent=32, vae_hidden=64, pert_hidden=64):
        super().__init__()
        self.vae, self.xi, self.phi = VAE(latent, vae_hidden), mlp(6, [pert_hidden, pert_hidden], 2), phi

    def forward(self, s, num_samples=1):
        s = s.repeat(num_samples, 1)
        a = self.vae.decode(s)
        return (a + self.phi * torch.tanh(self.xi(torch.cat([s, a], 1)))).clamp(-1, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transitions", type=int, default=10_240_000)
    ap.add_argument("--updates", type=int, default=300)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--latent", type=int, default=32)
    ap.add_argument("--vae-hidden", type=int, default=64)
    ap.add_argument("--pert-hidden", type=int, default=64)
    args = ap.parse_args()
    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    dev = torch.device("cuda", 0)
    torch.manual_seed(args.seed)
    T = 400
    n = args.transitions // T
    env = pkg.GpuCSTRVecEnv(n, device=dev, seed=args.seed, monitor=False)
    buf = pkg.GpuReplayBuffer(T * n, device=dev, n_envs=n, index_mode="philox", seed=args.seed)
    behaviour = pkg.ActorWeights(*[torch.zeros(s) for s in ((400, 4), (400,), (300, 400), (300,), (2, 300), (2,))], device=dev)
    roll = pkg.FusedRollout(env, buf, behaviour, sigma=0.0, actor_mode="fp32")
    env.reset()
    torch.cuda.synchronize()
    t0 = time.time()
    roll.collect(T, warmup=True)  # uniform random actions: one whole episode of every reactor, straight into the ring
    torch.cuda.synchronize()
    t_data = time.time() - t0
    assert buf.full and buf.size() * n == T * n
    mk = lambda: Actor(0.05, args.latent, args.vae_hidden, args.pert_hidden).to(dev)  # noqa: E731
    actor, actor_t = mk(), mk()
    critics = nn.ModuleList([mlp(6, [400, 300], 1) for _ in range(2)]).to(dev)
    critics_t = nn.ModuleList([mlp(6, [400, 300], 1) for _ in range(2)]).to(dev)
    actor_t.load_state_dict(actor.state_dict())
    critics_t.load_state_dict(critics.state_dict())
    opt_vae, opt_xi = torch.optim.Adam(actor.vae.parameters(), lr=1e-3), torch.optim.Adam(actor.xi.parameters(), lr=1e-3)
    opt_c = torch.optim.Adam(critics.parameters(), lr=1e-3)
    gamma, tau, delay, ncand, B = 0.99, 0.005, 2, 10, args.batch
    log = []
    torch.cuda.synchronize()
    t0 = time.time()
    for it in range(1, args.updates + 1):
        b = buf.sample(B)
        recon, mean, std = actor.vae(b.observations, b.actions)
        vae_loss = F.mse_loss(recon, b.actions) + 0.5 * (-0.5 * (1 + torch.log(std.pow(2)) - mean.pow(2) - std.pow(2)).mean())
        opt_vae.zero_grad(set_to_none=True)
        vae_loss.backward()
        opt_vae.step()
        with torch.no_grad():
            actor_t.vae.load_state_dict(actor.vae.state_dict())
            cand = actor_t(b.next_observations, ncand)
            nobs = b.next_observations.repeat(ncand, 1)
            # bcq.py:170-171 as written: the candidate-major (10 B, 1) column reshaped row-major to (B, 10) before the max
            q = torch.min(*[c(torch.cat([nobs, cand], 1)) for c in critics_t]).reshape(B, ncand).max(1)[0].unsqueeze(1)
            target = b.rewards + (1 - b.dones) * gamma * q
        critic_loss = sum(F.mse_loss(c(torch.cat([b.observations, b.actions], 1)), target) for c in critics)
        opt_c.zero_grad(set_to_none=True)
        critic_loss.backward()
        opt_c.step()
        if it % delay == 0:
            actor_loss = -critics[0](torch.cat([b.observations, actor(b.observations)], 1)).mean()
            opt_xi.zero_grad(set_to_none=True)
            actor_loss.backward()
            opt_xi.step()
            with torch.no_grad():
                for src, dst in ((critics, critics_t), (actor, actor_t)):
                    for p, t in zip(src.parameters(), dst.parameters()):
                        t.mul_(1 - tau).add_(p, alpha=tau)
        if it % 50 == 0 or it == 1:
            log.append((it, float(vae_loss.detach()), float(critic_loss.detach())))
            print(f"update {it:4d}  vae loss {log[-1][1]:.4f}  critic loss {log[-1][2]:.4f}", flush=True)
    torch.cuda.synchronize()
    t_train = time.time() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        buf.sample(1 << 20)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"transitions": T * n, "dataset_bytes": T * n * 64, "dataset_seconds": t_data, "dataset_transitions_per_s": T * n / t_data,
                      "updates": args.updates, "batch": B, "updates_per_s": args.updates / t_train,
                      "sample_1M_rows_ms": e0.elapsed_time(e1) / 50, "vae_loss_first_last": [log[0][1], log[-1][1]],
                      "critic_loss_first_last": [log[0][2], log[-1][2]]}))


if __name__ == "__main__":
    main()
