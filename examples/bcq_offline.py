#!/usr/bin/env python
"""BCQ offline training on a synthetic 10M-transition CSTR dataset held in the GPU-resident replay buffer
(BASELINE.json configs[3]).

    python examples/bcq_offline.py [--transitions 10240000] [--updates 300] [--batch 256] [--script-sizes]

1. The dataset is produced ON the device: ``FusedRollout.collect(400, warmup=True)`` drives 25,600 reactors for one full episode
   with uniform random actions (the behaviour policy of an offline dataset) and writes the 64-byte records straight into the
   ring — no pickle, no host copy.  (``GpuReplayBuffer.from_reference`` is the path for a dataset that already exists as a
   reference pickle; ``tests/test_gpu_reference_learn.py`` runs the reference's own ``BCQ.learn`` on one.)
2. Every gradient step is ``cstr_bcq_update`` (``FusedBCQUpdate``): Philox sample straight into the update's tensors, then the loop
   body of ``core/bcq/bcq.py:137-205`` — VAE reconstruction + 0.5 KL, ten candidates through the refreshed VAE and the target
   perturbation net, twin-critic min then the max over the reference's (B, 10) reshape, delayed perturbation step, polyak — replayed
   from one CUDA graph per ``actor_delay`` cycle.  ``--torch`` times the same update in eager torch on the same GPU beside it.
   Network sizes default to ``BCQPolicy``'s (core/bcq/policies.py:320-322); ``--script-sizes`` uses the experiment script's
   (``HalfCheetah_BCQ.py:55-58``: vae 700 / latent 12 / perturbation 400).
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


class TorchBCQ:
    """The same update in eager torch (autograd, torch.optim.Adam): the comparison arm, what the unmodified reference executes on a GPU."""

    def __init__(self, dev, latent, hv, hp):
        self.enc, self.dec, self.xi = mlp(6, [hv, hv], 2 * latent).to(dev), mlp(4 + latent, [hv, hv], 2).to(dev), mlp(6, [hp, hp], 2).to(dev)
        self.xi_t = mlp(6, [hp, hp], 2).to(dev)
        self.critics = nn.ModuleList([mlp(6, [400, 300], 1) for _ in range(2)]).to(dev)
        self.critics_t = nn.ModuleList([mlp(6, [400, 300], 1) for _ in range(2)]).to(dev)
        self.xi_t.load_state_dict(self.xi.state_dict())
        self.critics_t.load_state_dict(self.critics.state_dict())
        self.opt_vae = torch.optim.Adam(list(self.enc.parameters()) + list(self.dec.parameters()), lr=1e-3)
        self.opt_xi, self.opt_c = torch.optim.Adam(self.xi.parameters(), lr=1e-3), torch.optim.Adam(self.critics.parameters(), lr=1e-3)
        self.latent, self.n = latent, 0

    def decode(self, s):
        z = torch.randn(s.shape[0], self.latent, device=s.device).clamp(-0.5, 0.5)
        return torch.tanh(self.dec(torch.cat([s, z], 1)))

    def perturb(self, xi, s, a):
        return (a + 0.05 * torch.tanh(xi(torch.cat([s, a], 1)))).clamp(-1, 1)

    def update(self, b):
        self.n += 1
        B = b.observations.shape[0]
        mean, log_std = self.enc(torch.cat([b.observations, b.actions], 1)).chunk(2, 1)
        std = log_std.clamp(-4, 15).exp()
        recon = torch.tanh(self.dec(torch.cat([b.observations, mean + std * torch.randn_like(std)], 1)))
        vae_loss = F.mse_loss(recon, b.actions) + 0.5 * (-0.5 * (1 + torch.log(std.pow(2)) - mean.pow(2) - std.pow(2)).mean())
        self.opt_vae.zero_grad(set_to_none=True)
        vae_loss.backward()
        self.opt_vae.step()
        with torch.no_grad():
            nobs = b.next_observations.repeat(10, 1)
            cand = self.perturb(self.xi_t, nobs, self.decode(nobs))
            q = torch.min(*[c(torch.cat([nobs, cand], 1)) for c in self.critics_t]).reshape(B, 10).max(1)[0].unsqueeze(1)  # bcq.py:170-171 as written
            target = b.rewards + (1 - b.dones) * 0.99 * q
        critic_loss = sum(F.mse_loss(c(torch.cat([b.observations, b.actions], 1)), target) for c in self.critics)
        self.opt_c.zero_grad(set_to_none=True)
        critic_loss.backward()
        self.opt_c.step()
        if self.n % 2 == 0:
            a = self.perturb(self.xi, b.observations, self.decode(b.observations))
            actor_loss = -self.critics[0](torch.cat([b.observations, a], 1)).mean()
            self.opt_xi.zero_grad(set_to_none=True)
            actor_loss.backward()
            self.opt_xi.step()
            with torch.no_grad():
                for src, dst in ((self.critics, self.critics_t), (self.xi, self.xi_t)):
                    for p, t in zip(src.parameters(), dst.parameters()):
                        t.mul_(1 - 0.005).add_(p, alpha=0.005)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--transitions", type=int, default=10_240_000)
    ap.add_argument("--updates", type=int, default=300)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--script-sizes", action="store_true", help="vae 700 / latent 12 / perturbation 400 (HalfCheetah_BCQ.py:55-58)")
    ap.add_argument("--torch", action="store_true", help="also time the same update in eager torch on this GPU")
    args = ap.parse_args(argv)
    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    dev = torch.device("cuda", 0)
    torch.manual_seed(args.seed)
    latent, hv, hp = (12, 700, 400) if args.script_sizes else (32, 64, 64)
    T = 400
    n = args.transitions // T
    env = pkg.GpuCSTRVecEnv(n, device=dev, seed=args.seed, monitor=False)
    buf = pkg.GpuReplayBuffer(T * n, device=dev, n_envs=n, index_mode="philox", seed=args.seed)
    roll = pkg.FusedRollout(env, buf, None, sigma=0.0, actor_mode="fp32")
    env.reset()
    torch.cuda.synchronize()
    t0 = time.time()
    roll.collect(T, warmup=True)  # uniform random actions: one whole episode of every reactor, straight into the ring
    torch.cuda.synchronize()
    t_data = time.time() - t0
    assert buf.full and buf.size() * n == T * n
    ref = TorchBCQ(dev, latent, hv, hp)
    eng = pkg.FusedBCQUpdate(latent, hv, hp, [400, 300], args.batch, device=dev, seed=args.seed)
    P = lambda m: [p.detach() for p in m.parameters()]  # noqa: E731
    eng.load_nets({"vae_enc": P(ref.enc), "vae_dec": P(ref.dec), "pert": P(ref.xi), "critic0": P(ref.critics[0]), "critic1": P(ref.critics[1])})
    log = []
    torch.cuda.synchronize()
    t0 = time.time()
    done = 0
    while done < args.updates:
        k = min(50, args.updates - done)
        eng.train(k, buf, args.batch, graph=True)
        done += k
        vae_loss, critic_loss, actor_loss = eng.pop_losses()
        log.append((done, vae_loss, critic_loss))
        print(f"update {done:5d}  vae loss {vae_loss:.4f}  critic loss {critic_loss:.4f}  actor loss {actor_loss if actor_loss is None else round(actor_loss, 4)}", flush=True)
    torch.cuda.synchronize()
    t_train = time.time() - t0
    out = {"transitions": T * n, "dataset_bytes": T * n * 64, "dataset_seconds": t_data, "dataset_transitions_per_s": T * n / t_data,
           "updates": args.updates, "batch": args.batch, "sizes": {"latent": latent, "vae_hidden": hv, "pert_hidden": hp, "critic": [400, 300]},
           "fused_updates_per_s": args.updates / t_train, "fused_ms_per_update": 1e3 * t_train / args.updates,
           "vae_loss_first_last": [log[0][1], log[-1][1]], "critic_loss_first_last": [log[0][2], log[-1][2]]}
    if args.torch:
        k = min(args.updates, 100)
        for _ in range(5):
            ref.update(buf.sample(args.batch))
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(k):
            ref.update(buf.sample(args.batch))
        torch.cuda.synchronize()
        out["torch_eager_ms_per_update"] = 1e3 * (time.time() - t0) / k
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        buf.sample(1 << 20)
    e1.record()
    torch.cuda.synchronize()
    out["sample_1M_rows_ms"] = e0.elapsed_time(e1) / 50
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main()
