"""Times the TD3 gradient step (cstr_td3_update) against a plain torch fp32 autograd implementation of the same update on the
same GPU (what the unmodified reference's TD3.train executes on a CUDA device), at the reference's default batch (256) and
the large-batch config (4096).   python profiles/run_td3.py [--batch 256 4096] [--steps 200] [--once]"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn as nn
import torch.nn.functional as F


def flops_per_sample(h1, h2):
    fa, fc = 4 * h1 + h1 * h2 + 2 * h2, 6 * h1 + h1 * h2 + h2
    return 2 * (2.5 * fa + 9 * fc)  # MACs x 2, policy_delay = 2


def mlp(i, o, squash, dev):
    layers = [nn.Linear(i, 400), nn.ReLU(), nn.Linear(400, 300), nn.ReLU(), nn.Linear(300, o)]
    return nn.Sequential(*(layers + ([nn.Tanh()] if squash else []))).to(dev)


class TorchTD3:
    """td3.py:162-206 in eager torch (fp32, autograd, torch.optim.Adam, polyak) — the comparison arm."""

    def __init__(self, dev):
        self.actor, self.actor_t = mlp(4, 2, True, dev), mlp(4, 2, True, dev)
        self.critics = nn.ModuleList([mlp(6, 1, False, dev), mlp(6, 1, False, dev)])
        self.critics_t = nn.ModuleList([mlp(6, 1, False, dev), mlp(6, 1, False, dev)])
        self.actor_t.load_state_dict(self.actor.state_dict())
        self.critics_t.load_state_dict(self.critics.state_dict())
        self.opt_a, self.opt_c = torch.optim.Adam(self.actor.parameters(), lr=1e-3), torch.optim.Adam(self.critics.parameters(), lr=1e-3)
        self.n = 0

    def update(self, b):
        self.n += 1
        with torch.no_grad():
            noise = (b.actions.clone().normal_(0, 0.2)).clamp(-0.5, 0.5)
            na = (self.actor_t(b.next_observations) + noise).clamp(-1, 1)
            q = torch.cat([c(torch.cat([b.next_observations, na], 1)) for c in self.critics_t], 1)
            target = b.rewards + (1 - b.dones) * 0.99 * q.min(1, keepdim=True)[0]
        loss = sum(F.mse_loss(c(torch.cat([b.observations, b.actions], 1)), target) for c in self.critics)
        self.opt_c.zero_grad()
        loss.backward()
        self.opt_c.step()
        if self.n % 2 == 0:
            al = -self.critics[0](torch.cat([b.observations, self.actor(b.observations)], 1)).mean()
            self.opt_a.zero_grad()
            al.backward()
            self.opt_a.step()
            with torch.no_grad():
                for src, dst in ((self.critics, self.critics_t), (self.actor, self.actor_t)):
                    for p, t in zip(src.parameters(), dst.parameters()):
                        t.mul_(1 - 0.005)
                        torch.add(t, p, alpha=0.005, out=t)


def timed(fn, steps):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[256, 4096])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--once", action="store_true", help="few launches only (for ncu)")
    ap.add_argument("--no-torch", action="store_true")
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "tensor", "bf16"])
    ap.add_argument("--sac", action="store_true", help="time the SAC gradient step (cstr_sac_update, [256,256] nets) instead")
    ap.add_argument("--graph", action="store_true", help="time train(graph=True): one CUDA-graph launch per policy_delay updates")
    args = ap.parse_args()
    pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    n_envs = 65536
    buf = pkg.GpuReplayBuffer(16 * n_envs, device=dev, n_envs=n_envs, index_mode="philox")
    buf.records.uniform_(-1, 1)
    buf.records[..., 11:13] = 0
    buf.pos, buf.full = 0, True
    out = {}
    for B in args.batch if not args.sac else []:
        torch.manual_seed(0)
        eng = pkg.FusedTD3Update([400, 300], B, device=dev, gemm=args.gemm)
        ref = TorchTD3(dev)
        eng.adopt_modules(mlp(4, 2, True, dev), [mlp(6, 1, False, dev), mlp(6, 1, False, dev)], mlp(4, 2, True, dev),
                          [mlp(6, 1, False, dev), mlp(6, 1, False, dev)])
        steps = 4 if args.once else args.steps
        if args.graph:
            ms = timed(lambda: eng.train(2, buf, B, graph=True), steps // 2) / 2
        else:
            ms = timed(lambda: eng.update(buf.sample(B)), steps)
        row = {"fused_ms_per_update": ms, "fused_updates_per_s": 1e3 / ms, "fused_samples_per_s": B * 1e3 / ms,
               "algorithmic_tflops": flops_per_sample(400, 300) * B / (ms * 1e-3) / 1e12}
        if not args.no_torch and not args.once:
            ms_t = timed(lambda: ref.update(buf.sample(B)), steps)
            row.update(torch_eager_ms_per_update=ms_t, speedup_vs_torch_eager=ms_t / ms)
        out[f"batch_{B}"] = row
    for B in args.batch if args.sac else []:
        eng = pkg.FusedSACUpdate([256, 256], B, device=dev, gemm=args.gemm)
        eng.params[:eng._ent_offset].normal_(0, 0.05)
        eng.targets.copy_(eng.params)
        if args.graph:
            ms = timed(lambda: eng.train(2, buf, B, graph=True), args.steps // 2) / 2
        else:
            ms = timed(lambda: eng.update(buf.sample(B)), 4 if args.once else args.steps)
        out[f"sac_batch_{B}"] = {"fused_ms_per_update": ms, "fused_updates_per_s": 1e3 / ms}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
