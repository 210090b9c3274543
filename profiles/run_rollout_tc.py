"""Profiling driver: a few launches of the tcgen05 fused rollout (TD3 actor 4-400-300-2) on N reactors.
Usage: python profiles/run_rollout_tc.py [n_envs] [K] [launches]   (run plain first, then under ncu)"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 4
torch.manual_seed(0)
lin = [torch.nn.Linear(4, 400), torch.nn.Linear(400, 300), torch.nn.Linear(300, 2)]
actor = pkg.ActorWeights(lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias)
env = pkg.GpuCSTRVecEnv(n, seed=1, monitor=False)
env.reset()
buf = pkg.GpuReplayBuffer(2 * K * n, n_envs=n)
roll = pkg.FusedRollout(env, buf, actor, sigma=0.1, actor_mode="tc")
for _ in range(2):
    roll.collect(K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(launches):
    roll.collect(K)
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / launches
print(f"n={n} K={K}: {ms:.3f} ms/launch, {n * K / ms / 1e6:.1f} G transitions/s, {n * K * 244400 / ms / 1e9:.1f} TFLOP/s (actor)")
