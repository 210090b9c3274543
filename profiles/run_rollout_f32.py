"""Profiling driver: a few launches of the fp32 fused rollout (TD3 actor 4-400-300-2).  Usage: python profiles/run_rollout_f32.py [n_envs] [K]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
env = pkg.GpuCSTRVecEnv(ne, device=dev, seed=4, monitor=False)
env.reset()
buf = pkg.GpuReplayBuffer(8 * ne, device=dev, n_envs=ne, index_mode="philox")
torch.manual_seed(0)
lin = [torch.nn.Linear(4, 400), torch.nn.Linear(400, 300), torch.nn.Linear(300, 2)]
actor = pkg.ActorWeights(lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias, device=dev)
roll = pkg.FusedRollout(env, buf, actor, sigma=0.1, actor_mode="fp32")
for _ in range(3):
    roll.collect(K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
roll.collect(K)
e1.record()
torch.cuda.synchronize()
print("fp32 rollout transitions/s %.3e" % (K * ne / (e0.elapsed_time(e1) * 1e-3)))
