"""Profiling driver: replay add / sample(philox) / sample(indices) and the single VecEnv step at 1,048,576 reactors."""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
ne, rows, B = 1 << 20, 64, 1 << 22
buf = pkg.GpuReplayBuffer(rows * ne, n_envs=ne, index_mode="philox")
o = torch.rand((ne, 4), device="cuda")
a = torch.rand((ne, 2), device="cuda")
r = torch.rand(ne, device="cuda")
d = torch.zeros(ne, dtype=torch.uint8, device="cuda")
env = pkg.GpuCSTRVecEnv(ne, seed=0, monitor=False)
env.reset()
act = torch.rand((ne, 2), device="cuda") * 2 - 1
bi = torch.randint(0, rows, (B,), device="cuda")
ei = torch.randint(0, ne, (B,), device="cuda")


def timed(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


t_add = timed(lambda: buf.add(o, o, a, r, d, None, timeouts=d), 64)
t_smp = timed(lambda: buf.sample(B), 10)
t_gat = timed(lambda: buf.gather(bi, ei), 10)
t_stp = timed(lambda: env.step_tensor(act), 20)
print(f"replay_add      {t_add*1e3:8.1f} us  {ne/t_add/1e6:8.1f} G transitions/s  algorithmic {ne*52/t_add/1e6:7.1f} GB/s  moved {ne*106/t_add/1e6:7.1f} GB/s")
print(f"replay_sample   {t_smp*1e3:8.1f} us  {B/t_smp/1e6:8.1f} G samples/s      algorithmic {B*116/t_smp/1e6:7.1f} GB/s (philox: no index read)")
print(f"replay_gather   {t_gat*1e3:8.1f} us  {B/t_gat/1e6:8.1f} G samples/s      algorithmic {B*116/t_gat/1e6:7.1f} GB/s")
print(f"vec_step_f32    {t_stp*1e3:8.1f} us  {ne/t_stp/1e6:8.1f} G env-steps/s   algorithmic {ne*70/t_stp/1e6:7.1f} GB/s")
