"""Profiling driver: the tape kernel variants at BASELINE configs[1] size (65,536 reactors x 400 intervals).
Usage: python profiles/run_tape.py {strict|fast|f64} [hbm|philox]"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

pkg = importlib.import_module("pytorch-rl-enhancedstablebaselines_b200")
kind = sys.argv[1] if len(sys.argv) > 1 else "strict"
src = sys.argv[2] if len(sys.argv) > 2 else "hbm"
n, T = 65536, 400
if kind == "f64":
    env = pkg.GpuCSTRVecEnv(n, dtype="fp64", seed=1, monitor=False)
    dt = torch.float64
else:
    env = pkg.GpuCSTRVecEnv(n, math=kind, seed=1, monitor=False)
    dt = torch.float32
env.reset()
acts = (torch.rand((T, n, 2), device="cuda", dtype=dt) * 2 - 1) if src == "hbm" else None
for _ in range(3):
    env.tape(T, acts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    env.tape(T, acts)
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"tape {kind} actions={src}: {ms*1e3:.1f} us/launch, {n*T/ms/1e6:.1f} G env-steps/s")
