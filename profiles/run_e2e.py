"""Host-buffer (e2e) throughput of mbik_solve_batch for chunk-size tuning:
    MBIK_CHUNK_POSES=75776 python profiles/run_e2e.py [--poses 1048576] [--reps 5]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from many_bone_ik_b200 import BatchedIKRig, rigs

ap = argparse.ArgumentParser()
ap.add_argument("--poses", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rig = rigs.humanoid22()
R = BatchedIKRig(rig)
n = a.poses
T = torch.empty((n, rig.n_pins, 12), dtype=torch.float32, pin_memory=True)
for s in range(0, n, 1 << 16):
    e = min(n, s + (1 << 16))
    T[s:e] = torch.from_numpy(rigs.random_targets(rig, s, e - s))
O = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, pin_memory=True)
for _ in range(2):
    R.solve_raw(n, T.numpy(), O.numpy(), device=0)
t0 = time.perf_counter()
for _ in range(a.reps):
    R.solve_raw(n, T.numpy(), O.numpy(), device=0)
dt = (time.perf_counter() - t0) / a.reps
print(f"chunk={os.environ.get('MBIK_CHUNK_POSES', 'default')}: {dt * 1e3:.2f} ms per call, {n / dt / 1e6:.2f} M solves/s e2e")
