"""Tiny driver for ncu / sanitizer runs: a few device-resident launches of the solve kernel.
    python profiles/run_kernel.py [--rig humanoid22] [--poses 131072] [--launches 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE

ap = argparse.ArgumentParser()
ap.add_argument("--rig", default="humanoid22")
ap.add_argument("--poses", type=int, default=131072)
ap.add_argument("--launches", type=int, default=3)
a = ap.parse_args()
rig = rigs.RIGS[a.rig]()
R = BatchedIKRig(rig)
T = torch.from_numpy(rigs.random_targets(rig, 0, a.poses)).cuda()
O = torch.empty((a.poses, rig.n_bones, 10), dtype=torch.float32, device="cuda")
for i in range(a.launches):
    R.solve_raw(a.poses, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print(f"launch {i}: {R.last_kernel_ms(0):.3f} ms  ({a.poses / R.last_kernel_ms(0) / 1e3:.2f} M solves/s)")
