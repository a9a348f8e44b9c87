"""A few device-resident launches of a rig beyond 128 solved bones (tests/rig_cases.py LARGE_RIGS), for timing / ncu.
    python profiles/run_large_rig.py [--rig big_tree240] [--poses 75776] [--launches 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE

ap = argparse.ArgumentParser()
ap.add_argument("--rig", default="big_tree240")
ap.add_argument("--poses", type=int, default=75776)
ap.add_argument("--launches", type=int, default=3)
a = ap.parse_args()
rig = rig_cases.LARGE_RIGS[a.rig]()
R = BatchedIKRig(rig)
T = torch.from_numpy(rigs.random_targets(rig, 0, a.poses)).cuda()
O = torch.empty((a.poses, rig.n_bones, 10), dtype=torch.float32, device="cuda")
for i in range(a.launches):
    R.solve_raw(a.poses, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print(f"{a.rig} launch {i}: {R.last_kernel_ms(0):.3f} ms  ({a.poses / R.last_kernel_ms(0) / 1e3:.3f} M solves/s)")
