"""Throughput of the stabilisation (STAB) kernel variants: python profiles/run_stab.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE

for f in (rig_cases.humanoid_stabilized, rig_cases.chain_multibone_root_stabilized):
    rig = f()
    R = BatchedIKRig(rig)
    n = 148 * 512 * 4
    T = torch.from_numpy(rigs.random_targets(rig, 0, n)).cuda()
    O = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device="cuda")
    for i in range(3):
        R.solve_raw(n, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    ms = R.last_kernel_ms(0)
    print(f"{rig.name}: {ms:.3f} ms  ({n / ms / 1e3:.2f} M solves/s)")
