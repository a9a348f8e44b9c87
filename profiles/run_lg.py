"""L2-managed state experiment (MBIK_LG_HOT_FROM): kernel time and an output digest of one big device-resident batch.
    MBIK_LG_HOT_FROM=44 python profiles/run_lg.py --rig chain64 [--poses 75776]"""
import argparse
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE

ap = argparse.ArgumentParser()
ap.add_argument("--rig", default="chain64")
ap.add_argument("--poses", type=int, default=75776)
a = ap.parse_args()
rig = rigs.RIGS[a.rig]()
R = BatchedIKRig(rig)
T = torch.from_numpy(rigs.random_targets(rig, 0, a.poses)).cuda()
O = torch.empty((a.poses, rig.n_bones, 10), dtype=torch.float32, device="cuda")
ms = []
for i in range(4):
    O.zero_()
    R.solve_raw(a.poses, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ms.append(R.last_kernel_ms(0))
digest = hashlib.sha1(O.cpu().numpy().tobytes()).hexdigest()[:16]
print(f"{a.rig} hot_from={os.environ.get('MBIK_LG_HOT_FROM', '-')} poses={a.poses}: {min(ms[1:]):.2f} ms  ({a.poses / min(ms[1:]) / 1e3:.3f} M solves/s)  digest {digest}")
