import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE
for name in ("chain64", "quad80"):
    rig = rigs.RIGS[name](); R = BatchedIKRig(rig); n = 4096
    T = torch.from_numpy(rigs.random_targets(rig, 0, n)).cuda()
    O = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device="cuda")
    lat = []
    for i in range(40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); R.solve_raw(n, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream); b.record(); torch.cuda.synchronize()
        if i >= 10: lat.append(a.elapsed_time(b))
    print(name, "p50(4096) %.3f ms" % np.median(lat))
