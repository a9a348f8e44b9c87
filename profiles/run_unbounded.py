import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, rig_cases
from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE
for name, f in (("chain200", rig_cases.LARGE_RIGS["chain200"]), ("chain300", rig_cases.chain300)):
    rig = f(); R = BatchedIKRig(rig); n = 75776
    T = torch.from_numpy(rigs.random_targets(rig, 0, n)).cuda()
    O = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device="cuda")
    for i in range(3):
        R.solve_raw(n, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    R.solve_raw(n, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    b.record(); torch.cuda.synchronize()
    print(name, "iterations", rig.iterations, "%.2f ms per %d poses = %.3f M solves/s" % (a.elapsed_time(b), n, n / a.elapsed_time(b) / 1e3), "capacity", R.info["kernel_capacity"])
