"""p50 latency of one device-resident mbik_solve_batch call per batch size, for both kernel mappings
(one thread per pose in lockstep CTAs vs. one warp per concurrently solvable segment).
    python profiles/run_latency.py [--rig humanoid22] [--sizes 32,1024,4096,...] [--calls 120]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE, MBIK_SCHED_SEGMENT_PARALLEL, MBIK_SCHED_THROUGHPUT

ap = argparse.ArgumentParser()
ap.add_argument("--rig", default="humanoid22")
ap.add_argument("--sizes", default="32,1024,4096,8192,16384,32768,65536")
ap.add_argument("--calls", type=int, default=120)
ap.add_argument("--json", default="")
a = ap.parse_args()
rig = rigs.RIGS[a.rig]()
R = BatchedIKRig(rig)
print(f"{a.rig}: sp_roles {R.info['sp_roles']}, sp_phases {R.info['sp_phases']}, estimated gain {R.info['sp_gain']:.2f}")
stream = torch.cuda.current_stream().cuda_stream
rows = []
for n in [int(x) for x in a.sizes.split(",")]:
    T = torch.from_numpy(rigs.random_targets(rig, 0, n)).cuda()
    outs = {}
    row = {"poses": n}
    for name, fl in (("throughput", MBIK_SCHED_THROUGHPUT), ("segment_parallel", MBIK_SCHED_SEGMENT_PARALLEL), ("auto", 0)):
        O = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device="cuda")
        lat = []
        for i in range(a.calls + 20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            R.solve_raw(n, T, O, device=0, flags=MBIK_IO_DEVICE | fl, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            if i >= 20:
                lat.append(e0.elapsed_time(e1))
        outs[name] = O
        row[name + "_p50_ms"] = float(np.median(lat))
    row["identical"] = bool(torch.equal(outs["throughput"].view(torch.int32), outs["segment_parallel"].view(torch.int32)))
    rows.append(row)
    print(json.dumps(row))
if a.json:
    json.dump({"rig": a.rig, "rows": rows}, open(a.json, "w"), indent=1)
