"""Small solves of every rig family (default, stabilised, warm-start stream) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck  python profiles/sanitize_run.py
    compute-sanitizer --tool racecheck python profiles/sanitize_run.py [rig ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, IKStream, rigs

cases = dict(rigs.RIGS)
cases.update(rig_cases.EDGE_RIGS)
only = sys.argv[1:]
for name, f in cases.items():
    if only and name not in only:
        continue
    rig = f()
    R = BatchedIKRig(rig)
    for n in (33, 600):
        T = rigs.random_targets(rig, 0, n)
        for sched in ("throughput", "segment_parallel"):  # both kernel mappings (racecheck: shared-memory poses, team buffers)
            out, loc, st = R.solve(T, want_local=True, iterations=2, sched=sched)
            assert out.shape == (n, rig.n_bones, 10)
    S = IKStream(R, 64)
    S.submit(rigs.random_targets(rig, 0, 64), np.empty((64, rig.n_bones, 10), np.float32), None, iterations=1)
    S.sync()
    print("ok", name, flush=True)
