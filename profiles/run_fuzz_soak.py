"""Parity soak beyond the test suite's seeds: random rigs (tests/rig_cases.random_rig, seeds 1000...) solved in every kernel mapping
and compared bit for bit with the oracle; every 8th rig is a dense large one (64 ... 520 bones, many pins).  Prints one line per
failure and a summary.      python profiles/run_fuzz_soak.py [--rigs 300] [--poses 48]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rig_cases  # noqa: E402
from many_bone_ik_b200 import BatchedIKRig, MbikError, rigs  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rigs", type=int, default=300)
ap.add_argument("--poses", type=int, default=48)
ap.add_argument("--first", type=int, default=0)
ap.add_argument("--extreme", action="store_true", help="targets scaled by 1e5 ... 1e37 / 1e-30, 12 iterations: poses overflow to Inf / NaN part-way")
a = ap.parse_args()


def same(x, y):
    return x.shape == y.shape and bool(np.all((x == y) | (np.isnan(x) & np.isnan(y))))


t0 = time.time()
n_ok = n_fail = n_rejected = 0
variants = {}
for k in range(a.first, a.first + a.rigs):
    seed = 1000 + k
    rig = rig_cases.soak_rig(k)
    if a.extreme:
        rig.iterations = 12
    try:
        R = BatchedIKRig(rig)
    except MbikError as e:
        n_rejected += 1
        print("rejected", seed, e, flush=True)
        continue
    variants[R.info["kernel_capacity"]] = variants.get(R.info["kernel_capacity"], 0) + 1
    T = rigs.random_targets(rig, seed, a.poses)
    if a.extreme:
        erng = np.random.default_rng(seed)
        T[:, :, 9:] *= np.float32(10.0) ** erng.integers(5, 38, size=(a.poses, 1, 1)).astype(np.float32)
        if k % 3 == 0:
            T[:, :, :9] *= np.float32(10.0) ** erng.integers(0, 20, size=(a.poses, 1, 1)).astype(np.float32)
        if k % 4 == 0:
            T[::5] = np.float32(1e-30) * T[::5]
        if k % 5 == 0:
            T[1::7, :, 3] = np.nan
    start = rig_cases.perturbed_start_pose(rig, a.poses, seed=seed) if k % 3 == 0 else None
    ref = O.solve_batch(rig, T, start_pose=start, want_local=True, threads=8)
    bad = []
    for sched in ("throughput", "segment_parallel", "auto"):
        got = R.solve(T, start_pose=start, want_local=True, sched=sched)
        if not (same(got[0], ref[0]) and same(got[1], ref[1]) and np.array_equal(got[2], ref[2])):
            bad.append(sched)
    if bad:
        n_fail += 1
        print("MISMATCH seed", seed, "bones", rig.n_bones, "solved", R.info["n_solved"], "mappings", bad, flush=True)
    else:
        n_ok += 1
print(f"fuzz soak{' (extreme inputs)' if a.extreme else ''}: {n_ok} rigs bit-identical to the oracle in all three mappings, {n_fail} mismatching, {n_rejected} rejected by the flattener; "
      f"{a.poses} poses each; rigs per kernel capacity {dict(sorted(variants.items()))}; {time.time() - t0:.0f} s")
sys.exit(1 if n_fail else 0)
