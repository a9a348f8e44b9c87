"""Binding soak: the reference module's own classes (oracle/_ref, built from /root/reference/src in the build container) with
ManyBoneIK3D::_process_modification replaced by the libmbik.so binding (many_bone_ik_b200/host/godot_module_binding.h), against
the same scenes solved by the reference's CPU solver -- on the parity soak's random rigs (tests/rig_cases.soak_rig): per-node
Binding (fresh node per pose / long-lived node, rest / perturbed start pose), and CrowdBinding scenes of several random rigs
over frames (one launch per rig and frame).  Every Skeleton3D must hold the same bits.
      python profiles/run_binding_soak.py [--rigs 200] [--crowds 20]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rig_cases  # noqa: E402
from many_bone_ik_b200 import rigs  # noqa: E402
from oracle import reference_py as Rf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rigs", type=int, default=200)
ap.add_argument("--crowds", type=int, default=20)
a = ap.parse_args()
if not (Rf.available() and os.path.exists(Rf.BINDING_LIB)):
    print("oracle/_ref is not built (needs the build container); nothing to do")
    sys.exit(2)


def same(x, y):
    return x.shape == y.shape and bool(np.all((x == y) | (np.isnan(x) & np.isnan(y))))


t0 = time.time()
fails = []
n_nodes = 0
for k in range(a.rigs):
    rig = rig_cases.soak_rig(k)
    n = 4
    T = rigs.random_targets(rig, 300 + k, n)
    start = rig_cases.perturbed_start_pose(rig, n, seed=k)
    for sp, rebuild in ((None, True), (start, False)) if k % 2 else ((start, True), (None, False)):
        ref_out, ref_st = Rf.solve_batch(rig, T, start_pose=sp, rebuild_each=rebuild)
        rc, out, st = Rf.binding_solve_batch(rig, T, start_pose=sp, rebuild_each=rebuild)
        n_nodes += n
        if rc != 0 or not same(out, ref_out):
            fails.append((k, "binding", rc, rebuild, sp is not None))
            print("MISMATCH rig", k, rig.name, "bones", rig.n_bones, "rc", rc, "rebuild", rebuild, "start", sp is not None, flush=True)
print(f"per-node binding: {a.rigs} rigs, {n_nodes} node-frames, {len(fails)} failures; {time.time() - t0:.0f} s", flush=True)

t1 = time.time()
crowd_fail = 0
for c in range(a.crowds):
    rng = np.random.default_rng(50 + c)
    ks = [int(x) for x in rng.choice(400, size=4, replace=False)]
    parts = [(rig_cases.soak_rig(kk), int(rng.integers(3, 20))) for kk in ks]
    frames = 3
    crowd = Rf.BindingCrowd()
    refs, Ts = [], []
    for j, (rig, n) in enumerate(parts):
        sp = rig_cases.perturbed_start_pose(rig, n, seed=c * 10 + j) if j % 2 else None
        crowd.add(rig, n, start_pose=sp)
        Tf = np.stack([rigs.random_targets(rig, 100 * c + 10 * f + j, n) for f in range(frames)])
        Ts.append(Tf)
        refs.append(Rf.solve_frames(rig, Tf, start_pose=sp, threads=8))
    for f in range(frames):
        rc, outs, launches = crowd.frame([T[f] for T in Ts])
        ok = rc == 0 and all(same(outs[j], refs[j]["out"][f]) for j in range(len(parts)))
        if not ok:
            crowd_fail += 1
            fails.append((c, "crowd", rc, f, ks))
            print("MISMATCH crowd", c, "frame", f, "rigs", ks, "rc", rc, "launches", launches, flush=True)
    crowd.close()
print(f"crowd binding: {a.crowds} scenes x 4 random rigs x 3 frames, {crowd_fail} failing frames; {time.time() - t1:.0f} s")
print(f"binding soak: {len(fails)} failures {fails[:8]}")
sys.exit(1 if fails else 0)
