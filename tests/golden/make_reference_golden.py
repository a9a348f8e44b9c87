"""Generates tests/golden/reference_solves.npz by running the REFERENCE MODULE'S OWN CODE: oracle/_ref/libmbik_ref.so
is every translation unit under /root/reference/src compiled unmodified (make -C oracle ref) over the engine
stand-in oracle/godot_shim/ (engine math = oracle/godot_math.h).  /root/reference does not exist on the GPU box, so
its outputs are frozen here as fixtures: the CPU suite checks the restatement oracle against them, the GPU suite
checks the CUDA path against them.

Contents, for every benchmark and edge rig (rigs.RIGS + rig_cases.EDGE_RIGS) and for 24 random rigs:
  <name>_targets  [4, n_pins, 12]   seeded targets (rigs.random_targets, poses 0..3)
  <name>_out      [4, n_bones, 10]  position / quaternion / scale handed to Skeleton3D by the reference
  <name>_local    [4, n_bones, 12]  raw local transforms of the IK bones after the solve
  <name>_status   [4]               non-finite-reset flag
  <name>_order / _dir / _twist / _cones   setup facts of the rig the reference builds (solve order of the bones,
                                    bone-direction and twist-axes bases, cone + tangent-circle geometry)
plus warm-start frames (humanoid22, quad80: frame 2 started from frame 1's locals), frame sequences of a long-lived node
(4 frames, every frame re-seeded from the skeleton as the previous write-back left it) and stage vectors (QCP fits,
kusudama point-in-limits, clamp, swing-twist on seeded random inputs).

    python tests/golden/make_reference_golden.py      (needs /root/reference; run in the build container)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from many_bone_ik_b200 import rigs  # noqa: E402
from oracle import reference_py as Rf  # noqa: E402
import rig_cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N_RANDOM_RIGS = 24


def all_cases():
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    for seed in range(N_RANDOM_RIGS):
        cases[f"random_rig_{seed}"] = (lambda s=seed: rig_cases.random_rig(s))
    return cases


def stage_inputs():
    """Seeded inputs of the stage-level vectors (shared with the tests)."""
    rng = np.random.default_rng(20260)
    qcp = []
    for k in range(24):
        n = int(rng.integers(1, 46))
        moved = rng.normal(size=(n, 3)).astype(np.float32)
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = rng.uniform(0, 2.5)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        Rm = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        target = (moved.astype(np.float64) @ Rm.T + rng.normal(size=(n, 3)) * (0.05 if k % 3 else 0.0) + (rng.normal(size=3) if k % 2 else 0)).astype(np.float32)
        w = rng.uniform(0.0, 1.0, n)
        if k % 5 == 0:
            w[rng.integers(0, n)] = 0.0
        qcp.append((moved, target, w.astype(np.float64), bool(k % 2)))
    kus = []
    for k in range(48):
        nc = int(rng.integers(1, 5))
        cones = np.concatenate([rng.normal(size=(nc, 3)), rng.uniform(0.05, 1.2, (nc, 1))], axis=1).astype(np.float32)
        point = rng.normal(size=3).astype(np.float32)
        kus.append((cones, point))
    quats = rng.normal(size=(32, 4)).astype(np.float32)
    quats /= np.linalg.norm(quats, axis=1, keepdims=True).astype(np.float32)
    cos_half = np.cos(rng.uniform(0.01, 1.5, 32) / 2)
    return qcp, kus, quats.astype(np.float32), cos_half


def main():
    assert Rf.source_present(), "needs the reference sources (/root/reference)"
    Rf.build()
    rc, text = Rf.run_doctests()
    assert rc == 0, text
    out = {}
    for name, f in all_cases().items():
        rig = f()
        T = rigs.random_targets(rig, 0, 4)
        o, loc, st = Rf.solve_batch(rig, T, want_local=True, rebuild_each=True)
        facts = Rf.rig_facts(rig)
        out[name + "_targets"] = T
        out[name + "_out"] = o
        out[name + "_local"] = loc
        out[name + "_status"] = st
        out[name + "_order"] = facts["bone_order"]
        out[name + "_dir"] = facts["dir_basis"]
        out[name + "_twist"] = facts["twist_basis"]
        out[name + "_cones"] = Rf.cone_geometry(rig)
    for name in ("humanoid22", "quad80"):
        rig = rigs.RIGS[name]()
        T2 = rigs.random_targets(rig, 1000, 4)
        o2, loc2, st2 = Rf.solve_batch(rig, T2, start_pose=out[name + "_local"], want_local=True)
        out[name + "_warm_targets"] = T2
        out[name + "_warm_out"] = o2
        out[name + "_warm_local"] = loc2
        out[name + "_warm_status"] = st2
    # a node living across frames (ref_solve_frames): frame f+1 starts from what the skeleton holds after frame f's write-back
    for name in ("humanoid22", "quad80", "chain_diverging"):
        rig = all_cases()[name]()
        Tf = np.stack([rigs.random_targets(rig, 5000 + 100 * f, 6) for f in range(4)])
        fr = Rf.solve_frames(rig, Tf)
        out[name + "_frames_targets"] = Tf
        out[name + "_frames_out"] = fr["out"]
        out[name + "_frames_skeleton"] = fr["skeleton"]
        out[name + "_frames_status"] = fr["status"]
    qcp, kus, quats, cos_half = stage_inputs()
    out["stage_qcp"] = np.stack([np.concatenate(Rf.qcp_weighted_superpose(m, t, w, tr)) for m, t, w, tr in qcp])
    out["stage_kusudama"] = np.stack([np.concatenate([p, [ib]]).astype(np.float32) for p, ib in (Rf.kusudama_point_in_limits(c, pt) for c, pt in kus)])
    out["stage_clamp"] = np.stack([Rf.clamp_to_cos_half_angle(q, c) for q, c in zip(quats, cos_half)])
    out["stage_swing_twist"] = np.stack([np.concatenate(Rf.swing_twist_y(q)) for q in quats])
    np.savez_compressed(os.path.join(HERE, "reference_solves.npz"), **out)
    print("wrote reference_solves.npz:", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "reference_solves.npz")), "bytes")


if __name__ == "__main__":
    main()
