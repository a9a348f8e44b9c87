"""Generates tests/golden/oracle_solves.npz and reference_kats.npz from the CPU oracle (the restatement; the fixtures
written by the reference module's own code are tests/golden/reference_solves.npz, make_reference_golden.py).  The fixtures freeze (a) the numeric known-answer values of
the reference's own doctest cases and (b) oracle solve outputs for a few seeded poses of every benchmark rig, so
that any later change to the oracle or the kernel that moves a result is caught.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from many_bone_ik_b200 import rigs  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
import rig_cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # (a) reference KAT inputs and the oracle's outputs on them
    q = np.array([0, 0, np.sqrt(2) / 2, np.sqrt(2) / 2], np.float32)
    moved = np.array([[4, 5, 6], [7, 8, 9], [1, 2, 3]], np.float32)

    def qxform(q, v):  # Quaternion::xform in float32
        u = q[:3].astype(np.float32)
        uv = np.cross(u, v).astype(np.float32)
        return (v + ((uv * q[3]) + np.cross(u, uv).astype(np.float32)) * np.float32(2)).astype(np.float32)

    target = np.stack([qxform(q, m) for m in moved])
    rot, tr = O.qcp_weighted_superpose(moved, target, [1, 1, 1], False)
    target_t = (moved + np.array([1, 2, 3], np.float32)).astype(np.float32)
    rot_t, tr_t = O.qcp_weighted_superpose(moved, target_t, [1, 1, 1], True)
    pt, ib = O.kusudama_point_in_limits([[0, 0, 1, np.float32(np.deg2rad(np.float32(30.0)))]], [1, 0, 0])
    np.savez(os.path.join(HERE, "reference_kats.npz"), qcp_moved=moved, qcp_target=target, qcp_rot=rot, qcp_expected=q,
             qcp_target_t=target_t, qcp_rot_t=rot_t, qcp_tr_t=tr_t, kus_point=pt, kus_in_bounds=np.float32(ib))
    # (b) frozen oracle solves
    out = {}
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    for name, f in cases.items():
        rig = f()
        n = 4
        T = rigs.random_targets(rig, 0, n)
        o, loc, st = O.solve_batch(rig, T, want_local=True, rebuild_each=True)
        out[name + "_targets"] = T
        out[name + "_out"] = o
        out[name + "_local"] = loc
        out[name + "_status"] = st
    np.savez_compressed(os.path.join(HERE, "oracle_solves.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
