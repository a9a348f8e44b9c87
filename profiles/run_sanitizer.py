"""Every kernel instantiation family once, at a batch small enough for compute-sanitizer (run_sanitizer.sh wraps this in
memcheck / racecheck / initcheck / synccheck).  Checks only self-consistency (both mappings give the same bits, repeats
are bitwise equal); parity with the oracle is the job of tests/.

    python profiles/run_sanitizer.py [case ...]        cases: small sp stab lims glw glw_stab glw_lims dyn stream"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import limit_set_cases as LS  # noqa: E402
import rig_cases  # noqa: E402
from many_bone_ik_b200 import BatchedIKRig, IKStream, rigs  # noqa: E402


def same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def both_mappings(rig, n, seed=1, start=False):
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, seed, n)
    sp = rig_cases.perturbed_start_pose(rig, n, seed=seed + 1) if start else None
    a = R.solve(T, start_pose=sp, want_local=True, sched="throughput")
    b = R.solve(T, start_pose=sp, want_local=True, sched="segment_parallel")
    c = R.solve(T, start_pose=sp, want_local=True)
    for x, y, z in zip(a, b, c):
        assert same(x, y) and same(x, z), rig.name
    return R, T, a


def case_small():
    for mk in (rigs.humanoid22, rig_cases.two_roots, rig_cases.star_mixed_pins, rig_cases.scaled_bones, rig_cases.no_pins):
        both_mappings(mk(), 333, start=True)
    both_mappings(rig_cases.random_rig(10), 200)
    both_mappings(rig_cases.random_rig(3), 200)


def case_sp():
    for mk in (rigs.humanoid22, rigs.quad80, rig_cases.big_tree120):
        both_mappings(mk(), 97)


def case_stab():
    for mk in (rig_cases.humanoid_stabilized, rig_cases.two_roots_stabilized, rig_cases.chain_multibone_root_stabilized,
               rig_cases.star_stabilized_constraint_mode):
        both_mappings(mk(), 150)


def case_lims():
    for mk, n in ((rigs.humanoid22, 300), (rig_cases.humanoid_stabilized, 100), (rigs.quad80, 100)):
        rig = mk()
        R = BatchedIKRig(rig)
        sets = LS.variants(rig, 3)
        h = R.create_limit_sets(sets)
        T = rigs.random_targets(rig, 4, n)
        idx = (np.arange(n) % 3).astype(np.int32)
        a, _ = R.solve_with_limits(h, idx, T, sched="throughput")
        b, _ = R.solve_with_limits(h, idx, T)
        assert same(a, b), rig.name
        R.destroy_limit_sets(h)


def glw_rig(rig, n):
    """sched="throughput" on a 64-bone-and-up variant = the streamed-walk instantiation at any batch size"""
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 7, n)
    a = R.solve(T, want_local=True, sched="throughput")
    b = R.solve(T, want_local=True)  # segment-parallel where the rig has a schedule for it
    c = R.solve(T[-40:], want_local=True, sched="throughput")
    for x, y, z in zip(a, b, c):
        assert same(x, y) and same(x[-40:], z), rig.name


def case_glw():
    for mk in (rigs.chain64, rigs.quad80, rig_cases.chain150, rig_cases.chain200, rig_cases.big_tree240):
        glw_rig(mk(), 160)


def case_glw_stab():
    glw_rig(rig_cases.chain64_stabilized(), 100)


def case_glw_lims():
    rig = rigs.chain64()
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 2)
    h = R.create_limit_sets(sets)
    T = rigs.random_targets(rig, 5, 100)
    idx = (np.arange(100) % 2).astype(np.int32)
    a, _ = R.solve_with_limits(h, idx, T, sched="throughput")
    b, _ = R.solve_with_limits(h, idx, T, sched="throughput")
    assert same(a, b)
    R.destroy_limit_sets(h)


def case_dyn():
    rig = rig_cases.chain300()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 3, 70)
    a = R.solve(T, want_local=True)
    b = R.solve(T, want_local=True, start_pose=rig_cases.perturbed_start_pose(rig, 70, seed=8))
    assert not same(a[0], b[0])
    rig.stabilization_passes = 2
    rig.name = "chain300_stabilized"
    R2 = BatchedIKRig(rig)
    R2.solve(T[:40])
    sets = LS.variants(rig_cases.chain300(), 2)
    h = R.create_limit_sets(sets)
    R.solve_with_limits(h, (np.arange(40) % 2).astype(np.int32), T[:40])
    R.destroy_limit_sets(h)


def case_stream():
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 300
    for solved_only in (False, True):
        S = IKStream(R, n, device=0, solved_only=solved_only)
        out = np.empty((n, R.info["n_solved"] if solved_only else rig.n_bones, 10), np.float32)
        st = np.empty(n, np.uint32)
        for f in range(3):
            S.submit(rigs.random_targets(rig, 20 + f, n), out, st)
            S.sync()
        S.read_local()
        S.close()
    T = rigs.random_targets(rig, 1, n)
    a, _ = R.solve(T, solved_only=True)
    b, _ = R.solve(T)
    assert same(a, b[:, R.bone_order()])
    R.solve(T, newton_iters=3)


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}
if __name__ == "__main__":
    for name in sys.argv[1:] or list(CASES):
        CASES[name]()
        print("ok", name, flush=True)
