"""CPU tests of the ORACLE (test infrastructure): the reference's own doctest known-answer cases and the
frozen golden fixtures.  No GPU, no product code under test here except the rig generators.

Reference KATs restated (values quoted from the reference test headers, /root/reference/tests/):
  test_qcp.h:40-57      Weighted Superpose      expected q = (0, 0, sqrt(2)/2, sqrt(2)/2) within CMP_EPSILON
  test_qcp.h:59-85      Weighted Translation    expected translation (1, 2, 3), identity rotation
  test_ik_kusudama_3d.h:127-156  point (1,0,0) vs 30 deg cone about +Z -> (0.5, 0, 0.8660254), bounds == -1
"""
import os

import numpy as np
import pytest

import rig_cases
from many_bone_ik_b200 import rigs
from oracle import oracle_py as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CMP_EPSILON = 1e-5


def test_reference_doctest_cases_pass_in_the_oracle_binary():
    rc, out = O.run_kat()
    assert rc == 0, out
    assert "0 failed" in out
    assert out.count("PASS") >= 15


def test_qcp_weighted_superpose_kat():
    k = np.load(os.path.join(GOLD, "reference_kats.npz"))
    rot, tr = O.qcp_weighted_superpose(k["qcp_moved"], k["qcp_target"], [1, 1, 1], False)
    expected = np.array([0, 0, np.sqrt(2) / 2, np.sqrt(2) / 2])
    assert np.all(np.abs(rot - expected) < CMP_EPSILON)  # the reference's own assertion
    assert np.array_equal(rot, k["qcp_rot"])               # frozen bits


def test_qcp_weighted_translation_kat():
    k = np.load(os.path.join(GOLD, "reference_kats.npz"))
    rot, tr = O.qcp_weighted_superpose(k["qcp_moved"], k["qcp_target_t"], [1, 1, 1], True)
    assert np.all(np.abs(tr - np.array([1, 2, 3])) < CMP_EPSILON)
    assert np.all(np.abs(np.abs(rot) - np.array([0, 0, 0, 1])) < CMP_EPSILON)
    assert np.array_equal(rot, k["qcp_rot_t"]) and np.array_equal(tr, k["qcp_tr_t"])


def test_kusudama_single_cone_kat():
    r30 = np.float32(np.deg2rad(np.float32(30.0)))
    pt, ib = O.kusudama_point_in_limits([[0, 0, 1, r30]], [1, 0, 0])
    assert ib == -1.0
    assert np.all(np.abs(pt - np.array([0.5, 0.0, 0.8660254])) < CMP_EPSILON)
    # inside the cone: point returned unchanged, bounds > 0 (test_ik_kusudama_3d.h:38-64)
    p_in = np.array([0.0, 0.1, 1.0], np.float32)
    pt2, ib2 = O.kusudama_point_in_limits([[0, 0, 1, r30]], p_in)
    assert ib2 > 0
    assert np.allclose(pt2, p_in / np.linalg.norm(p_in), atol=1e-6)
    # radius ~ 0: outside -> returns the control point (test_ik_kusudama_3d.h:96-124)
    pt3, ib3 = O.kusudama_point_in_limits([[0, 0, 1, 0.0]], [1, 0, 0])
    assert ib3 < 0
    assert np.all(np.abs(pt3 - np.array([0, 0, 1])) < 1e-4)


def test_clamp_to_cos_half_angle_properties():
    rng = np.random.default_rng(0)
    for _ in range(200):
        q = rng.normal(size=4)
        q = (q / np.linalg.norm(q)).astype(np.float32)
        ch = float(np.cos(rng.uniform(0.01, 1.5) / 2))
        out = O.clamp_to_cos_half_angle(q, ch)
        assert out[3] >= min(ch, abs(q[3])) - 1e-6          # rotation angle limited to 2*acos(ch)
        assert abs(np.linalg.norm(out) - 1.0) < 1e-5          # unit in exact arithmetic (not renormalised)
        if abs(q[3]) >= ch:                                   # inside the limit: only the sign flip
            assert np.array_equal(out, q if q[3] >= 0 else -q)


def test_swing_twist_recomposes():
    rng = np.random.default_rng(1)

    def qmul(a, b):
        ax, ay, az, aw = a
        bx, by, bz, bw = b
        return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                         aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])

    for _ in range(100):
        q = rng.normal(size=4)
        q = (q / np.linalg.norm(q)).astype(np.float32)
        sw, tw = O.swing_twist_y(q)
        assert abs(tw[0]) < 1e-6 and abs(tw[2]) < 1e-6       # twist is about +Y
        r = qmul(sw.astype(np.float64), tw.astype(np.float64))
        qq = q if q[3] >= 0 else -q
        assert min(np.abs(r - qq).max(), np.abs(r + qq).max()) < 1e-5


ALL_RIGS = dict(rigs.RIGS)
ALL_RIGS.update(rig_cases.EDGE_RIGS)


@pytest.mark.parametrize("name", sorted(ALL_RIGS))
def test_oracle_reproduces_frozen_solves(name):
    """The golden file was written by tests/golden/make_golden.py; any change of the oracle (or of the rig
    generators / the target RNG) that moves a single bit of a solve fails here."""
    g = np.load(os.path.join(GOLD, "oracle_solves.npz"))
    rig = ALL_RIGS[name]()
    T = rigs.random_targets(rig, 0, 4)
    assert np.array_equal(T, g[name + "_targets"]), "target generator changed"
    out, loc, st = O.solve_batch(rig, T, want_local=True, threads=2)
    assert np.array_equal(loc, g[name + "_local"], equal_nan=True)
    assert np.array_equal(out, g[name + "_out"], equal_nan=True)
    assert np.array_equal(st, g[name + "_status"])


def test_oracle_persistent_instance_equals_rebuild_each_pose():
    """solve_batch reuses one ManyBoneIK3D per thread and re-seeds it per pose; that must equal a fresh
    _bone_list_changed() rebuild per pose (what the C ABI's semantics are defined as)."""
    for f in (rigs.humanoid22, rig_cases.star_mixed_pins):
        rig = f()
        T = rigs.random_targets(rig, 100, 6)
        a = O.solve_batch(rig, T, want_local=True, rebuild_each=False)
        b = O.solve_batch(rig, T, want_local=True, rebuild_each=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y, equal_nan=True)


def test_oracle_thread_partition_invariance():
    rig = rigs.humanoid22()
    T = rigs.random_targets(rig, 0, 37)
    a = O.solve_batch(rig, T, threads=1)
    b = O.solve_batch(rig, T, threads=5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_solved_locals_stay_rigid():
    """The solve only ever composes rotations (and root-segment translations): every solved local basis stays
    orthonormal to float32 accuracy on the benchmark rigs.  (Whether the effector error shrinks is NOT asserted:
    the reference takes target headings relative to the effector's own bone, src/ik_effector_3d.cpp:97, so its
    fits are not the textbook EWBIK ones -- parity, not quality, is the contract)"""
    for f in (rigs.humanoid22, rigs.quad80):
        rig = f()
        T = rigs.random_targets(rig, 0, 8)
        out, loc, st = O.solve_batch(rig, T, want_local=True, threads=4)
        B = loc[..., :9].reshape(-1, 3, 3).astype(np.float64)
        assert np.all(np.isfinite(B))
        err = np.abs(B @ B.transpose(0, 2, 1) - np.eye(3)).max()
        assert err < 1e-4, err
        assert np.all(st == 0)
