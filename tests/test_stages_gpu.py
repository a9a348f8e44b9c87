"""Per-stage checks of the CUDA code (test plan items 1-2): the reference's own unit-test vectors
asserted directly on the kernel's device functions (mbik_stage_* probes), and randomized differential tests of
each stage against the oracle's restatement of the same reference function -- bit for bit."""
import ctypes as C

import numpy as np
import pytest

from many_bone_ik_b200 import BatchedIKRig, _capi, rigs
from many_bone_ik_b200.rigs import DEG, Rig, _xf
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
CMP_EPSILON = 1e-5


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def gpu_qcp(moved, target, weight, translate):
    lib = _capi.load_library()
    moved = np.ascontiguousarray(moved, np.float32)
    target = np.ascontiguousarray(target, np.float32)
    weight = np.ascontiguousarray(weight, np.float64)
    out = np.zeros(7, np.float32)
    rc = lib.mbik_stage_qcp(0, moved.shape[0], _p(moved), _p(target), _p(weight), int(bool(translate)), _p(out))
    assert rc == 0
    return out[:4].copy(), out[4:].copy()


def gpu_clamp(quats, cos_half):
    lib = _capi.load_library()
    quats = np.ascontiguousarray(quats, np.float32).reshape(-1, 4)
    cos_half = np.ascontiguousarray(cos_half, np.float64)
    out = np.zeros_like(quats)
    assert lib.mbik_stage_clamp(0, quats.shape[0], _p(quats), _p(cos_half), _p(out)) == 0
    return out


def cone_rig(cones):
    """Two-bone chain whose second bone carries a kusudama with `cones` ([(cx, cy, cz, radius)])."""
    rest = np.stack([_xf(np.eye(3), (0, 0, 0)), _xf(np.eye(3), (0, 0.3, 0)), _xf(np.eye(3), (0, 0.3, 0))]).astype(np.float32)
    r = Rig("cone_probe", ["a", "b", "c"], np.array([-1, 0, 1], np.int32), rest, iterations=1)
    r.pins = [dict(bone=2, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2))]
    r.constraints = [dict(bone=1, twist_from=0.0, twist_range=1.0, cones=[tuple(float(x) for x in c) for c in cones])]
    return r


def gpu_point_in_limits(cones, points):
    lib = _capi.load_library()
    R = BatchedIKRig(cone_rig(cones))
    points = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
    out = np.zeros((points.shape[0], 4), np.float32)
    assert lib.mbik_stage_point_in_limits(R.handle, 0, 1, points.shape[0], _p(points), _p(out)) == 0
    return out[:, :3].copy(), out[:, 3].copy()


def qxform(q, v):  # Quaternion::xform in float32
    u = q[:3].astype(np.float32)
    uv = np.cross(u, v).astype(np.float32)
    return (v + ((uv * q[3]) + np.cross(u, uv).astype(np.float32)) * np.float32(2)).astype(np.float32)


# ---- the reference's own test vectors, on the GPU stage code -----------------------------------------
def test_reference_qcp_weighted_superpose_on_gpu():
    """reference tests/test_qcp.h:40-57"""
    q = np.array([0, 0, np.sqrt(2) / 2, np.sqrt(2) / 2], np.float32)
    moved = np.array([[4, 5, 6], [7, 8, 9], [1, 2, 3]], np.float32)
    target = np.stack([qxform(q, m) for m in moved])
    rot, tr = gpu_qcp(moved, target, [1.0, 1.0, 1.0], False)
    assert np.all(np.abs(rot - q) < CMP_EPSILON)


def test_reference_qcp_weighted_translation_on_gpu():
    """reference tests/test_qcp.h:59-85"""
    moved = np.array([[4, 5, 6], [7, 8, 9], [1, 2, 3]], np.float32)
    target = (moved + np.array([1, 2, 3], np.float32)).astype(np.float32)
    rot, tr = gpu_qcp(moved, target, [1.0, 1.0, 1.0], True)
    assert np.all(np.abs(tr - np.array([1, 2, 3])) < CMP_EPSILON)
    assert np.all(np.abs(np.abs(rot) - np.array([0, 0, 0, 1])) < CMP_EPSILON)


def test_reference_kusudama_single_cone_cases_on_gpu():
    """reference tests/test_ik_kusudama_3d.h:38-156"""
    r30 = np.float32(np.deg2rad(np.float32(30.0)))
    pt, ib = gpu_point_in_limits([(0, 0, 1, r30)], [[1, 0, 0]])
    assert ib[0] == -1.0 and np.all(np.abs(pt[0] - np.array([0.5, 0.0, 0.8660254])) < CMP_EPSILON)
    p_in = np.array([0.0, 0.1, 1.0], np.float32)
    pt, ib = gpu_point_in_limits([(0, 0, 1, r30)], [p_in])
    assert ib[0] > 0 and np.allclose(pt[0], p_in / np.linalg.norm(p_in), atol=1e-6)
    pt, ib = gpu_point_in_limits([(0, 0, 1, 0.0)], [[1, 0, 0]])
    assert ib[0] < 0 and np.all(np.abs(pt[0] - np.array([0, 0, 1])) < 1e-4)


# ---- differential, stage by stage, against the oracle ------------------------------------------------
@pytest.mark.parametrize("translate", [False, True])
def test_qcp_stage_equals_oracle(translate):
    rng = np.random.default_rng(5 + int(translate))
    for trial in range(200):
        n = int(rng.integers(1, 64))
        moved = rng.normal(size=(n, 3)).astype(np.float32) * np.float32(rng.choice([0.01, 1.0, 30.0]))
        if trial % 5 == 0:  # near-exact fits: the adjugate collapses (qsqr < 1e-6 -> identity)
            ang = rng.uniform(0, 0.01)
            target = moved + rng.normal(size=(n, 3)).astype(np.float32) * np.float32(ang)
        else:
            target = rng.normal(size=(n, 3)).astype(np.float32)
        w = rng.choice([0.0, 0.25, 1.0, 0.7], size=n) if trial % 7 else np.ones(n)
        g_rot, g_tr = gpu_qcp(moved, target, w, translate)
        o_rot, o_tr = O.qcp_weighted_superpose(moved, target, w, translate)
        assert np.array_equal(g_rot, o_rot, equal_nan=True), (trial, n, g_rot, o_rot)
        assert np.array_equal(g_tr, o_tr, equal_nan=True), (trial, n, g_tr, o_tr)


def test_clamp_stage_equals_oracle():
    rng = np.random.default_rng(9)
    n = 2000
    q = rng.normal(size=(n, 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    q[:50, 3] = np.float32(1.0)  # previous_coefficient == 0 branch candidates
    q[50:100, :3] = 0
    ch = np.cos(rng.uniform(0.001, 3.1, size=n) / 2.0)
    got = gpu_clamp(q, ch)
    ref = np.stack([O.clamp_to_cos_half_angle(q[i], ch[i]) for i in range(n)])
    assert np.array_equal(got, ref, equal_nan=True)


@pytest.mark.parametrize("n_cones", [1, 2, 3, 4])
def test_point_in_limits_stage_equals_oracle(n_cones):
    rng = np.random.default_rng(20 + n_cones)
    for trial in range(12):
        cones = []
        base = rng.normal(size=3)
        for _ in range(n_cones):
            c = base + rng.normal(size=3) * 0.8
            c /= np.linalg.norm(c)
            cones.append((np.float32(c[0]), np.float32(c[1]), np.float32(c[2]), np.float32(rng.uniform(0.05, 1.2))))
        pts = rng.normal(size=(300, 3)).astype(np.float32)
        pts[:20] *= np.float32(1e-3)
        pts[20:40] = np.array([c[:3] for c in cones] * 20, np.float32)[:20]  # exactly on control points
        g_pt, g_ib = gpu_point_in_limits(cones, pts)
        for i in range(pts.shape[0]):
            o_pt, o_ib = O.kusudama_point_in_limits(np.array(cones, np.float32), pts[i])
            assert np.array_equal(g_pt[i], o_pt, equal_nan=True), (trial, i, g_pt[i], o_pt)
            assert g_ib[i] == o_ib, (trial, i, g_ib[i], o_ib)


def test_reference_cases_in_cpp_through_the_c_abi(tmp_path):
    """tests/cpp/reference_cases_gpu.cpp: the reference's doctest cases, same names / inputs / CHECKs, in C++ against
    libmbik.so (stage probes + the ManyBoneIK3D facade for the kusudama set-up)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "many_bone_ik_b200")
    exe = os.path.join(str(tmp_path), "refcases")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(root, "tests", "cpp", "reference_cases_gpu.cpp"),
                           "-L" + pkg, "-l:libmbik.so", "-Wl,-rpath," + pkg, "-ldl", "-lpthread", "-lrt"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
