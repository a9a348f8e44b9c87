"""Properties of the Godot core/math shim under the oracle (oracle/godot_math.h).  The engine source is not part of
the reference tree, so the shim restates Godot 4.3/4.4 semantics from knowledge of the engine;
these tests check every function the solve path calls against float64 mathematics (what the function must compute,
to float32 accuracy) and against the documented engine conventions (row-major Basis, Hamilton product order, w-last
quaternions, shortest-arc special cases), independently of the oracle's IK code."""
import numpy as np
import pytest

from oracle import oracle_py as O

RNG = np.random.default_rng(1234)


def rand_quat():
    q = RNG.normal(size=4)
    return (q / np.linalg.norm(q)).astype(np.float32)


def quat_to_mat64(q):
    x, y, z, w = [float(v) for v in q]
    n = x * x + y * y + z * z + w * w
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
                     [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


def same_rotation(qa, qb, tol=2e-6):
    return min(np.abs(qa - qb).max(), np.abs(qa + qb).max()) < tol


def test_basis_from_quaternion_is_the_standard_rotation_matrix_rows():
    for _ in range(200):
        q = rand_quat()
        B = O.math_probe("basis_from_quat", q).reshape(3, 3).astype(np.float64)
        assert np.abs(B - quat_to_mat64(q)).max() < 1e-6          # Basis stores ROWS; xform(v) = rows . v
    # a non-unit quaternion still yields a pure rotation (s = 2 / |q|^2)
    q = rand_quat() * np.float32(3.0)
    B = O.math_probe("basis_from_quat", q).reshape(3, 3).astype(np.float64)
    assert np.abs(B @ B.T - np.eye(3)).max() < 1e-5


def test_quaternion_xform_matches_matrix_and_hamilton_convention():
    for _ in range(200):
        q, v = rand_quat(), RNG.normal(size=3).astype(np.float32)
        got = O.math_probe("quat_xform", q, v).astype(np.float64)
        assert np.abs(got - quat_to_mat64(q) @ v.astype(np.float64)).max() < 1e-5
    # 90 degrees about +Z maps +X to +Y (right-handed, w last)
    s = np.float32(np.sqrt(0.5))
    assert np.allclose(O.math_probe("quat_xform", [0, 0, s, s], [1, 0, 0]), [0, 1, 0], atol=1e-6)


def test_get_quaternion_inverts_basis_from_quaternion_on_every_shepperd_branch():
    # rotations by ~180 degrees about x, y, z and generic ones exercise the four trace / diagonal branches
    specials = [np.array([1, 0.01, 0.02, 0.01]), np.array([0.01, 1, 0.02, 0.01]), np.array([0.02, 0.01, 1, 0.01]), np.array([0.1, 0.2, 0.3, 1.0])]
    for q in specials + [rand_quat() for _ in range(300)]:
        q = (np.asarray(q, np.float64) / np.linalg.norm(q)).astype(np.float32)
        B = O.math_probe("basis_from_quat", q)
        assert same_rotation(O.math_probe("get_quaternion", B), q)
        assert same_rotation(O.math_probe("get_rotation_quaternion", B), q)


def test_orthonormalized_is_gram_schmidt_on_columns_x_first():
    for _ in range(200):
        M = (quat_to_mat64(rand_quat()) + RNG.normal(size=(3, 3)) * 0.05).astype(np.float32)
        B = O.math_probe("orthonormalized", M).reshape(3, 3).astype(np.float64)
        assert np.abs(B.T @ B - np.eye(3)).max() < 1e-5
        x = M[:, 0].astype(np.float64)
        assert np.abs(B[:, 0] - x / np.linalg.norm(x)).max() < 1e-6   # the x COLUMN is only normalised
        y = M[:, 1].astype(np.float64) - B[:, 0] * (B[:, 0] @ M[:, 1].astype(np.float64))
        assert np.abs(B[:, 1] - y / np.linalg.norm(y)).max() < 1e-5


def test_rotation_quaternion_of_a_mirrored_basis_flips_it_first():
    q = rand_quat()
    B = O.math_probe("basis_from_quat", q).reshape(3, 3)
    mirrored = (B * np.float32(-1.0)).astype(np.float32)                # det < 0 -> scaled by -1 before conversion
    assert same_rotation(O.math_probe("get_rotation_quaternion", mirrored), q)
    sc = O.math_probe("get_scale", (B @ np.diag([2.0, 3.0, 0.5]).astype(np.float32)).astype(np.float32))
    assert np.allclose(sc, [2.0, 3.0, 0.5], atol=1e-5)
    assert np.allclose(O.math_probe("get_scale", mirrored), [-1, -1, -1], atol=1e-6)   # sign of the determinant


def test_inverse_and_affine_inverse():
    for _ in range(100):
        M = (quat_to_mat64(rand_quat()) @ np.diag(RNG.uniform(0.5, 2.0, 3))).astype(np.float32)
        inv = O.math_probe("inverse", M).reshape(3, 3).astype(np.float64)
        assert np.abs(inv @ M.astype(np.float64) - np.eye(3)).max() < 1e-5
        o = RNG.normal(size=3).astype(np.float32)
        ai = O.math_probe("affine_inverse", M, o)
        Bi, oi = ai[:9].reshape(3, 3).astype(np.float64), ai[9:].astype(np.float64)
        p = RNG.normal(size=3)
        assert np.abs(Bi @ (M.astype(np.float64) @ p + o) + oi - p).max() < 1e-5


def test_basis_product_composes_right_to_left():
    a, b = rand_quat(), rand_quat()
    A, B = O.math_probe("basis_from_quat", a), O.math_probe("basis_from_quat", b)
    AB = O.math_probe("basis_mul", A, B).reshape(3, 3).astype(np.float64)
    assert np.abs(AB - quat_to_mat64(a) @ quat_to_mat64(b)).max() < 1e-5   # (A * B).xform(v) = A.xform(B.xform(v))


def test_shortest_arc_constructor():
    for _ in range(200):
        v0, v1 = RNG.normal(size=3).astype(np.float32) * np.float32(RNG.uniform(0.1, 5)), RNG.normal(size=3).astype(np.float32)
        q = O.math_probe("shortest_arc", v0, v1)
        assert abs(np.linalg.norm(q) - 1) < 1e-5                       # inputs are normalised first (Godot >= 4.3)
        r = O.math_probe("quat_xform", q, v0 / np.linalg.norm(v0))
        assert np.abs(r - v1 / np.linalg.norm(v1)).max() < 1e-4                # ill-conditioned near anti-parallel inputs
        axis = q[:3].astype(np.float64)
        assert abs(axis @ v0.astype(np.float64)) < 1e-4 * np.linalg.norm(v0) and abs(axis @ v1.astype(np.float64)) < 1e-4 * np.linalg.norm(v1)
    # parallel -> identity; anti-parallel -> half turn about a perpendicular axis
    assert np.array_equal(O.math_probe("shortest_arc", [0, 2, 0], [0, 5, 0]), np.array([0, 0, 0, 1], np.float32))
    q = O.math_probe("shortest_arc", [0, 1, 0], [0, -1, 0])
    assert q[3] == 0 and abs(np.linalg.norm(q[:3]) - 1) < 1e-6 and abs(q[1]) < 1e-6


def test_axis_angle_constructor_and_slerp():
    q = O.math_probe("quat_axis_angle", [0, 0, 1], [np.pi / 2])
    assert np.allclose(q, [0, 0, np.sqrt(0.5), np.sqrt(0.5)], atol=1e-6)
    assert np.array_equal(O.math_probe("quat_axis_angle", [0, 0, 0], [1.0]), np.zeros(4, np.float32))   # zero axis -> zero quaternion
    a, b = rand_quat(), rand_quat()
    A, B = O.math_probe("basis_from_quat", a), O.math_probe("basis_from_quat", b)
    assert np.abs(O.math_probe("basis_slerp", A, B, [0.0]) - A).max() < 2e-6       # weight 0 returns `from` (the solve's no-op)
    assert np.abs(O.math_probe("basis_slerp", A, B, [1.0]) - B).max() < 2e-6
    H = O.math_probe("basis_slerp", A, B, [0.5]).reshape(3, 3).astype(np.float64)
    assert np.abs(H @ H.T - np.eye(3)).max() < 1e-5


def test_basis_from_quaternion_and_scale_scales_the_columns():
    """Basis::set_quaternion_scale (what Skeleton3D::get_bone_pose() recomposes a pose from): R(q) * diag(scale), i.e. column
    j of the rotation scaled by scale[j]; get_rotation_quaternion / get_scale take it apart again."""
    rng = np.random.default_rng(12)
    for _ in range(20):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        sc = rng.uniform(0.3, 2.5, 3)
        R = O.math_probe("basis_from_quat", q).reshape(3, 3).astype(np.float64)
        B = O.math_probe("basis_from_quat_scale", q, sc).reshape(3, 3)
        assert np.allclose(B, R @ np.diag(sc), atol=2e-6)
        assert np.allclose(O.math_probe("get_scale", B), sc, atol=1e-5)
        assert same_rotation(O.math_probe("get_rotation_quaternion", B), q)
    ident = O.math_probe("basis_from_quat_scale", [0, 0, 0, 1], [1, 1, 1]).reshape(3, 3)
    assert np.array_equal(ident, np.eye(3, dtype=np.float32))
