"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/_build/liboracle.so (the CPU restatement of the
reference solver).  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from many_bone_ik_b200._capi import RigDesc, rig_to_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle.so")
KAT = os.path.join(_HERE, "_build", "oracle_kat")


def build(force=False):
    if force or not (os.path.exists(LIB) and os.path.exists(KAT)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.orc_solve_batch.argtypes = [C.POINTER(RigDesc), C.c_size_t, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_uint, vp]
        L.orc_rig_facts.argtypes = [C.POINTER(RigDesc), vp, C.c_int32, vp, vp, vp]
        L.orc_step_weights.argtypes = [C.POINTER(RigDesc), C.c_int32, vp, C.c_int32]
        L.orc_cone_geometry.argtypes = [C.POINTER(RigDesc), vp, C.c_int32]
        L.orc_qcp_weighted_superpose.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        L.orc_kusudama_point_in_limits.argtypes = [vp, C.c_int, vp, vp]
        L.orc_clamp_to_cos_half_angle.argtypes = [vp, C.c_double, vp]
        L.orc_swing_twist_y.argtypes = [vp, vp]
        L.orc_math_probe.argtypes = [C.c_int, vp, vp]
        L.orc_recompose_pose.argtypes = [C.c_size_t, vp, vp]
        L.orc_recompose_pose.restype = None
        L.orc_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hardware_threads():
    return int(lib().orc_hardware_threads())


def solve_batch(rig, targets, start_pose=None, iterations=-1, threads=1, rebuild_each=False, want_local=False, want_counters=False):
    """Returns (out_pose [n, n_bones, 10], [out_local [n, n_bones, 12]], status [n], [counters])."""
    L = lib()
    desc, keep = rig_to_desc(rig)
    targets = np.ascontiguousarray(targets, np.float32)
    n = targets.shape[0]
    assert targets.shape == (n, rig.n_pins, 12)
    if start_pose is not None:
        start_pose = np.ascontiguousarray(start_pose, np.float32)
        assert start_pose.shape == (n, rig.n_bones, 12)
    out = np.zeros((n, rig.n_bones, 10), np.float32)
    loc = np.zeros((n, rig.n_bones, 12), np.float32) if want_local else None
    st = np.zeros(n, np.uint32)
    cnt = np.zeros(3, np.int64)
    rc = L.orc_solve_batch(C.byref(desc), n, _p(targets), _p(start_pose), _p(out), _p(loc), _p(st), int(iterations), int(threads),
                           1 if rebuild_each else 0, _p(cnt))
    assert rc == 0
    res = [out]
    if want_local:
        res.append(loc)
    res.append(st)
    if want_counters:
        res.append(dict(bone_steps=int(cnt[0]), swing_calls=int(cnt[1]), swing_rectified=int(cnt[2])))
    return tuple(res)


def recompose_pose(rig, out_pose, start_pose=None):
    """The skeleton state after a frame's write-back, [n, n_bones, 12]: for the bones of bone_list what
    Skeleton3D::get_bone_pose() returns once out_pose's position / rotation / scale are in the skeleton
    (Transform3D(Basis(rotation, scale), position)); the other bones keep start_pose (default: rest)."""
    out_pose = np.ascontiguousarray(out_pose, np.float32)
    n = out_pose.shape[0]
    assert out_pose.shape == (n, rig.n_bones, 10)
    rec = np.zeros((n, rig.n_bones, 12), np.float32)
    lib().orc_recompose_pose(n * rig.n_bones, _p(out_pose), _p(rec))
    base = np.ascontiguousarray(start_pose, np.float32).copy() if start_pose is not None else np.broadcast_to(
        np.asarray(rig.rest_local, np.float32).reshape(1, rig.n_bones, 12), (n, rig.n_bones, 12)).copy()
    solved = rig_facts(rig)["bone_order"]
    base[:, solved] = rec[:, solved]
    return base


def rig_facts(rig):
    L = lib()
    desc, keep = rig_to_desc(rig)
    cap = rig.n_bones
    order = np.full(cap, -1, np.int32)
    nseg = np.zeros(1, np.int32)
    dirb = np.zeros((cap, 9), np.float32)
    twb = np.zeros((cap, 9), np.float32)
    n = L.orc_rig_facts(C.byref(desc), _p(order), cap, _p(nseg), _p(dirb), _p(twb))
    return dict(bone_order=order[:n].copy(), n_segments=int(nseg[0]), dir_basis=dirb[:n].copy(), twist_basis=twb[:n].copy())


def step_weights(rig, step):
    L = lib()
    desc, keep = rig_to_desc(rig)
    w = np.zeros(256, np.float64)
    n = L.orc_step_weights(C.byref(desc), int(step), _p(w), 256)
    return w[:max(n, 0)].copy()


def cone_geometry(rig):
    L = lib()
    desc, keep = rig_to_desc(rig)
    nc = sum(len(c["cones"]) for c in rig.constraints)
    out = np.zeros((max(nc, 1), 9), np.float32)
    n = L.orc_cone_geometry(C.byref(desc), _p(out), nc)
    return out[:n].copy()


def qcp_weighted_superpose(moved, target, weight, translate):
    L = lib()
    moved = np.ascontiguousarray(moved, np.float32)
    target = np.ascontiguousarray(target, np.float32)
    weight = np.ascontiguousarray(weight, np.float64)
    out = np.zeros(7, np.float32)
    L.orc_qcp_weighted_superpose(_p(moved), _p(target), _p(weight), moved.shape[0], int(bool(translate)), _p(out))
    return out[:4].copy(), out[4:].copy()


def kusudama_point_in_limits(cones, point):
    L = lib()
    cones = np.ascontiguousarray(cones, np.float32).reshape(-1, 4)
    point = np.ascontiguousarray(point, np.float32)
    out = np.zeros(4, np.float32)
    L.orc_kusudama_point_in_limits(_p(cones), cones.shape[0], _p(point), _p(out))
    return out[:3].copy(), float(out[3])


def clamp_to_cos_half_angle(q, cos_half):
    L = lib()
    q = np.ascontiguousarray(q, np.float32)
    out = np.zeros(4, np.float32)
    L.orc_clamp_to_cos_half_angle(_p(q), float(cos_half), _p(out))
    return out


def swing_twist_y(q):
    L = lib()
    q = np.ascontiguousarray(q, np.float32)
    out = np.zeros(8, np.float32)
    L.orc_swing_twist_y(_p(q), _p(out))
    return out[:4].copy(), out[4:].copy()


MATH_OPS = dict(basis_from_quat=0, get_quaternion=1, get_rotation_quaternion=2, orthonormalized=3, inverse=4, shortest_arc=5,
                quat_xform=6, basis_slerp=7, quat_axis_angle=8, affine_inverse=9, basis_mul=10, get_scale=11, basis_from_quat_scale=12)
_MATH_OUT = {0: 9, 1: 4, 2: 4, 3: 9, 4: 9, 5: 4, 6: 3, 7: 9, 8: 4, 9: 12, 10: 9, 11: 3, 12: 9}


def math_probe(op, *args):
    """One function of the Godot core/math shim (oracle/godot_math.h) on float32 inputs."""
    L = lib()
    code = MATH_OPS[op]
    inp = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in args]), np.float32)
    out = np.zeros(_MATH_OUT[code], np.float32)
    assert L.orc_math_probe(code, _p(inp), _p(out)) == 0
    return out


def run_kat():
    build()
    r = subprocess.run([KAT], capture_output=True, text=True)
    return r.returncode, r.stdout
