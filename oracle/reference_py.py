"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/_ref/libmbik_ref.so: the reference module's OWN
sources (/root/reference/src, compiled unmodified by `make -C oracle ref`) driven headless over the engine
stand-in oracle/godot_shim/.  Same call shapes as oracle_py (the restatement), so tests can run one against
the other.  Imported only by tests/, tests/golden/make_reference_golden.py and bench.py's cpu_baseline /
--impl reference legs; never by the product package.

/root/reference does not exist on the GPU box: there the prebuilt oracle/_ref/ files that travelled with the
snapshot are used as they are; `available()` says whether they are present."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from many_bone_ik_b200._capi import RigDesc, rig_to_desc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libmbik_ref.so")
DOCTESTS = os.path.join(_HERE, "_ref", "ref_doctests")
BINDING_LIB = os.path.join(_HERE, "_ref", "libmbik_ref_binding.so")
REFERENCE = os.environ.get("MBIK_REFERENCE_DIR", "/root/reference")


def source_present():
    return os.path.exists(os.path.join(REFERENCE, "src", "many_bone_ik_3d.cpp"))


def build(force=False):
    """(Re)builds oracle/_ref when the reference sources are present; otherwise keeps whatever is prebuilt."""
    if source_present():
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j8", "ref", f"REFERENCE={REFERENCE}"] + (["-B"] if force else []))
    return LIB if os.path.exists(LIB) else None


def build_binding(force=False):
    """oracle/_ref/libmbik_ref_binding.so: the reference's classes with _process_modification replaced by the libmbik.so
    binding (many_bone_ik_b200/host/godot_module_binding.h).  Needs many_bone_ik_b200/libmbik.so built first."""
    if source_present():
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j8", "ref_binding", f"REFERENCE={REFERENCE}"] + (["-B"] if force else []))
    return BINDING_LIB if os.path.exists(BINDING_LIB) else None


def available():
    return os.path.exists(LIB) or (source_present() and build() is not None)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmbik_ref.so is not built and the reference sources are not present")
        L = C.CDLL(LIB)
        vp = C.c_void_p
        L.ref_solve_batch.argtypes = [C.POINTER(RigDesc), C.c_size_t, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_uint]
        L.ref_solve_frames.argtypes = [C.POINTER(RigDesc), C.c_size_t, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        L.ref_rig_facts.argtypes = [C.POINTER(RigDesc), vp, C.c_int32, vp, vp, vp]
        L.ref_step_weights.argtypes = [C.POINTER(RigDesc), C.c_int32, vp, C.c_int32]
        L.ref_cone_geometry.argtypes = [C.POINTER(RigDesc), vp, C.c_int32]
        L.ref_qcp_weighted_superpose.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
        L.ref_kusudama_point_in_limits.argtypes = [vp, C.c_int, vp, vp]
        L.ref_clamp_to_cos_half_angle.argtypes = [vp, C.c_double, vp]
        L.ref_swing_twist_y.argtypes = [vp, vp]
        L.ref_hardware_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def hardware_threads():
    return int(lib().ref_hardware_threads())


def solve_batch(rig, targets, start_pose=None, iterations=-1, threads=1, rebuild_each=False, want_local=False):
    """Returns (out_pose [n, n_bones, 10], [out_local [n, n_bones, 12]], status [n])."""
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
    rc = L.ref_solve_batch(C.byref(desc), n, _p(targets), _p(start_pose), _p(out), _p(loc), _p(st), int(iterations), int(threads),
                           1 if rebuild_each else 0)
    assert rc == 0
    return (out, loc, st) if want_local else (out, st)


def solve_frames(rig, targets, start_pose=None, iterations=-1, threads=1):
    """One long-lived node per pose over a frame sequence: frame 0 starts from start_pose (or rest), every later frame
    from what the skeleton holds after the previous write-back (the module's own modification_processed re-seed).
    targets [frames, n, n_pins, 12] -> dict(out [frames, n, n_bones, 10], local (raw IK locals), skeleton
    (get_bone_pose() after the frame) [frames, n, n_bones, 12], status [frames, n])."""
    L = lib()
    desc, keep = rig_to_desc(rig)
    targets = np.ascontiguousarray(targets, np.float32)
    f, n = targets.shape[:2]
    assert targets.shape == (f, n, rig.n_pins, 12)
    if start_pose is not None:
        start_pose = np.ascontiguousarray(start_pose, np.float32)
        assert start_pose.shape == (n, rig.n_bones, 12)
    out = np.zeros((f, n, rig.n_bones, 10), np.float32)
    loc = np.zeros((f, n, rig.n_bones, 12), np.float32)
    skel = np.zeros((f, n, rig.n_bones, 12), np.float32)
    st = np.zeros((f, n), np.uint32)
    rc = L.ref_solve_frames(C.byref(desc), n, f, _p(targets), _p(start_pose), _p(out), _p(loc), _p(skel), _p(st), int(iterations), int(threads))
    assert rc == 0
    return dict(out=out, local=loc, skeleton=skel, status=st)


def rig_facts(rig):
    L = lib()
    desc, keep = rig_to_desc(rig)
    cap = rig.n_bones
    order = np.full(cap, -1, np.int32)
    nseg = np.zeros(1, np.int32)
    dirb = np.zeros((cap, 9), np.float32)
    twb = np.zeros((cap, 9), np.float32)
    n = L.ref_rig_facts(C.byref(desc), _p(order), cap, _p(nseg), _p(dirb), _p(twb))
    return dict(bone_order=order[:n].copy(), n_segments=int(nseg[0]), dir_basis=dirb[:n].copy(), twist_basis=twb[:n].copy())


def step_weights(rig, step):
    L = lib()
    desc, keep = rig_to_desc(rig)
    w = np.zeros(256, np.float64)
    n = L.ref_step_weights(C.byref(desc), int(step), _p(w), 256)
    return w[:max(n, 0)].copy()


def cone_geometry(rig):
    L = lib()
    desc, keep = rig_to_desc(rig)
    nc = sum(len(c["cones"]) for c in rig.constraints)
    out = np.zeros((max(nc, 1), 9), np.float32)
    n = L.ref_cone_geometry(C.byref(desc), _p(out), nc)
    return out[:n].copy()


def qcp_weighted_superpose(moved, target, weight, translate):
    L = lib()
    moved = np.ascontiguousarray(moved, np.float32)
    target = np.ascontiguousarray(target, np.float32)
    weight = np.ascontiguousarray(weight, np.float64)
    out = np.zeros(7, np.float32)
    L.ref_qcp_weighted_superpose(_p(moved), _p(target), _p(weight), moved.shape[0], int(bool(translate)), _p(out))
    return out[:4].copy(), out[4:].copy()


def kusudama_point_in_limits(cones, point):
    L = lib()
    cones = np.ascontiguousarray(cones, np.float32).reshape(-1, 4)
    point = np.ascontiguousarray(point, np.float32)
    out = np.zeros(4, np.float32)
    L.ref_kusudama_point_in_limits(_p(cones), cones.shape[0], _p(point), _p(out))
    return out[:3].copy(), float(out[3])


def clamp_to_cos_half_angle(q, cos_half):
    L = lib()
    q = np.ascontiguousarray(q, np.float32)
    out = np.zeros(4, np.float32)
    L.ref_clamp_to_cos_half_angle(_p(q), float(cos_half), _p(out))
    return out


def swing_twist_y(q):
    L = lib()
    q = np.ascontiguousarray(q, np.float32)
    out = np.zeros(8, np.float32)
    L.ref_swing_twist_y(_p(q), _p(out))
    return out[:4].copy(), out[4:].copy()


_blib = None


def binding_solve_batch(rig, targets, start_pose=None, iterations=-1, rebuild_each=False):
    """The same scene as solve_batch, but every frame is solved by the CUDA path through the module binding.
    Returns (rc, out_pose [n, n_bones, 10], status [n]); rc = 0 or the first mbik error the binding saw."""
    global _blib
    if _blib is None:
        if not os.path.exists(BINDING_LIB) and build_binding() is None:
            raise RuntimeError("oracle/_ref/libmbik_ref_binding.so is not built and the reference sources are not present")
        _blib = C.CDLL(BINDING_LIB)
        vp = C.c_void_p
        _blib.ref_binding_solve_batch.argtypes = [C.POINTER(RigDesc), C.c_size_t, vp, vp, vp, vp, C.c_int, C.c_uint]
    desc, keep = rig_to_desc(rig)
    targets = np.ascontiguousarray(targets, np.float32)
    n = targets.shape[0]
    if start_pose is not None:
        start_pose = np.ascontiguousarray(start_pose, np.float32)
    out = np.zeros((n, rig.n_bones, 10), np.float32)
    st = np.zeros(n, np.uint32)
    rc = _blib.ref_binding_solve_batch(C.byref(desc), n, _p(targets), _p(start_pose), _p(out), _p(st), int(iterations), 1 if rebuild_each else 0)
    return int(rc), out, st


class BindingCrowd:
    """A scene of long-lived reference nodes (each its own Skeleton3D + ManyBoneIK3D) whose _process_modification is the
    deferred binding mbik_godot::CrowdBinding: nodes sharing a rig are solved by ONE mbik_solve_batch per frame."""

    def __init__(self):
        global _blib
        if not os.path.exists(BINDING_LIB) and build_binding() is None:
            raise RuntimeError("oracle/_ref/libmbik_ref_binding.so is not built and the reference sources are not present")
        self.L = C.CDLL(BINDING_LIB)
        vp = C.c_void_p
        self.L.ref_crowd_create.restype = vp
        self.L.ref_crowd_destroy.argtypes = [vp]
        self.L.ref_crowd_add.argtypes = [vp, C.POINTER(RigDesc), C.c_int, vp, C.c_int]
        self.L.ref_crowd_frame.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int)]
        self.h = C.c_void_p(self.L.ref_crowd_create())
        self.parts = []  # (rig, n_nodes)

    def add(self, rig, n_nodes, start_pose=None, iterations=-1):
        desc, keep = rig_to_desc(rig)
        if start_pose is not None:
            start_pose = np.ascontiguousarray(start_pose, np.float32)
            assert start_pose.shape == (n_nodes, rig.n_bones, 12)
        rc = self.L.ref_crowd_add(self.h, C.byref(desc), int(n_nodes), _p(start_pose), int(iterations))
        assert rc > 0, rc
        self.parts.append((rig, int(n_nodes)))

    def frame(self, targets_per_part):
        """targets_per_part: one [n_nodes, n_pins, 12] array per add() call.  Returns (rc, [out_pose per part], launches)."""
        flat = np.concatenate([np.ascontiguousarray(t, np.float32).reshape(-1) for t in targets_per_part])
        total = sum(n * rig.n_bones * 10 for rig, n in self.parts)
        out = np.zeros(total, np.float32)
        launches = C.c_int(0)
        rc = self.L.ref_crowd_frame(self.h, _p(flat), _p(out), None, C.byref(launches))
        outs, at = [], 0
        for rig, n in self.parts:
            outs.append(out[at:at + n * rig.n_bones * 10].reshape(n, rig.n_bones, 10).copy())
            at += n * rig.n_bones * 10
        return int(rc), outs, int(launches.value)

    def close(self):
        if self.h:
            self.L.ref_crowd_destroy(self.h)
            self.h = None


def run_doctests():
    """The reference's own 15 doctest cases on its own code; returns (exit code, stdout)."""
    if not os.path.exists(DOCTESTS):
        build()
    r = subprocess.run([DOCTESTS], capture_output=True, text=True)
    return r.returncode, r.stdout
