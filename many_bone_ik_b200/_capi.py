"""ctypes mirror of include/mbik.h (structs + prototypes) and the loader for libmbik.so.

The library is built in-tree (many_bone_ik_b200/csrc/build.py -> many_bone_ik_b200/libmbik.so); there is
no Python or CPU fallback: if it is missing, importing the solver raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MBIK_LIB: an alternative build of the same library (kernel experiments, profiles/build_variant.py)
LIB_PATH = os.environ.get("MBIK_LIB") or os.path.join(_HERE, "libmbik.so")


class PinDesc(C.Structure):
    _fields_ = [("bone", C.c_int32), ("weight", C.c_float), ("motion_propagation_factor", C.c_float),
                ("direction_priorities", C.c_float * 3)]


class ConeDesc(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float)]


class ConstraintDesc(C.Structure):
    _fields_ = [("bone", C.c_int32), ("twist_from", C.c_float), ("twist_range", C.c_float),
                ("n_cones", C.c_int32), ("cone_offset", C.c_int32)]


class RigDesc(C.Structure):
    _fields_ = [("n_bones", C.c_int32), ("parent", C.POINTER(C.c_int32)), ("rest_local", C.POINTER(C.c_float)),
                ("n_pins", C.c_int32), ("pins", C.POINTER(PinDesc)),
                ("n_constraints", C.c_int32), ("constraints", C.POINTER(ConstraintDesc)),
                ("cones", C.POINTER(ConeDesc)),
                ("n_bone_damp", C.c_int32), ("bone_damp", C.POINTER(C.c_float)),
                ("default_damp", C.c_float), ("iterations_per_frame", C.c_int32),
                ("stabilization_passes", C.c_int32), ("constraint_mode", C.c_int32)]


class SolveParams(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32), ("stream", C.c_void_p),
                ("newton_iters", C.c_int32), ("reserved", C.c_int32)]


class LimitSetsInfo(C.Structure):
    _fields_ = [("n_sets", C.c_int32), ("bytes_per_set", C.c_uint32), ("table_bytes", C.c_int64), ("author_seconds", C.c_double),
                ("author_threads", C.c_int32)]


class RigInfo(C.Structure):
    _fields_ = [("n_bones", C.c_int32), ("n_solved", C.c_int32), ("n_segments", C.c_int32), ("n_steps", C.c_int32),
                ("n_effectors", C.c_int32), ("n_pins", C.c_int32), ("max_headings", C.c_int32), ("n_cones", C.c_int32),
                ("iterations", C.c_int32), ("kernel_capacity", C.c_int32), ("rig_blob_bytes", C.c_int64),
                ("flops_per_solve", C.c_double), ("max_segment_len", C.c_int32), ("max_walk_stack", C.c_int32),
                ("sp_roles", C.c_int32), ("sp_phases", C.c_int32), ("sp_gain", C.c_double)]


MBIK_IO_HOST = 0
MBIK_IO_DEVICE = 1
MBIK_SCHED_THROUGHPUT = 2
MBIK_SCHED_SEGMENT_PARALLEL = 4
MBIK_OUT_SOLVED_ONLY = 8
MBIK_LOCAL_RECOMPOSED = 16

# every symbol include/mbik.h declares (tests check the library exports exactly these)
EXPORTED_SYMBOLS = [
    "mbik_device_count", "mbik_strerror", "mbik_last_error", "mbik_rig_create", "mbik_rig_destroy",
    "mbik_rig_get_info", "mbik_rig_get_bone_order", "mbik_rig_get_schedule", "mbik_rig_get_step_weights", "mbik_rig_get_bone_frames",
    "mbik_rig_get_cone_geometry", "mbik_solve_batch", "mbik_solve_batch_multi", "mbik_alloc_pinned",
    "mbik_free_pinned", "mbik_last_kernel_ms", "mbik_measure_fp32_tflops", "mbik_selftest",
    "mbik_stream_create", "mbik_stream_destroy", "mbik_stream_submit", "mbik_stream_sync", "mbik_stream_read_local",
    "mbik_stream_reset", "mbik_stream_frames", "mbik_stage_qcp", "mbik_stage_clamp", "mbik_stage_point_in_limits",
    "mbik_limit_sets_create", "mbik_limit_sets_destroy", "mbik_solve_batch_limits",
    "mbik_limit_sets_create_async", "mbik_limit_sets_wait", "mbik_limit_sets_get_info", "mbik_limit_sets_get_geometry", "mbik_stream_create_ex", "mbik_stage_qcp_newton",
]


def constraints_to_arrays(constraint_sets):
    """[n_sets] lists of constraint dicts (same rows as the rig's) -> (ConstraintDesc[n_sets * rows], ConeDesc[n_sets * cones_per_set],
    rows, cones_per_set).  cone_offset is relative to the set's own block."""
    rows = len(constraint_sets[0])
    cones_per_set = sum(len(c["cones"]) for c in constraint_sets[0])
    cons = (ConstraintDesc * max(1, len(constraint_sets) * rows))()
    cones = (ConeDesc * max(1, len(constraint_sets) * cones_per_set))()
    for s, cs in enumerate(constraint_sets):
        if len(cs) != rows or sum(len(c["cones"]) for c in cs) != cones_per_set:
            raise ValueError("every limit set needs the rig's constraint rows and cone counts")
        k = 0
        for i, c in enumerate(cs):
            d = cons[s * rows + i]
            d.bone, d.twist_from, d.twist_range = int(c["bone"]), float(c["twist_from"]), float(c["twist_range"])
            d.n_cones, d.cone_offset = len(c["cones"]), k
            for (cx, cy, cz, r) in c["cones"]:
                cone = cones[s * cones_per_set + k]
                cone.center[0], cone.center[1], cone.center[2], cone.radius = float(cx), float(cy), float(cz), float(r)
                k += 1
    return cons, cones, rows, cones_per_set


def rig_to_desc(rig):
    """Build a RigDesc for `rig` (many_bone_ik_b200.rigs.Rig).  Returns (desc, keepalive)."""
    keep = []
    parent = np.ascontiguousarray(rig.parent, np.int32)
    rest = np.ascontiguousarray(rig.rest_local, np.float32).reshape(-1)
    pins = (PinDesc * max(1, len(rig.pins)))()
    for i, p in enumerate(rig.pins):
        pins[i].bone = int(p["bone"])
        pins[i].weight = float(p["weight"])
        pins[i].motion_propagation_factor = float(p["mpf"])
        for a in range(3):
            pins[i].direction_priorities[a] = float(p["priorities"][a])
    n_cones = sum(len(c["cones"]) for c in rig.constraints)
    cons = (ConstraintDesc * max(1, len(rig.constraints)))()
    cones = (ConeDesc * max(1, n_cones))()
    k = 0
    for i, c in enumerate(rig.constraints):
        cons[i].bone = int(c["bone"])
        cons[i].twist_from = float(c["twist_from"])
        cons[i].twist_range = float(c["twist_range"])
        cons[i].n_cones = len(c["cones"])
        cons[i].cone_offset = k
        for (cx, cy, cz, r) in c["cones"]:
            cones[k].center[0], cones[k].center[1], cones[k].center[2] = float(cx), float(cy), float(cz)
            cones[k].radius = float(r)
            k += 1
    damp = np.ascontiguousarray(rig.bone_damp, np.float32)
    keep += [parent, rest, pins, cons, cones, damp]
    d = RigDesc()
    d.n_bones = rig.n_bones
    d.parent = parent.ctypes.data_as(C.POINTER(C.c_int32))
    d.rest_local = rest.ctypes.data_as(C.POINTER(C.c_float))
    d.n_pins = len(rig.pins)
    d.pins = C.cast(pins, C.POINTER(PinDesc))
    d.n_constraints = len(rig.constraints)
    d.constraints = C.cast(cons, C.POINTER(ConstraintDesc))
    d.cones = C.cast(cones, C.POINTER(ConeDesc))
    d.n_bone_damp = int(damp.shape[0])
    d.bone_damp = damp.ctypes.data_as(C.POINTER(C.c_float))
    d.default_damp = float(rig.default_damp)
    d.iterations_per_frame = int(rig.iterations)
    d.stabilization_passes = int(rig.stabilization_passes)
    d.constraint_mode = int(bool(rig.constraint_mode))
    return d, keep


_lib = None


def load_library():
    """Load libmbik.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m many_bone_ik_b200.csrc.build` "
            "(or __graft_entry__.build()).  many_bone_ik_b200 has no CPU/Python fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, fp, dp, u32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    lib.mbik_device_count.restype = C.c_int
    lib.mbik_strerror.restype = C.c_char_p
    lib.mbik_strerror.argtypes = [C.c_int]
    lib.mbik_last_error.restype = C.c_char_p
    lib.mbik_rig_create.argtypes = [C.POINTER(RigDesc), C.POINTER(vp)]
    lib.mbik_rig_destroy.argtypes = [vp]
    lib.mbik_rig_get_info.argtypes = [vp, C.POINTER(RigInfo)]
    lib.mbik_rig_get_bone_order.argtypes = [vp, i32p]
    lib.mbik_rig_get_schedule.argtypes = [vp, i32p, C.c_int32]
    lib.mbik_rig_get_step_weights.argtypes = [vp, C.c_int32, dp, C.c_int32]
    lib.mbik_rig_get_bone_frames.argtypes = [vp, fp, fp]
    lib.mbik_rig_get_cone_geometry.argtypes = [vp, fp]
    lib.mbik_solve_batch.argtypes = [vp, C.POINTER(SolveParams), C.c_size_t, vp, vp, vp, vp, vp]
    lib.mbik_solve_batch_multi.argtypes = [vp, C.POINTER(SolveParams), C.c_size_t, vp, vp, vp, vp, vp, i32p, C.c_int32]
    lib.mbik_alloc_pinned.restype = vp
    lib.mbik_alloc_pinned.argtypes = [C.c_size_t]
    lib.mbik_free_pinned.argtypes = [vp]
    lib.mbik_last_kernel_ms.argtypes = [vp, C.c_int32, fp]
    lib.mbik_measure_fp32_tflops.argtypes = [C.c_int32, C.c_int32, dp]
    lib.mbik_selftest.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.mbik_stream_create.argtypes = [vp, C.c_int32, C.c_size_t, vp, C.POINTER(vp)]
    lib.mbik_stream_destroy.argtypes = [vp]
    lib.mbik_stream_submit.argtypes = [vp, vp, vp, vp, C.c_int32]
    lib.mbik_stream_sync.argtypes = [vp]
    lib.mbik_stream_read_local.argtypes = [vp, vp]
    lib.mbik_stream_reset.argtypes = [vp, vp]
    lib.mbik_stream_frames.argtypes = [vp]
    lib.mbik_stream_frames.restype = C.c_int64
    lib.mbik_stage_qcp.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, vp]
    lib.mbik_stage_clamp.argtypes = [C.c_int32, C.c_int32, vp, vp, vp]
    lib.mbik_stage_point_in_limits.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp]
    lib.mbik_limit_sets_create.argtypes = [vp, C.c_int32, C.POINTER(ConstraintDesc), C.POINTER(ConeDesc), C.c_int32, C.POINTER(vp)]
    lib.mbik_limit_sets_destroy.argtypes = [vp]
    lib.mbik_limit_sets_create_async.argtypes = lib.mbik_limit_sets_create.argtypes
    lib.mbik_limit_sets_wait.argtypes = [vp]
    lib.mbik_limit_sets_get_info.argtypes = [vp, C.POINTER(LimitSetsInfo)]
    lib.mbik_limit_sets_get_geometry.argtypes = [vp, C.c_int32, fp, fp]
    lib.mbik_stream_create_ex.argtypes = [vp, C.c_int32, C.c_size_t, vp, C.c_uint32, C.POINTER(vp)]
    lib.mbik_stage_qcp_newton.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, C.c_int32, C.c_int32, vp]
    lib.mbik_solve_batch_limits.argtypes = [vp, vp, C.POINTER(SolveParams), C.c_size_t, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib
