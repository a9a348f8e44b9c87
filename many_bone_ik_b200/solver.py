"""Thin Python binding of the C ABI (include/mbik.h) -- what tests and bench.py call.

`BatchedIKRig` wraps an opaque `mbik_rig*`.  Host buffers are numpy arrays; device buffers are anything with a
`data_ptr()` (torch CUDA tensors) or raw integer device pointers.  No compute happens in Python, and there
is no fallback: every solve goes through libmbik.so's CUDA kernel or raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import (MBIK_IO_DEVICE, MBIK_IO_HOST, MBIK_LOCAL_RECOMPOSED, MBIK_OUT_SOLVED_ONLY, MBIK_SCHED_SEGMENT_PARALLEL, MBIK_SCHED_THROUGHPUT,
                    LimitSetsInfo, RigInfo, SolveParams)

# kernel mapping selectors of mbik_solve_params.flags (include/mbik.h); identical results, for tests and measurements
SCHED_FLAGS = {"auto": 0, "throughput": MBIK_SCHED_THROUGHPUT, "segment_parallel": MBIK_SCHED_SEGMENT_PARALLEL}


class MbikError(RuntimeError):
    def __init__(self, code, where):
        lib = _capi.load_library()
        self.code = code
        detail = lib.mbik_last_error().decode(errors="replace")
        super().__init__(f"{where}: {lib.mbik_strerror(code).decode()} ({code}): {detail}")


def device_count():
    return int(_capi.load_library().mbik_device_count())


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))


class BatchedIKRig:
    """One rig = one reference `ManyBoneIK3D` configuration after `_bone_list_changed()`."""

    def __init__(self, rig):
        self.lib = _capi.load_library()
        self.rig = rig
        desc, keep = _capi.rig_to_desc(rig)
        h = C.c_void_p()
        rc = self.lib.mbik_rig_create(C.byref(desc), C.byref(h))
        if rc != 0:
            raise MbikError(rc, "mbik_rig_create")
        self.handle = h
        info = RigInfo()
        rc = self.lib.mbik_rig_get_info(self.handle, C.byref(info))
        if rc != 0:
            raise MbikError(rc, "mbik_rig_get_info")
        self.info = {f[0]: getattr(info, f[0]) for f in RigInfo._fields_}
        self.n_bones = rig.n_bones
        self.n_pins = rig.n_pins

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mbik_rig_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- schedule facts (host only; no GPU needed) -------------------------------------------------
    def bone_order(self):
        out = np.zeros(self.info["n_solved"], np.int32)
        self.lib.mbik_rig_get_bone_order(self.handle, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def schedule(self):
        """Segment-parallel schedule: rows (phase, warp, first step, end step, team size, member index)."""
        n = self.lib.mbik_rig_get_schedule(self.handle, None, 0)
        out = np.zeros((max(n, 1), 6), np.int32)
        self.lib.mbik_rig_get_schedule(self.handle, out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out[:n]

    def step_weights(self, step):
        buf = np.zeros(512, np.float64)
        n = self.lib.mbik_rig_get_step_weights(self.handle, int(step), buf.ctypes.data_as(C.POINTER(C.c_double)), 512)
        if n < 0:
            raise MbikError(n, "mbik_rig_get_step_weights")
        return buf[:n].copy()

    def bone_frames(self):
        n = self.info["n_solved"]
        d = np.zeros((n, 9), np.float32)
        t = np.zeros((n, 9), np.float32)
        fp = C.POINTER(C.c_float)
        self.lib.mbik_rig_get_bone_frames(self.handle, d.ctypes.data_as(fp), t.ctypes.data_as(fp))
        return d, t

    def cone_geometry(self):
        nc = sum(len(c["cones"]) for c in self.rig.constraints)
        out = np.zeros((max(nc, 1), 9), np.float32)
        n = self.lib.mbik_rig_get_cone_geometry(self.handle, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out[:max(n, 0)].copy()

    # ---- the hot path -------------------------------------------------------------------------------
    def solve(self, targets, start_pose=None, iterations=-1, device=-1, want_local=False, devices=None, sched="auto", newton_iters=0,
              solved_only=False, recomposed_local=False):
        """Host-buffer solve.  targets [n, n_pins, 12] float32 -> out_pose [n, n_bones, 10]
        (+ out_local [n, n_bones, 12]) + status [n].  `devices`: list of ordinals -> sharded multi-GPU call.
        `sched`: kernel mapping, one of SCHED_FLAGS (same bits either way).  `newton_iters` > 0 leaves reference parity
        (mbik_solve_params::newton_iters).  `solved_only`: out_pose [n, n_solved, 10] in bone_list order
        (MBIK_OUT_SOLVED_ONLY).  `recomposed_local`: out_local as Skeleton3D::get_bone_pose() returns it (MBIK_LOCAL_RECOMPOSED)."""
        targets = np.ascontiguousarray(targets, np.float32)
        n = targets.shape[0]
        if targets.shape != (n, self.n_pins, 12):
            raise ValueError(f"targets must be [n, {self.n_pins}, 12]")
        if start_pose is not None:
            start_pose = np.ascontiguousarray(start_pose, np.float32)
            if start_pose.shape != (n, self.n_bones, 12):
                raise ValueError(f"start_pose must be [n, {self.n_bones}, 12]")
        out = np.empty((n, self.info["n_solved"] if solved_only else self.n_bones, 10), np.float32)
        loc = np.empty((n, self.n_bones, 12), np.float32) if want_local else None
        st = np.zeros(n, np.uint32)
        flags = MBIK_IO_HOST | SCHED_FLAGS[sched] | (MBIK_OUT_SOLVED_ONLY if solved_only else 0) | (MBIK_LOCAL_RECOMPOSED if recomposed_local else 0)
        p = SolveParams(int(iterations), int(device), flags, None, int(newton_iters), 0)
        if devices is None:
            rc = self.lib.mbik_solve_batch(self.handle, C.byref(p), n, _ptr(targets), _ptr(start_pose), _ptr(out), _ptr(loc), _ptr(st))
        else:
            devs = np.ascontiguousarray(devices, np.int32)
            rc = self.lib.mbik_solve_batch_multi(self.handle, C.byref(p), n, _ptr(targets), _ptr(start_pose), _ptr(out), _ptr(loc), _ptr(st),
                                                 devs.ctypes.data_as(C.POINTER(C.c_int32)), int(devs.shape[0]))
        if rc != 0:
            raise MbikError(rc, "mbik_solve_batch")
        return (out, loc, st) if want_local else (out, st)

    # ---- per-pose limit sets -----------------------------------------------------
    def create_limit_sets(self, constraint_sets, asynchronous=False):
        """constraint_sets: [n_sets] lists of constraint dicts with the rig's rows (same bones and cone counts, other
        values).  Returns an opaque handle for solve_with_limits(); free it with destroy_limit_sets().
        asynchronous=True returns while the host authoring still runs (mbik_limit_sets_create_async)."""
        cons, cones, rows, cones_per_set = _capi.constraints_to_arrays(constraint_sets)
        h = C.c_void_p()
        fn = self.lib.mbik_limit_sets_create_async if asynchronous else self.lib.mbik_limit_sets_create
        rc = fn(self.handle, len(constraint_sets), C.cast(cons, C.POINTER(_capi.ConstraintDesc)),
                C.cast(cones, C.POINTER(_capi.ConeDesc)), int(cones_per_set), C.byref(h))
        if rc != 0:
            raise MbikError(rc, "mbik_limit_sets_create")
        return h

    def limit_sets_info(self, sets):
        """Waits for the authoring; {n_sets, bytes_per_set, table_bytes, author_seconds, author_threads}."""
        info = LimitSetsInfo()
        rc = self.lib.mbik_limit_sets_get_info(sets, C.byref(info))
        if rc != 0:
            raise MbikError(rc, "mbik_limit_sets_get_info")
        return {f[0]: getattr(info, f[0]) for f in LimitSetsInfo._fields_}

    def limit_set_geometry(self, sets, set_index):
        """(cones [n_cones, 9], twist_basis [n_solved, 9]) of one authored set: cone_geometry() / bone_frames()[1] layouts."""
        fp = C.POINTER(C.c_float)
        n = self.lib.mbik_limit_sets_get_geometry(sets, int(set_index), None, None)
        if n < 0:
            raise MbikError(n, "mbik_limit_sets_get_geometry")
        cones = np.zeros((n, 9), np.float32)
        twist = np.zeros((self.info["n_solved"], 9), np.float32)
        rc = self.lib.mbik_limit_sets_get_geometry(sets, int(set_index), cones.ctypes.data_as(fp), twist.ctypes.data_as(fp))
        if rc < 0:
            raise MbikError(rc, "mbik_limit_sets_get_geometry")
        return cones, twist

    def destroy_limit_sets(self, sets):
        self.lib.mbik_limit_sets_destroy(sets)

    def solve_with_limits(self, sets, set_index, targets, start_pose=None, iterations=-1, device=-1, want_local=False, sched="auto"):
        """Host-buffer solve where pose k uses limit set set_index[k]."""
        targets = np.ascontiguousarray(targets, np.float32)
        n = targets.shape[0]
        if targets.shape != (n, self.n_pins, 12):
            raise ValueError(f"targets must be [n, {self.n_pins}, 12]")
        set_index = np.ascontiguousarray(set_index, np.int32)
        if set_index.shape != (n,):
            raise ValueError("set_index must be [n]")
        if start_pose is not None:
            start_pose = np.ascontiguousarray(start_pose, np.float32)
            if start_pose.shape != (n, self.n_bones, 12):
                raise ValueError(f"start_pose must be [n, {self.n_bones}, 12]")
        out = np.empty((n, self.n_bones, 10), np.float32)
        loc = np.empty((n, self.n_bones, 12), np.float32) if want_local else None
        st = np.zeros(n, np.uint32)
        p = SolveParams(int(iterations), int(device), MBIK_IO_HOST | SCHED_FLAGS[sched], None, 0, 0)
        rc = self.lib.mbik_solve_batch_limits(self.handle, sets, C.byref(p), n, _ptr(set_index), _ptr(targets), _ptr(start_pose), _ptr(out), _ptr(loc),
                                              _ptr(st))
        if rc != 0:
            raise MbikError(rc, "mbik_solve_batch_limits")
        return (out, loc, st) if want_local else (out, st)

    def solve_with_limits_raw(self, sets, n_poses, set_index, targets, out_pose, start_pose=None, out_local=None, out_status=None, iterations=-1,
                              device=-1, flags=MBIK_IO_HOST, stream=None):
        p = SolveParams(int(iterations), int(device), int(flags), C.c_void_p(int(stream)) if stream else None, 0, 0)
        rc = self.lib.mbik_solve_batch_limits(self.handle, sets, C.byref(p), int(n_poses), _ptr(set_index), _ptr(targets), _ptr(start_pose),
                                              _ptr(out_pose), _ptr(out_local), _ptr(out_status))
        if rc != 0:
            raise MbikError(rc, "mbik_solve_batch_limits")

    def solve_raw(self, n_poses, targets, out_pose, start_pose=None, out_local=None, out_status=None, iterations=-1,
                  device=-1, flags=MBIK_IO_HOST, stream=None, newton_iters=0, devices=None):
        """Zero-copy call: pointers (numpy / torch tensors / ints) are passed straight through.
        With flags=MBIK_IO_DEVICE the launch is asynchronous on `stream`.  `devices`: mbik_solve_batch_multi (host buffers)."""
        p = SolveParams(int(iterations), int(device), int(flags), C.c_void_p(int(stream)) if stream else None, int(newton_iters), 0)
        if devices is not None:
            devs = np.ascontiguousarray(devices, np.int32)
            rc = self.lib.mbik_solve_batch_multi(self.handle, C.byref(p), int(n_poses), _ptr(targets), _ptr(start_pose), _ptr(out_pose),
                                                 _ptr(out_local), _ptr(out_status), devs.ctypes.data_as(C.POINTER(C.c_int32)), int(devs.shape[0]))
            if rc != 0:
                raise MbikError(rc, "mbik_solve_batch_multi")
            return
        rc = self.lib.mbik_solve_batch(self.handle, C.byref(p), int(n_poses), _ptr(targets), _ptr(start_pose), _ptr(out_pose),
                                       _ptr(out_local), _ptr(out_status))
        if rc != 0:
            raise MbikError(rc, "mbik_solve_batch")

    def last_kernel_ms(self, device=0):
        ms = C.c_float(0)
        rc = self.lib.mbik_last_kernel_ms(self.handle, int(device), C.byref(ms))
        if rc != 0:
            raise MbikError(rc, "mbik_last_kernel_ms")
        return float(ms.value)


class IKStream:
    """Warm-start streaming (include/mbik.h, mbik_stream_*): the poses of `n_poses` skeletons stay resident on the
    device; every frame uploads targets only and starts from the previous frame's solution, as the reference
    re-seeds its IK bones from the skeleton after each frame (src/many_bone_ik_3d.cpp:1084, :91-102)."""

    def __init__(self, rig: BatchedIKRig, n_poses, device=0, initial_pose=None, solved_only=False):
        self.lib = rig.lib
        self.rig = rig
        self.n = int(n_poses)
        self.out_rows = rig.info["n_solved"] if solved_only else rig.n_bones
        if initial_pose is not None:
            initial_pose = np.ascontiguousarray(initial_pose, np.float32)
            if initial_pose.shape != (self.n, rig.n_bones, 12):
                raise ValueError(f"initial_pose must be [{self.n}, {rig.n_bones}, 12]")
        h = C.c_void_p()
        rc = self.lib.mbik_stream_create_ex(rig.handle, int(device), self.n, _ptr(initial_pose), MBIK_OUT_SOLVED_ONLY if solved_only else 0, C.byref(h))
        if rc != 0:
            raise MbikError(rc, "mbik_stream_create")
        self.handle = h
        self._keep = []

    def submit(self, targets, out_pose=None, out_status=None, iterations=-1):
        """Asynchronous: buffers must stay alive until sync() (they are kept referenced here)."""
        if isinstance(targets, np.ndarray):
            targets = np.ascontiguousarray(targets, np.float32)
            if targets.shape != (self.n, self.rig.n_pins, 12):
                raise ValueError(f"targets must be [{self.n}, {self.rig.n_pins}, 12]")
        self._keep.append((targets, out_pose, out_status))
        rc = self.lib.mbik_stream_submit(self.handle, _ptr(targets), _ptr(out_pose), _ptr(out_status), int(iterations))
        if rc != 0:
            raise MbikError(rc, "mbik_stream_submit")

    def sync(self):
        rc = self.lib.mbik_stream_sync(self.handle)
        self._keep.clear()
        if rc != 0:
            raise MbikError(rc, "mbik_stream_sync")

    def read_local(self):
        out = np.empty((self.n, self.rig.n_bones, 12), np.float32)
        rc = self.lib.mbik_stream_read_local(self.handle, _ptr(out))
        self._keep.clear()
        if rc != 0:
            raise MbikError(rc, "mbik_stream_read_local")
        return out

    def reset(self, initial_pose=None):
        if initial_pose is not None:
            initial_pose = np.ascontiguousarray(initial_pose, np.float32)
        rc = self.lib.mbik_stream_reset(self.handle, _ptr(initial_pose))
        if rc != 0:
            raise MbikError(rc, "mbik_stream_reset")

    @property
    def frames(self):
        return int(self.lib.mbik_stream_frames(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mbik_stream_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
