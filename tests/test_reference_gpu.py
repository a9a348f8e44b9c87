"""GPU parity against the REFERENCE MODULE'S OWN CODE (not the restatement).

tests/golden/reference_solves.npz holds outputs of oracle/_ref/libmbik_ref.so -- /root/reference/src compiled
unmodified over the engine stand-in (tests/golden/make_reference_golden.py).  The CUDA path, through the C ABI, must
reproduce them bit for bit (NaN == NaN) in both kernel mappings.  Where the prebuilt reference library travelled to
the GPU box (oracle/_ref is git-ignored but not gpurun-ignored) it is also called live on larger batches."""
import os
import sys

import numpy as np
import pytest

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, rigs
from oracle import reference_py as Rf
from test_stages_gpu import gpu_clamp, gpu_point_in_limits, gpu_qcp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_reference_golden as MRG  # noqa: E402

CASES = MRG.all_cases()
SCHEDS = ("throughput", "segment_parallel", "auto")
needs_ref = pytest.mark.skipif(not os.path.exists(Rf.LIB), reason="prebuilt oracle/_ref/libmbik_ref.so did not travel to this box")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "reference_solves.npz"))


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_reference_fixtures(name, gold):
    rig = CASES[name]()
    R = BatchedIKRig(rig)
    for sched in SCHEDS:
        out, loc, st = R.solve(gold[name + "_targets"], want_local=True, sched=sched)
        assert _same(loc, gold[name + "_local"]), sched
        assert _same(out, gold[name + "_out"]), sched
        assert np.array_equal(st, gold[name + "_status"]), sched


@pytest.mark.parametrize("name", ["humanoid22", "quad80"])
def test_cuda_warm_start_reproduces_reference_fixtures(name, gold):
    R = BatchedIKRig(rigs.RIGS[name]())
    for sched in SCHEDS:
        out, loc, st = R.solve(gold[name + "_warm_targets"], start_pose=gold[name + "_local"], want_local=True, sched=sched)
        assert _same(loc, gold[name + "_warm_local"]) and _same(out, gold[name + "_warm_out"]) and np.array_equal(st, gold[name + "_warm_status"]), sched


def test_cuda_stages_reproduce_reference_fixtures(gold):
    qcp, kus, quats, cos_half = MRG.stage_inputs()
    got = np.stack([np.concatenate(gpu_qcp(m, t, w, tr)) for m, t, w, tr in qcp])
    assert _same(got, gold["stage_qcp"])
    assert _same(gpu_clamp(quats, cos_half), gold["stage_clamp"])
    for (cones, point), want in zip(kus, gold["stage_kusudama"]):
        p, ib = gpu_point_in_limits(cones, point[None])
        assert _same(np.concatenate([p[0], ib]).astype(np.float32), want)


@needs_ref
@pytest.mark.parametrize("name,n", [("humanoid22", 4096), ("chain64", 96), ("quad80", 256), ("star_mixed_pins", 512), ("humanoid_stabilized", 512),
                                    ("big_tree120", 128)])
def test_cuda_equals_live_reference(name, n):
    """BASELINE configs[1] (the 4096-pose humanoid batch) and the other rigs, CUDA vs the reference's own code."""
    rig = CASES[name]()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 7000, n)
    ref = Rf.solve_batch(rig, T, want_local=True, threads=Rf.hardware_threads())
    for sched in SCHEDS:
        out, loc, st = R.solve(T, want_local=True, sched=sched)
        assert _same(loc, ref[1]), sched
        assert _same(out, ref[0]), sched
        assert np.array_equal(st, ref[2]), sched


@needs_ref
@pytest.mark.parametrize("name", sorted(rig_cases.LARGE_RIGS))
def test_cuda_equals_live_reference_beyond_128_solved_bones(name):
    """The {256, 256, 32} kernel variant vs the reference module's own code (which has no bone limit)."""
    rig = rig_cases.LARGE_RIGS[name]()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 7000, 32)
    ref = Rf.solve_batch(rig, T, want_local=True, threads=Rf.hardware_threads())
    out, loc, st = R.solve(T, want_local=True)
    assert _same(loc, ref[1]) and _same(out, ref[0]) and np.array_equal(st, ref[2])


@needs_ref
def test_cuda_frame_sequence_equals_live_reference():
    """Five warm-started frames (each frame starts from the previous frame's solved locals, the reference's
    frame-to-frame seeding, src/many_bone_ik_3d.cpp:1084): CUDA and the reference's own code stay bit-identical."""
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 64
    start_g = start_r = None
    for frame in range(5):
        T = rigs.random_targets(rig, 100 * frame, n)
        og, lg, sg = R.solve(T, start_pose=start_g, want_local=True)
        orf, lrf, srf = Rf.solve_batch(rig, T, start_pose=start_r, want_local=True, threads=4)
        assert _same(lg, lrf) and _same(og, orf) and np.array_equal(sg, srf), frame
        start_g, start_r = lg, lrf


@pytest.mark.parametrize("name", ["humanoid22", "quad80", "chain_diverging"])
def test_stream_reproduces_reference_frame_sequences(name, gold):
    """mbik_stream_* against the committed fixtures of the reference module's own node living across frames
    (ref_solve_frames: nothing re-seeds the skeleton from outside; chain_diverging goes non-finite and is reset)."""
    from many_bone_ik_b200 import IKStream
    rig = CASES[name]()
    R = BatchedIKRig(rig)
    Tf = gold[name + "_frames_targets"]
    frames, n = Tf.shape[:2]
    S = IKStream(R, n, device=0)
    outs = [np.empty((n, rig.n_bones, 10), np.float32) for _ in range(frames)]
    sts = [np.empty(n, np.uint32) for _ in range(frames)]
    for f in range(frames):
        S.submit(np.ascontiguousarray(Tf[f]), outs[f], sts[f])
    S.sync()
    for f in range(frames):
        assert _same(outs[f], gold[name + "_frames_out"][f]), f
        assert np.array_equal(sts[f], gold[name + "_frames_status"][f]), f
    assert _same(S.read_local(), gold[name + "_frames_skeleton"][-1])


@needs_ref
def test_stream_equals_live_reference_node_across_frames():
    """300 long-lived reference nodes over 5 frames vs one device-resident stream (solved-only output layout)."""
    from many_bone_ik_b200 import IKStream
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n, frames = 300, 5
    Tf = np.stack([rigs.random_targets(rig, 31 * f, n) for f in range(frames)])
    sp = rig_cases.perturbed_start_pose(rig, n, seed=5)
    ref = Rf.solve_frames(rig, Tf, start_pose=sp, threads=8)
    order = R.bone_order()
    S = IKStream(R, n, device=0, initial_pose=sp, solved_only=True)
    outs = [np.empty((n, len(order), 10), np.float32) for _ in range(frames)]
    for f in range(frames):
        S.submit(np.ascontiguousarray(Tf[f]), outs[f], None)
    S.sync()
    for f in range(frames):
        assert _same(outs[f], ref["out"][f][:, order]), f
    assert _same(S.read_local(), ref["skeleton"][-1])


needs_binding = pytest.mark.skipif(not os.path.exists(Rf.BINDING_LIB), reason="prebuilt oracle/_ref/libmbik_ref_binding.so did not travel to this box")


@needs_ref
@needs_binding
@pytest.mark.parametrize("name", ["humanoid22", "chain64", "quad80", "star_mixed_pins", "two_roots", "humanoid_stabilized", "humanoid_constraint_mode",
                                  "no_pins", "scaled_bones", "random_rig_7", "random_rig_19"])
def test_module_binding_drop_in_equals_the_reference(name):
    """The drop-in claim at the reference's own class boundary.  Two identical headless scenes built from the reference's
    own classes and configured through its property paths; in one, ManyBoneIK3D::_process_modification is the reference's
    CPU solver, in the other it is the binding a maintainer would add (many_bone_ik_b200/host/godot_module_binding.h ->
    mbik_rig_create / mbik_solve_batch, one pose per frame).  The Skeleton3D must receive bit-identical position /
    rotation / scale for every bone, for a fresh node per pose, a long-lived node over many frames, and perturbed poses."""
    rig = CASES[name]()
    n = 6
    T = rigs.random_targets(rig, 300, n)
    start = rig_cases.perturbed_start_pose(rig, n, seed=3)
    for sp in (None, start):
        for rebuild in (True, False):
            ref_out, ref_st = Rf.solve_batch(rig, T, start_pose=sp, rebuild_each=rebuild)
            rc, out, st = Rf.binding_solve_batch(rig, T, start_pose=sp, rebuild_each=rebuild)
            assert rc == 0
            assert _same(out, ref_out), (name, rebuild)


@needs_ref
@needs_binding
def test_crowd_binding_one_launch_per_rig_equals_the_reference_nodes():
    """mbik_godot::CrowdBinding (many_bone_ik_b200/host/godot_module_binding.h): 96 humanoid nodes and 70 quadruped nodes,
    each a reference ManyBoneIK3D with its own Skeleton3D living across 4 frames, enqueue their frames; one flush per
    frame solves every rig's nodes in ONE mbik_solve_batch (2 launches per frame instead of 166).  Every skeleton must
    receive, bit for bit, what the reference's own _process_modification leaves on it -- frame after frame, i.e. including
    the re-seed from the recomposed skeleton pose."""
    parts = [(rigs.humanoid22(), 96), (rigs.quad80(), 70)]
    frames = 4
    crowd = Rf.BindingCrowd()
    refs, Ts = [], []
    for k, (rig, n) in enumerate(parts):
        sp = rig_cases.perturbed_start_pose(rig, n, seed=40 + k)
        crowd.add(rig, n, start_pose=sp)
        Tf = np.stack([rigs.random_targets(rig, 7000 + 50 * f + k, n) for f in range(frames)])
        Ts.append(Tf)
        refs.append(Rf.solve_frames(rig, Tf, start_pose=sp, threads=8))
    for f in range(frames):
        rc, outs, launches = crowd.frame([T[f] for T in Ts])
        assert rc == 0 and launches == len(parts), (rc, launches)
        for k in range(len(parts)):
            assert _same(outs[k], refs[k]["out"][f]), (f, parts[k][0].name)
    crowd.close()
