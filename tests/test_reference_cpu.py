"""The reference module's OWN code as the pin (CPU, no GPU).

oracle/_ref/libmbik_ref.so = every translation unit under /root/reference/src compiled unmodified over the engine
stand-in oracle/godot_shim/ (make -C oracle ref).  Its outputs are frozen in tests/golden/reference_solves.npz
(tests/golden/make_reference_golden.py).  These tests check, bit for bit (NaN == NaN):
  * the restatement oracle (oracle/ewbik_oracle.cpp) against the committed reference fixtures -- always;
  * the product's host flattener (solve order, bone-direction / twist frames, cone geometry) against them -- always;
  * the oracle against the live reference library on more poses, both life-cycle modes, warm starts and random
    rigs -- when oracle/_ref is built (it is built here from /root/reference and travels to the GPU box prebuilt);
  * the reference's own 15 doctest cases on its own code -- when oracle/_ref is built;
  * that the committed fixtures are what the reference produces today -- when /root/reference is present."""
import os
import sys

import numpy as np
import pytest

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, rigs
from oracle import oracle_py as O
from oracle import reference_py as Rf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_reference_golden as MRG  # noqa: E402

CASES = MRG.all_cases()
needs_ref = pytest.mark.skipif(not Rf.available(), reason="oracle/_ref/libmbik_ref.so not built and /root/reference not present")
needs_src = pytest.mark.skipif(not Rf.source_present(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "reference_solves.npz"))


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@needs_ref
def test_reference_own_doctest_cases_pass_on_its_own_code():
    rc, text = Rf.run_doctests()
    assert rc == 0, text
    assert "15 cases, 52 checks, 0 failed" in text.splitlines()[-1]


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_reference_fixtures(name, gold):
    rig = CASES[name]()
    T = rigs.random_targets(rig, 0, 4)
    assert np.array_equal(T, gold[name + "_targets"]), "target generator changed: regenerate the fixtures"
    for rebuild in (True, False):
        out, loc, st = O.solve_batch(rig, T, want_local=True, rebuild_each=rebuild, threads=2)
        assert _same(loc, gold[name + "_local"]), (name, rebuild)
        assert _same(out, gold[name + "_out"]), (name, rebuild)
        assert np.array_equal(st, gold[name + "_status"]), (name, rebuild)
    facts = O.rig_facts(rig)
    assert np.array_equal(facts["bone_order"], gold[name + "_order"])
    assert _same(facts["dir_basis"], gold[name + "_dir"])
    assert _same(facts["twist_basis"], gold[name + "_twist"])
    assert _same(O.cone_geometry(rig), gold[name + "_cones"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_flattener_equals_reference_setup(name, gold):
    """mbik_rig_create (the product's host flattener, no GPU needed) against what the reference's own
    _bone_list_changed built: solve order, bone-direction and twist-axes frames, cone / tangent-circle geometry."""
    R = BatchedIKRig(CASES[name]())
    assert np.array_equal(R.bone_order(), gold[name + "_order"])
    d, t = R.bone_frames()
    assert _same(d, gold[name + "_dir"])
    assert _same(t, gold[name + "_twist"])
    geo = R.cone_geometry()
    assert _same(geo, gold[name + "_cones"][:geo.shape[0]]) and geo.shape[0] == gold[name + "_cones"].shape[0]


@pytest.mark.parametrize("name", ["humanoid22", "quad80"])
def test_oracle_warm_start_equals_reference_fixtures(name, gold):
    rig = rigs.RIGS[name]()
    out, loc, st = O.solve_batch(rig, gold[name + "_warm_targets"], start_pose=gold[name + "_local"], want_local=True)
    assert _same(loc, gold[name + "_warm_local"]) and _same(out, gold[name + "_warm_out"]) and np.array_equal(st, gold[name + "_warm_status"])


def test_oracle_stages_equal_reference_fixtures(gold):
    qcp, kus, quats, cos_half = MRG.stage_inputs()
    got = np.stack([np.concatenate(O.qcp_weighted_superpose(m, t, w, tr)) for m, t, w, tr in qcp])
    assert _same(got, gold["stage_qcp"])
    got = np.stack([np.concatenate([p, [ib]]).astype(np.float32) for p, ib in (O.kusudama_point_in_limits(c, pt) for c, pt in kus)])
    assert _same(got, gold["stage_kusudama"])
    assert _same(np.stack([O.clamp_to_cos_half_angle(q, c) for q, c in zip(quats, cos_half)]), gold["stage_clamp"])
    assert _same(np.stack([np.concatenate(O.swing_twist_y(q)) for q in quats]), gold["stage_swing_twist"])


@needs_ref
@pytest.mark.parametrize("name", sorted(list(rigs.RIGS) + list(rig_cases.EDGE_RIGS)))
def test_oracle_equals_live_reference(name):
    """More poses than the fixtures hold, other seeds, the long-lived-node life cycle (one scene re-seeded per
    frame) and a perturbed start pose."""
    rig = CASES[name]()
    n = 12
    T = rigs.random_targets(rig, 5000, n)
    start = rig_cases.perturbed_start_pose(rig, n, seed=11)
    for sp in (None, start):
        a = O.solve_batch(rig, T, start_pose=sp, want_local=True)
        b = Rf.solve_batch(rig, T, start_pose=sp, want_local=True, threads=3)
        for x, y in zip(a, b):
            assert _same(x, y), name


@needs_ref
@pytest.mark.parametrize("seed", range(24, 64))
def test_oracle_equals_live_reference_random_rigs(seed):
    rig = rig_cases.random_rig(seed)
    T = rigs.random_targets(rig, 0, 5)
    a = O.solve_batch(rig, T, want_local=True, rebuild_each=True)
    b = Rf.solve_batch(rig, T, want_local=True, rebuild_each=True)
    for x, y in zip(a, b):
        assert _same(x, y)
    fo, fr = O.rig_facts(rig), Rf.rig_facts(rig)
    for k in fo:
        assert _same(fo[k], fr[k]), k
    for step in range(len(fo["bone_order"])):
        assert np.array_equal(O.step_weights(rig, step), Rf.step_weights(rig, step))


@needs_ref
def test_reference_iteration_override_and_thread_partition():
    rig = rigs.humanoid22()
    T = rigs.random_targets(rig, 0, 21)
    for it in (0, 1, 3):
        a = O.solve_batch(rig, T, iterations=it, want_local=True)
        b = Rf.solve_batch(rig, T, iterations=it, want_local=True, threads=4)
        for x, y in zip(a, b):
            assert _same(x, y), it


@needs_src
def test_committed_reference_fixtures_are_current(gold):
    """Regenerates a sample of the fixtures from the reference sources and compares with the committed file."""
    for name in ("humanoid22", "chain64", "quad80", "star_mixed_pins", "random_rig_3"):
        rig = CASES[name]()
        out, loc, st = Rf.solve_batch(rig, gold[name + "_targets"], want_local=True, rebuild_each=True)
        assert _same(loc, gold[name + "_local"]) and _same(out, gold[name + "_out"]) and np.array_equal(st, gold[name + "_status"])


needs_binding = pytest.mark.skipif(not (os.path.exists(Rf.BINDING_LIB) or Rf.source_present()),
                                   reason="oracle/_ref/libmbik_ref_binding.so not built and /root/reference not present")


@needs_binding
def test_module_binding_without_a_gpu_reports_and_leaves_the_pose_untouched():
    """many_bone_ik_b200/host/godot_module_binding.h compiled against the reference's own classes: with no CUDA device the
    binding's _process_modification logs MBIK_ERR_NO_DEVICE (there is no CPU fallback) and the skeleton keeps its pose,
    the behaviour of the reference's ERR_FAIL_* early-outs."""
    from many_bone_ik_b200 import device_count
    if device_count() > 0:
        pytest.skip("a CUDA device is present: the binding solves (covered by tests/test_reference_gpu.py)")
    rig = rigs.humanoid22()
    T = rigs.random_targets(rig, 0, 3)
    start = rig_cases.perturbed_start_pose(rig, 3, seed=5)
    rc, out, st = Rf.binding_solve_batch(rig, T, start_pose=start)
    assert rc == -4  # MBIK_ERR_NO_DEVICE (include/mbik.h)
    untouched, _ = Rf.solve_batch(rig, T, start_pose=start, iterations=0)
    assert np.array_equal(out, untouched)


@needs_ref
def test_reference_node_across_frames_equals_oracle_reseeded_from_the_recomposed_pose():
    """What the module's frame-to-frame seeding is: a long-lived ManyBoneIK3D re-seeds its IK bones from
    Skeleton3D::get_bone_pose(), which -- once IKBone3D::set_skeleton_bone_pose wrote position / rotation / scale -- is
    the recomposed Transform3D(Basis(rotation, scale), position), with the non-finite reset, NOT the raw IK transforms.
    ref_solve_frames (no outside re-seeding) == per-frame oracle solves started from oracle_py.recompose_pose()."""
    for name in ("humanoid22", "quad80", "chain_diverging", "scaled_bones"):
        rig = CASES[name]()
        n, frames = 8, 4
        Tf = np.stack([rigs.random_targets(rig, 900 + 10 * f, n) for f in range(frames)])
        ref = Rf.solve_frames(rig, Tf, threads=4)
        start = None
        for f in range(frames):
            o, loc, st = O.solve_batch(rig, Tf[f], start_pose=start, want_local=True, threads=4)
            assert _same(o, ref["out"][f]) and _same(loc, ref["local"][f]) and np.array_equal(st, ref["status"][f]), (name, f)
            start = O.recompose_pose(rig, o, start)
            assert _same(start, ref["skeleton"][f]), (name, f)
