"""Per-pose limit sets, host side: creation, validation, error behaviour (no GPU)."""
import copy

import numpy as np
import pytest

import limit_set_cases as LS
from many_bone_ik_b200 import BatchedIKRig, MbikError, device_count, rigs


def test_limit_sets_create_and_destroy():
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    sets = R.create_limit_sets(LS.variants(rig, 3))
    assert sets.value
    R.destroy_limit_sets(sets)


def test_limit_sets_must_keep_the_rigs_rows():
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    bad = copy.deepcopy(rig.constraints)
    bad[0]["bone"] = bad[1]["bone"]  # other bone in row 0
    with pytest.raises(MbikError):
        R.create_limit_sets([rig.constraints, bad])
    bad = copy.deepcopy(rig.constraints)
    bad[3]["cones"] = bad[3]["cones"] + [(0.0, 1.0, 0.0, 0.3)]  # other cone count
    bad[4]["cones"] = bad[4]["cones"][:-1] if len(bad[4]["cones"]) > 1 else bad[4]["cones"]
    if sum(len(c["cones"]) for c in bad) == sum(len(c["cones"]) for c in rig.constraints):
        with pytest.raises(MbikError):
            R.create_limit_sets([rig.constraints, bad])


def test_limit_sets_accepted_on_stabilised_rigs():
    import rig_cases
    rig = rig_cases.humanoid_stabilized()
    R = BatchedIKRig(rig)
    h = R.create_limit_sets([rig.constraints])
    assert R.limit_sets_info(h)["n_sets"] == 1
    R.destroy_limit_sets(h)


def test_limit_sets_parallel_and_asynchronous_authoring_build_the_same_table(monkeypatch):
    """The host authoring runs on a thread pool; one thread, all threads and the asynchronous entry point must describe
    the same table (sizes here; the GPU tests compare the solves), and an invalid set is reported by the wait."""
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 24)
    monkeypatch.setenv("MBIK_AUTHOR_THREADS", "1")
    h1 = R.create_limit_sets(sets)
    i1 = R.limit_sets_info(h1)
    monkeypatch.delenv("MBIK_AUTHOR_THREADS")
    h2 = R.create_limit_sets(sets, asynchronous=True)
    i2 = R.limit_sets_info(h2)  # waits
    assert i1["author_threads"] == 1 and i2["author_threads"] >= 1
    assert i1["n_sets"] == i2["n_sets"] == 24 and i1["table_bytes"] == i2["table_bytes"] == 24 * i1["bytes_per_set"]
    assert i1["bytes_per_set"] == R.info["n_cones"] * 160 + R.info["n_solved"] * 208
    R.destroy_limit_sets(h1)
    R.destroy_limit_sets(h2)
    bad = copy.deepcopy(rig.constraints)
    bad[0]["bone"] = bad[1]["bone"]
    h3 = R.create_limit_sets([rig.constraints, bad], asynchronous=True)  # accepted: the rows are checked by the workers
    with pytest.raises(MbikError):
        R.limit_sets_info(h3)
    R.destroy_limit_sets(h3)


@pytest.mark.parametrize("name", ["humanoid22", "quad80", "chain64", "humanoid_stabilized", "scaled_bones", "chain300"])
def test_authored_set_geometry_equals_a_rig_flattened_with_that_set(name):
    """A set is authored by the limit-dependent part of the flattener alone (author_constraints on the rig's topology), not
    by flattening the rig again: its cone / tangent-circle triples and twist frames must be the bits that a rig created with
    the set's constraint values holds -- and those are pinned to the reference's own geometry by test_reference_cpu /
    test_host_cpu.  Same through the thread pool and the asynchronous entry point."""
    import rig_cases
    rig = (rigs.RIGS.get(name) or getattr(rig_cases, name))()
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 5, seed=11)
    handles = [R.create_limit_sets(sets), R.create_limit_sets(sets, asynchronous=True)]
    for s_i, cons in enumerate(sets):
        Rs = BatchedIKRig(LS.rig_with(rig, cons))
        want_cones, want_twist = Rs.cone_geometry(), Rs.bone_frames()[1]
        for h in handles:
            cones, twist = R.limit_set_geometry(h, s_i)
            assert cones.shape == want_cones.shape and np.array_equal(cones.view(np.uint32), want_cones.view(np.uint32))
            assert np.array_equal(twist.view(np.uint32), want_twist.view(np.uint32))
    if len(sets) > 1 and R.info["n_cones"] > 0:
        assert not np.array_equal(R.limit_set_geometry(handles[0], 0)[0], R.limit_set_geometry(handles[0], 1)[0])
    with pytest.raises(MbikError):
        R.limit_set_geometry(handles[0], len(sets))
    for h in handles:
        R.destroy_limit_sets(h)


def test_limit_set_solve_without_gpu_fails_loudly():
    if device_count() > 0:
        pytest.skip("a CUDA device is present")
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    sets = R.create_limit_sets(LS.variants(rig, 2))
    T = rigs.random_targets(rig, 0, 4)
    with pytest.raises(MbikError) as e:
        R.solve_with_limits(sets, np.zeros(4, np.int32), T)
    assert e.value.code == -4  # MBIK_ERR_NO_DEVICE: there is no CPU fallback
    R.destroy_limit_sets(sets)
