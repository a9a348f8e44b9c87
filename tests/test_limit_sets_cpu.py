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


def test_limit_sets_rejected_on_stabilised_rigs():
    import rig_cases
    rig = rig_cases.humanoid_stabilized()
    R = BatchedIKRig(rig)
    with pytest.raises(MbikError):
        R.create_limit_sets([rig.constraints])


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
