"""Per-pose limit sets on the GPU: pose k solved with limit set set_index[k] must equal, bit for bit, the same pose solved
by the CPU oracle (and by the reference module's own code where oracle/_ref travelled) on a rig that carries that set's
constraint values -- i.e. one batched launch == K differently limited characters."""
import os

import numpy as np
import pytest

import limit_set_cases as LS
import rig_cases
from many_bone_ik_b200 import BatchedIKRig, rigs
from many_bone_ik_b200._capi import MBIK_IO_DEVICE
from oracle import oracle_py as O
from oracle import reference_py as Rf

pytestmark = pytest.mark.gpu

CASES = dict(rigs.RIGS)
CASES.update({k: rig_cases.EDGE_RIGS[k] for k in ("star_mixed_pins", "two_roots", "chain_multibone_root", "humanoid_constraint_mode", "big_tree120")})


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def _expected(rig, sets, idx, T, start, solver):
    out = loc = st = None
    for s, cons in enumerate(sets):
        o, l, t = solver(LS.rig_with(rig, cons), T, start_pose=start, want_local=True)
        if out is None:
            out, loc, st = np.empty_like(o), np.empty_like(l), np.empty_like(t)
        m = idx == s
        out[m], loc[m], st[m] = o[m], l[m], t[m]
    return out, loc, st


@pytest.mark.parametrize("name", sorted(CASES))
def test_limit_sets_equal_per_set_rigs(name):
    rig = CASES[name]()
    n = 700 if name in ("humanoid22", "star_mixed_pins") else 96
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 4)
    h = R.create_limit_sets(sets)
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 4, n).astype(np.int32)
    T = rigs.random_targets(rig, 40, n)
    for start in (None, rig_cases.perturbed_start_pose(rig, n, seed=9)):
        want = _expected(rig, sets, idx, T, start, lambda r, t, **kw: O.solve_batch(r, t, threads=8, **kw))
        for sched in ("auto", "throughput", "segment_parallel"):  # both kernel mappings read the pose's limit-set record
            got = R.solve_with_limits(h, idx, T, start_pose=start, want_local=True, sched=sched)
            assert _same(got[1], want[1]) and _same(got[0], want[0]) and np.array_equal(got[2], want[2]), sched
    R.destroy_limit_sets(h)


def test_limit_sets_with_stabilisation_passes():
    """stabilization_passes > 0 and per-pose limit sets in one launch (the STAB x LIMS instantiation)."""
    rig = rig_cases.humanoid_stabilized()
    n = 400
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 3, seed=21)
    h = R.create_limit_sets(sets, asynchronous=True)  # the solve waits for the authoring
    idx = (np.arange(n) % 3).astype(np.int32)
    T = rigs.random_targets(rig, 77, n)
    want = _expected(rig, sets, idx, T, None, lambda r, t, **kw: O.solve_batch(r, t, threads=8, **kw))
    got = R.solve_with_limits(h, idx, T, want_local=True)
    assert _same(got[1], want[1]) and _same(got[0], want[0]) and np.array_equal(got[2], want[2])
    R.destroy_limit_sets(h)


@pytest.mark.skipif(not os.path.exists(Rf.LIB), reason="prebuilt oracle/_ref/libmbik_ref.so did not travel to this box")
def test_limit_sets_equal_the_reference_code_per_set():
    rig = rigs.humanoid22()
    n = 256
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 3, seed=8)
    h = R.create_limit_sets(sets)
    idx = (np.arange(n) % 3).astype(np.int32)
    T = rigs.random_targets(rig, 900, n)
    got = R.solve_with_limits(h, idx, T, want_local=True)
    want = _expected(rig, sets, idx, T, None, lambda r, t, **kw: Rf.solve_batch(r, t, threads=8, **kw))
    assert _same(got[1], want[1]) and _same(got[0], want[0]) and np.array_equal(got[2], want[2])
    R.destroy_limit_sets(h)


def test_limit_set_zero_equals_the_plain_solve_and_indices_are_clamped():
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 2)
    h = R.create_limit_sets(sets)
    n = 1500
    T = rigs.random_targets(rig, 0, n)
    plain = R.solve(T, want_local=True, sched="throughput")
    got = R.solve_with_limits(h, np.full(n, -7, np.int32), T, want_local=True)  # clamped to set 0 = the rig's own limits
    for a, b in zip(got, plain):
        assert _same(a, b)
    hi = R.solve_with_limits(h, np.full(n, 99, np.int32), T, want_local=True)   # clamped to the last set
    one = R.solve_with_limits(h, np.ones(n, np.int32), T, want_local=True)
    for a, b in zip(hi, one):
        assert _same(a, b)
    R.destroy_limit_sets(h)


def test_limit_sets_on_a_rig_beyond_128_solved_bones():
    """mbik_solve_batch_limits through the {256, 256, 32} variant (tail layout): per-set rigs in the oracle."""
    rig = rig_cases.LARGE_RIGS["big_tree240"]()
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 3)
    h = R.create_limit_sets(sets)
    n = 96
    T = rigs.random_targets(rig, 0, n)
    idx = (np.arange(n) % 3).astype(np.int32)
    want = _expected(rig, sets, idx, T, None, lambda r, t, **kw: O.solve_batch(r, t, threads=8, **kw))
    got = R.solve_with_limits(h, idx, T, want_local=True)
    for a, b in zip(got, want):
        assert _same(a, b)
    R.destroy_limit_sets(h)


def test_limit_sets_device_io_equals_host_io():
    import torch
    rig = rigs.quad80()
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 3)
    h = R.create_limit_sets(sets)
    n = 148 * 512 + 77  # more than one kernel wave: the host path pipelines chunks
    idx = (np.arange(n) % 3).astype(np.int32)
    T = rigs.random_targets(rig, 0, n)
    host = R.solve_with_limits(h, idx, T)
    t_dev, i_dev = torch.from_numpy(T).cuda(), torch.from_numpy(idx).cuda()
    o_dev = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device="cuda")
    R.solve_with_limits_raw(h, n, i_dev, t_dev, o_dev, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert _same(o_dev.cpu().numpy(), host[0])
    R.destroy_limit_sets(h)
