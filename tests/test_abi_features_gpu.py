"""GPU tests of the round-2 C-ABI additions: solved-only output layout, newton_iters, wave-balanced CTA sizes, CTA sizes
whose scratch does not fit beside a large rig blob."""
import copy

import numpy as np
import pytest

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, _capi, rigs
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", ["humanoid22", "quad80", "two_roots", "no_pins"])
def test_solved_only_layout_equals_the_solved_rows_of_the_default_layout(name):
    """MBIK_OUT_SOLVED_ONLY: [n][n_solved][10] in bone_list order -- the bones set_skeleton_bone_pose writes
    (src/many_bone_ik_3d.cpp:104-116) -- bit-identical to the same bones' rows of the full layout, in both mappings,
    through the chunked host path (more than one kernel wave) and with a start pose."""
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    rig = cases[name]()
    R = BatchedIKRig(rig)
    order = R.bone_order()
    n = 148 * 512 + 333 if name == "humanoid22" else 500
    T = rigs.random_targets(rig, 11, n)
    start = rig_cases.perturbed_start_pose(rig, n, seed=2)
    for sp in (None, start):
        for sched in ("throughput", "segment_parallel", "auto"):
            full, st_full = R.solve(T, start_pose=sp, sched=sched)
            compact, st = R.solve(T, start_pose=sp, sched=sched, solved_only=True)
            assert compact.shape == (n, len(order), 10)
            assert _same(compact, full[:, order]) and np.array_equal(st, st_full), (sched, sp is None)


def test_newton_iters_zero_is_the_parity_path_and_more_changes_the_result():
    """mbik_solve_params::newton_iters: 0 == the reference (no eigenvalue refinement, src/math/qcp.cpp:205,215);
    > 0 is a different solver (Theobald's Newton-Raphson on the characteristic polynomial)."""
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 512
    T = rigs.random_targets(rig, 0, n)
    ref = O.solve_batch(rig, T, want_local=True, threads=8)
    for sched in ("throughput", "segment_parallel"):
        zero = R.solve(T, want_local=True, sched=sched, newton_iters=0)
        neg = R.solve(T, want_local=True, sched=sched, newton_iters=-3)
        for a, b, c in zip(zero, neg, ref):
            assert _same(a, b) and _same(a, c)
        newton = R.solve(T, want_local=True, sched=sched, newton_iters=30)
        assert not _same(newton[1], zero[1])
        assert np.isfinite(newton[1]).all()
    # the two mappings agree with each other under Newton too (same arithmetic, same order)
    a = R.solve(T, want_local=True, sched="throughput", newton_iters=30)
    b = R.solve(T, want_local=True, sched="segment_parallel", newton_iters=30)
    assert _same(a[1], b[1])


def _kabsch_quaternion(moved, target, w):
    """Optimal rotation taking `moved` onto `target` (Horn's 4x4 eigenproblem, float64) as xyzw."""
    M = (w[:, None, None] * target[:, :, None] * moved[:, None, :]).sum(0)  # sum w * target moved^T
    Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz = M.T.reshape(-1)  # S_ab = sum w moved_a target_b
    K = np.array([[Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx],
                  [Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz],
                  [Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy],
                  [Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz]])
    vals, vecs = np.linalg.eigh(K)
    q = vecs[:, -1]  # wxyz
    return np.array([q[1], q[2], q[3], q[0]])


def test_qcp_stage_with_newton_is_the_optimal_rotation():
    """the reference's QCP under-rotates whenever the fit is imperfect.  With newton_iters the stage
    returns the Kabsch / Horn optimum (to float32 precision), without it the reference's value (checked elsewhere)."""
    lib = _capi.load_library()
    rng = np.random.default_rng(4)
    worst_plain = 0.0
    for k in range(20):
        n = int(rng.integers(3, 20))
        moved = rng.normal(size=(n, 3)).astype(np.float32)
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = rng.uniform(0.2, 2.5)
        Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        Rm = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
        target = (moved.astype(np.float64) @ Rm.T + rng.normal(size=(n, 3)) * 0.3).astype(np.float32)
        w = rng.uniform(0.2, 1.0, n)
        want = _kabsch_quaternion(moved.astype(np.float64), target.astype(np.float64), w)
        out = np.zeros(7, np.float32)
        for iters, tol in ((0, None), (50, 2e-3)):
            rc = lib.mbik_stage_qcp_newton(0, n, moved.ctypes.data, target.ctypes.data, np.ascontiguousarray(w).ctypes.data, 0, iters, out.ctypes.data)
            assert rc == 0
            q = out[:4].astype(np.float64)
            err = 2 * np.arccos(min(1.0, abs(float(q @ want)) / np.linalg.norm(q)))
            if tol is None:
                worst_plain = max(worst_plain, err)
            else:
                assert err < tol, (k, err)
    assert worst_plain > 0.05  # the unrefined eigenvalue really is a different (smaller) rotation on noisy fits


@pytest.mark.parametrize("n", [131072, 40000, 148 * 512 * 2 + 1])
def test_wave_balanced_cta_sizes_keep_every_pose_bit_identical(n):
    """Large batches pick 384- / 448- / 512-thread CTAs so that the last wave is full (mbik_kernel.cu:
    throughput_block_threads); a pose's result cannot depend on it: slices of the big batch == the same poses solved as
    small batches (which run other CTA sizes / the other mapping) == the oracle."""
    import torch
    from many_bone_ik_b200._capi import MBIK_IO_DEVICE, MBIK_SCHED_THROUGHPUT
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 0, n)
    t_dev = torch.from_numpy(T).cuda()
    o_dev = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device="cuda")
    R.solve_raw(n, t_dev, o_dev, device=0, flags=MBIK_IO_DEVICE | MBIK_SCHED_THROUGHPUT, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    big = o_dev.cpu().numpy()
    for lo in (0, n // 2 - 100, n - 257):
        small, _ = R.solve(T[lo:lo + 257])
        assert _same(big[lo:lo + 257], small), lo
    ref, _ = O.solve_batch(rig, T[-64:], threads=8)
    assert _same(big[-64:], ref)


def _many_cones_chain(n_solved=60, cones_per_bone=10):
    """<= 64 solved bones with ~100 KB of cone data: the {64, 16, 8} kernel variant with a blob too large for the
    shared-memory scratch of its 128-thread CTAs."""
    rig = copy.deepcopy(rigs.chain64())
    rig.name = "chain64_many_cones"
    rig.iterations = 2
    rig.constraints = []
    rng = np.random.default_rng(1)
    for b in range(4, n_solved, 1):
        cones = []
        for j in range(cones_per_bone):
            d = np.array([0.0, 1.0, 0.0]) + rng.normal(size=3) * 0.4
            d /= np.linalg.norm(d)
            cones.append((float(np.float32(d[0])), float(np.float32(d[1])), float(np.float32(d[2])), float(np.float32(rng.uniform(0.3, 0.7)))))
        rig.constraints.append(dict(bone=b, twist_from=float(np.float32(-0.3)), twist_range=float(np.float32(0.9)), cones=cones))
    return rig


@pytest.mark.parametrize("n", [3000, 10000, 19000, 148 * 512 + 10000])
def test_large_blob_on_the_64_bone_variant_runs_at_every_batch_size(n):
    rig = _many_cones_chain()
    R = BatchedIKRig(rig)
    assert R.info["kernel_capacity"] == 64 and R.info["rig_blob_bytes"] > 84 * 1024
    T = rigs.random_targets(rig, 0, n)
    out, st = R.solve(T, sched="throughput")  # used to fail with MBIK_ERR_CUDA between 32 and 128 poses per SM
    small, _ = R.solve(T[:96], sched="throughput")
    assert _same(out[:96], small)
    ref, ref_st = O.solve_batch(rig, T[-48:], threads=8)
    assert _same(out[-48:], ref) and np.array_equal(st[-48:], ref_st)


def test_unbounded_variant_equals_the_oracle(monkeypatch):
    """Rigs past every compiled capacity (here 300 solved bones; the reference has no limit) run the variant that sizes
    nothing at compile time -- constants read in place, per-pose state in a stream-ordered global workspace, the batch cut
    into launches that reuse it.  Same bits as the oracle: plain, with a start pose, through the host and the device path,
    with several launches per call, with per-pose limit sets."""
    import torch
    import limit_set_cases as LS
    from many_bone_ik_b200._capi import MBIK_IO_DEVICE
    rig = rig_cases.chain300()
    R = BatchedIKRig(rig)
    assert R.info["kernel_capacity"] == 16383
    n = 200
    T = rigs.random_targets(rig, 3, n)
    sp = rig_cases.perturbed_start_pose(rig, n, seed=8)
    for start in (None, sp):
        ref = O.solve_batch(rig, T, start_pose=start, want_local=True, threads=8)
        got = R.solve(T, start_pose=start, want_local=True)
        for a, b in zip(got, ref):
            assert _same(a, b)
    monkeypatch.setenv("MBIK_DYN_THREADS", "128")  # 200 poses = two launches sharing one workspace
    got2 = R.solve(T, start_pose=sp, want_local=True)
    for a, b in zip(got2, got):
        assert _same(a, b)
    t_dev = torch.from_numpy(T).cuda()
    s_dev = torch.from_numpy(sp).cuda()
    o_dev = torch.empty((n, R.info["n_solved"], 10), dtype=torch.float32, device="cuda")
    R.solve_raw(n, t_dev, o_dev, start_pose=s_dev, device=0, flags=MBIK_IO_DEVICE | _capi.MBIK_OUT_SOLVED_ONLY, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert _same(o_dev.cpu().numpy(), got[0][:, R.bone_order()])
    monkeypatch.delenv("MBIK_DYN_THREADS")
    sets = LS.variants(rig, 2)
    h = R.create_limit_sets(sets)
    idx = (np.arange(64) % 2).astype(np.int32)
    out, st = R.solve_with_limits(h, idx, T[:64])
    for s_i in range(2):
        want, _ = O.solve_batch(LS.rig_with(rig, sets[s_i]), T[:64], threads=8)
        assert _same(out[idx == s_i], want[idx == s_i])
    R.destroy_limit_sets(h)


@pytest.mark.parametrize("name,n", [("chain64", 19000), ("chain64", 148 * 512 * 4 + 1000), ("chain150", 19000), ("chain200", 19000),
                                    ("quad80", 19000), ("quad80", 80000), ("big_tree120", 19000), ("big_tree240", 19000), ("random_rig_64a", 19000),
                                    ("random_rig_64b", 19000)])
def test_streamed_walk_instantiation_keeps_every_pose_bit_identical(name, n):
    """Large batches of rigs with long effector walks (chains) run the streamed-walk instantiation: local poses in a global
    float4 workspace, walk children through a cp.async ring (mbik_kernel_body.cuh, GLW); more than four waves = several
    launches sharing one workspace.  A pose's bits cannot depend on it: slices of the big batch == the same poses solved
    as small batches (thread-local state) == the oracle; host path (chunked, three streams) == device path."""
    import torch
    from many_bone_ik_b200._capi import MBIK_IO_DEVICE
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.LARGE_RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    # trees exercise the walk stack (branch points pushed and re-read) and short plain runs; random 40..64-bone rigs the rest
    cases["random_rig_64a"] = lambda: rig_cases.random_rig(3022, n_bones=100)  # 54 solved bones, walk stack 3
    cases["random_rig_64b"] = lambda: rig_cases.random_rig(3015, n_bones=64)   # 36 solved bones, walk stack 3
    rig = cases[name]()
    R = BatchedIKRig(rig)
    assert R.info["kernel_capacity"] >= 64, "the streamed walk belongs to the 64-bone-and-up variants"
    T = rigs.random_targets(rig, 0, n)
    t_dev = torch.from_numpy(T).cuda()
    o_dev = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device="cuda")
    st_dev = torch.empty(n, dtype=torch.int32, device="cuda")
    R.solve_raw(n, t_dev, o_dev, out_status=st_dev, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    big = o_dev.cpu().numpy()
    for lo in (0, n // 2 - 50, n - 130):
        small, st = R.solve(T[lo:lo + 130])
        assert _same(big[lo:lo + 130], small), lo
        assert np.array_equal(st_dev.cpu().numpy()[lo:lo + 130].astype(np.uint32), st)
    ref, _ = O.solve_batch(rig, T[-24:], threads=8)
    assert _same(big[-24:], ref)
    if n < 100000:
        host, _ = R.solve(T)
        assert _same(host, big)


def test_unbounded_variant_with_stabilisation_passes():
    rig = rig_cases.chain300()
    rig.stabilization_passes = 2
    rig.name = "chain300_stabilized"
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 9, 96)
    ref = O.solve_batch(rig, T, want_local=True, threads=8)
    got = R.solve(T, want_local=True)
    for a, b in zip(got, ref):
        assert _same(a, b)


def test_streamed_walk_with_limit_sets_and_with_stabilisation():
    """The streamed-walk instantiation also exists with per-pose limit sets and with stabilisation passes (large batches of
    long-walk rigs); slices of the big batch == the same poses solved as small batches == the oracle."""
    import limit_set_cases as LS
    n = 19000
    # limit sets
    rig = rigs.chain64()
    R = BatchedIKRig(rig)
    sets = LS.variants(rig, 3)
    h = R.create_limit_sets(sets)
    T = rigs.random_targets(rig, 5, n)
    idx = (np.arange(n) % 3).astype(np.int32)
    big, st = R.solve_with_limits(h, idx, T)
    small, _ = R.solve_with_limits(h, idx[-200:], T[-200:])
    assert _same(big[-200:], small)
    for s_i in range(3):
        want, _ = O.solve_batch(LS.rig_with(rig, sets[s_i]), T[:30], threads=8)
        m = idx[:30] == s_i
        assert _same(big[:30][m], want[m])
    R.destroy_limit_sets(h)
    # stabilisation
    rig = rig_cases.chain64_stabilized()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 6, n)
    big, st = R.solve(T, sched="throughput")
    small, _ = R.solve(T[-150:], sched="throughput")
    assert _same(big[-150:], small)
    want, wst = O.solve_batch(rig, T[:24], threads=8)
    assert _same(big[:24], want) and np.array_equal(st[:24], wst)
