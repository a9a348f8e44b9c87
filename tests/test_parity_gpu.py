"""GPU parity: the CUDA solve (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: BIT-EXACT.  BASELINE.json states 1e-4 rad / 1e-5 chain-length tolerances, but the reference's constraint
snaps amplify float32 rounding differences by O(chain length) per iteration, so anything short of reproducing
the reference arithmetic bit for bit drifts out of tolerance on long chains; the kernel therefore performs the
same individually rounded IEEE operations as the reference, and these tests assert equality (NaN == NaN,
-0 == +0).  The north_star tolerances are asserted as well (test_north_star_tolerances) so the stated bar is
written down in a test."""
import os

import numpy as np
import pytest

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, MbikError, _capi, rigs
from oracle import oracle_py as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _assert_same(rig, got, ref):
    (out, loc, st), (ref_out, ref_loc, ref_st) = got, ref
    n = loc.shape[0]
    bad = np.argwhere(~np.all((loc == ref_loc) | (np.isnan(loc) & np.isnan(ref_loc)), axis=(1, 2))).ravel()
    assert bad.size == 0, f"{rig.name}: {bad.size}/{n} poses differ in local transforms, first {bad[:5]}, max abs diff {np.nanmax(np.abs(loc - ref_loc))}"
    assert np.array_equal(out, ref_out, equal_nan=True), f"{rig.name}: out_pose differs, max abs diff {np.nanmax(np.abs(out - ref_out))}"
    assert np.array_equal(st, ref_st)


# Both kernel mappings are product paths (the library picks by batch size): tests that probe one behaviour on a small
# batch run it through each mapping explicitly.
SCHEDS = ("throughput", "segment_parallel")


def _compare(rig, n, iterations=-1, start_pose=None, first=0):
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, first, n)
    ref = O.solve_batch(rig, T, start_pose=start_pose, iterations=iterations, want_local=True, threads=8)
    # both kernel mappings (one thread per pose in lockstep CTAs / one warp per concurrently solvable segment) and
    # whatever the library picks for this batch size: all three must equal the oracle bit for bit
    for sched in ("throughput", "segment_parallel", "auto"):
        got = R.solve(T, start_pose=start_pose, iterations=iterations, want_local=True, sched=sched)
        _assert_same(rig, got, ref)
    return got


@pytest.mark.parametrize("name,n", [("humanoid22", 4096), ("chain64", 256), ("quad80", 256)])
def test_bit_exact_default_configs(name, n):
    _compare(rigs.RIGS[name](), n)


@pytest.mark.parametrize("name,n", [("humanoid22", 6000), ("quad80", 4800), ("big_tree120", 4800)])
def test_segment_parallel_several_groups_per_sm(name, n):
    """Batches of more than one 32-pose group per SM run the 128-register build of the segment-parallel kernel
    (mbik_kernel.cu: segment_parallel_choice); same bits as the oracle, ragged last group included."""
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    rig = cases[name]()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 0, n)
    ref = O.solve_batch(rig, T, want_local=True, threads=8)
    for sched in ("segment_parallel", "auto"):
        _assert_same(rig, R.solve(T, want_local=True, sched=sched), ref)


@pytest.mark.parametrize("name,n", [("humanoid22", 4096 + 17), ("quad80", 2048), ("big_tree120", 1024)])
def test_segment_parallel_repeats_bitwise(name, n):
    """The warps of a pose group exchange poses and headings through shared memory between barriers; a missing or
    misplaced barrier would show as run-to-run differences.  25 repeats of a full-machine batch must return the same
    bits every time (compute-sanitizer's racecheck is not available on this pool), and equal the lockstep mapping."""
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    rig = cases[name]()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 3, n)
    first = R.solve(T, want_local=True, sched="throughput")
    for rep in range(25):
        got = R.solve(T, want_local=True, sched="segment_parallel")
        for a, b in zip(got, first):
            assert np.array_equal(a, b, equal_nan=True), f"{name}: repeat {rep} differs"


@pytest.mark.parametrize("iterations", [0, 1, 2, 15])
def test_humanoid_iteration_counts(iterations):
    _compare(rigs.humanoid22(), 257, iterations=iterations)


@pytest.mark.parametrize("name", sorted(rig_cases.EDGE_RIGS))
def test_bit_exact_edge_rigs(name):
    """multi-bone translating root segment, divergence to NaN + write-back reset, two skeleton roots, QCP
    single-heading branch, 3-axis pins, mpf = 0 cut-off, weight-0 pins, zero-cone rows, constraint_mode."""
    out, loc, st = _compare(rig_cases.EDGE_RIGS[name](), 192)
    if name == "chain_diverging":
        assert st.any(), "this rig is meant to exercise the non-finite reset path"


@pytest.mark.parametrize("name", sorted(rig_cases.LARGE_RIGS))
def test_bit_exact_rigs_beyond_128_solved_bones(name):
    """129..256 solved bones: the {256, 256, 32} kernel variant (thread-per-pose mapping only), with the rig constants
    fully shared-memory resident (chain150) and in the tail layout (chain200, big_tree240: walk list in global memory);
    a batch larger than one CTA per SM's worth of poses is covered by the second call."""
    rig = rig_cases.LARGE_RIGS[name]()
    _compare(rig, 48)
    R = BatchedIKRig(rig)
    n = 700
    T = rigs.random_targets(rig, 100, n)
    idx = np.arange(0, n, 53)
    ref = O.solve_batch(rig, T[idx], want_local=True, threads=8)
    out, loc, st = R.solve(T, want_local=True)
    _assert_same(rig, (out[idx], loc[idx], st[idx]), ref)


@pytest.mark.parametrize("seed", range(24))
def test_bit_exact_random_rigs(seed):
    """Fuzz: random trees, pins (zero weights, mpf cut-offs, 0-3 priority axes), 0-4 cones per row, per-bone damping,
    stabilisation, constraint mode -- solve == oracle bit for bit."""
    _compare(rig_cases.random_rig(seed), 96)


@pytest.mark.parametrize("seed,n_bones,n_pins", [(201, 70, 12), (202, 100, 16), (203, 140, 24), (204, 200, 30), (205, 256, 40), (206, 300, 50),
                                                 (207, 420, 60), (208, 64, 10), (209, 128, 20), (210, 90, 14), (211, 330, 64), (212, 500, 80)])
def test_bit_exact_random_dense_rigs(seed, n_bones, n_pins):
    """The same fuzz on large random trees with many pins: 28 ... 290 solved bones, i.e. every compiled kernel variant (32 /
    64 / 128 / 256 bones, with and without stabilisation passes) and the unbounded one, in every mapping the variant has,
    plain and from a perturbed start pose."""
    rig = rig_cases.random_rig(seed, n_bones=n_bones, n_pins=n_pins)
    _compare(rig, 128)
    _compare(rig, 64, start_pose=rig_cases.perturbed_start_pose(rig, 64, seed=seed), first=7)


@pytest.mark.parametrize("n_arms,arm_len", [(40, 1), (40, 2), (70, 2), (90, 3)])
def test_stabilisation_with_long_effector_lists(n_arms, arm_len):
    """Stabilisation passes with 40 ... 90 effectors in one list (the reference has no limit, src/ik_bone_segment_3d.cpp:114-180):
    the 64 / 128 / 256-bone variants and the unbounded one, every mapping the variant has, plain and from a perturbed pose."""
    rig = rig_cases.star_stabilized(n_arms, arm_len)
    _compare(rig, 96)
    _compare(rig, 40, start_pose=rig_cases.perturbed_start_pose(rig, 40, seed=n_arms), first=3)


@pytest.mark.parametrize("k", [39, 343])
def test_soak_rigs_that_exposed_the_effector_list_limit(k):
    """The two rigs of the 400-rig parity soak (profiles/r2_fuzz_soak.log) that the stabilisation variants' former 32-effector
    limit rejected: 68 effectors on the unbounded variant, 86 on the 256-bone one."""
    _compare(rig_cases.soak_rig(k), 64)


@pytest.mark.parametrize("name", ["humanoid22", "quad80", "star_mixed_pins"])
def test_bit_exact_with_start_pose(name):
    """Warm start: seeding from caller-supplied local poses (IKBone3D::set_initial_pose, src/ik_bone_3d.cpp:161)."""
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    rig = cases[name]()
    n = 96
    _compare(rig, n, start_pose=rig_cases.perturbed_start_pose(rig, n))


def test_frame_to_frame_warm_start_chain():
    """Three consecutive frames, each seeded with what the skeleton holds after the previous frame's write-back -- the
    recomposed position / rotation / scale (MBIK_LOCAL_RECOMPOSED) -- as the reference re-seeds from the skeleton after
    every solve (src/many_bone_ik_3d.cpp:1084, :91-102).  Oracle side: its out_pose recomposed by the engine stand-in."""
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 64
    start = None
    ref_start = None
    for frame in range(3):
        T = rigs.random_targets(rig, 1000 * frame, n)
        raw = R.solve(T, start_pose=start, want_local=True)
        got = R.solve(T, start_pose=start, want_local=True, recomposed_local=True)
        ref = O.solve_batch(rig, T, start_pose=ref_start, want_local=True, threads=8)
        _assert_same(rig, raw, ref)
        ref_start = O.recompose_pose(rig, ref[0], ref_start)
        assert np.array_equal(got[0], ref[0], equal_nan=True) and np.array_equal(got[1], ref_start, equal_nan=True), frame
        start = got[1]


@pytest.mark.parametrize("name", sorted(list(rigs.RIGS) + list(rig_cases.EDGE_RIGS)))
def test_cuda_reproduces_golden_fixtures(name):
    """The committed fixtures (tests/golden/oracle_solves.npz) -- no oracle call in this test."""
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    g = np.load(os.path.join(GOLD, "oracle_solves.npz"))
    rig = cases[name]()
    R = BatchedIKRig(rig)
    for sched in SCHEDS:
        out, loc, st = R.solve(g[name + "_targets"], want_local=True, sched=sched)
        assert np.array_equal(loc, g[name + "_local"], equal_nan=True), sched
        assert np.array_equal(out, g[name + "_out"], equal_nan=True), sched
        assert np.array_equal(st, g[name + "_status"]), sched


@pytest.mark.parametrize("n", [1, 31, 33, 383, 385, 148 * 384 + 5])
def test_ragged_batch_sizes(n):
    """Batch sizes around the warp / CTA / wave boundaries; pose k's result must not depend on the batch."""
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 0, n)
    m = min(n, 64)
    ref = O.solve_batch(rig, T[:m], want_local=True, threads=8)
    ref_t = O.solve_batch(rig, T[-32:], want_local=True, threads=8) if n > 64 else None
    for sched in SCHEDS + ("auto",):
        out, loc, st = R.solve(T, want_local=True, sched=sched)
        _assert_same(rig, (out[:m], loc[:m], st[:m]), ref)
        if n > 64:  # tail of the batch: last poses against the oracle too
            _assert_same(rig, (out[-32:], loc[-32:], st[-32:]), ref_t)


def test_empty_batch_is_a_no_op():
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    out, st = R.solve(np.zeros((0, 5, 12), np.float32))
    assert out.shape == (0, 22, 10) and st.shape == (0,)


def test_unsolved_bones_pass_through():
    """Bones outside bone_list keep their input pose (reference writes only solved bones, src/ik_bone_3d.cpp:170)."""
    rig = rigs.quad80()
    R = BatchedIKRig(rig)
    n = 16
    sp = rig_cases.perturbed_start_pose(rig, n)
    out, loc, st = R.solve(rigs.random_targets(rig, 0, n), start_pose=sp, want_local=True)
    solved = set(int(b) for b in R.bone_order())
    unsolved = [b for b in range(rig.n_bones) if b not in solved]
    assert len(unsolved) == 17
    assert np.array_equal(loc[:, unsolved], sp[:, unsolved])


def test_device_io_path_equals_host_io_path():
    """MBIK_IO_DEVICE (caller's device buffers + stream, asynchronous) == MBIK_IO_HOST (pipelined staging)."""
    import torch
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 3 * 148 * 384 + 77  # several chunks of the host pipeline
    T = rigs.random_targets(rig, 0, n)
    out_h, loc_h, st_h = R.solve(T, want_local=True)
    dev = torch.device("cuda", 0)
    t_d = torch.from_numpy(T).to(dev)
    o_d = torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device=dev)
    l_d = torch.empty((n, rig.n_bones, 12), dtype=torch.float32, device=dev)
    s_d = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        R.solve_raw(n, t_d, o_d, out_local=l_d, out_status=s_d, device=0, flags=_capi.MBIK_IO_DEVICE, stream=st.cuda_stream)
    st.synchronize()
    assert np.array_equal(o_d.cpu().numpy(), out_h, equal_nan=True)
    assert np.array_equal(l_d.cpu().numpy(), loc_h, equal_nan=True)
    assert np.array_equal(s_d.cpu().numpy().astype(np.uint32), st_h)
    assert R.last_kernel_ms(0) > 0


def test_multi_device_shard_invariance():
    """mbik_solve_batch_multi: pose k's result is bitwise independent of the number of shards.
    With one visible GPU the same device is listed several times, which still exercises the split."""
    from many_bone_ik_b200 import device_count
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 5000
    T = rigs.random_targets(rig, 0, n)
    base = R.solve(T, want_local=True)
    ndev = device_count()
    for devs in ([0], [0, 0], [0, 0, 0], list(range(ndev))):
        got = R.solve(T, want_local=True, devices=devs)
        for a, b in zip(got, base):
            assert np.array_equal(a, b, equal_nan=True), devs


def test_invalid_device_ordinal():
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    with pytest.raises(MbikError) as ei:
        R.solve(rigs.random_targets(rig, 0, 4), device=99)
    assert ei.value.code == -1


def _quat_angle(qa, qb):
    # angle between two unit quaternions (double cover aware); asin form is accurate near 0, unlike acos(dot)
    a, b = qa.astype(np.float64), qb.astype(np.float64)
    d = np.minimum(np.linalg.norm(a - b, axis=-1), np.linalg.norm(a + b, axis=-1))
    return 4.0 * np.arcsin(np.clip(d / 2.0, 0.0, 1.0))


@pytest.mark.parametrize("name", ["humanoid22", "chain64", "quad80"])
def test_north_star_tolerances(name):
    """BASELINE.json north_star: per-bone rotations within 1e-4 rad, effector residuals within 1e-5 of chain
    length, constraint satisfaction identical.  (Implied by bit-exactness; stated here as the written bar.)"""
    rig = rigs.RIGS[name]()
    R = BatchedIKRig(rig)
    n = 128
    T = rigs.random_targets(rig, 7, n)
    out, loc, st = R.solve(T, want_local=True)
    ref_out, ref_loc, ref_st = O.solve_batch(rig, T, want_local=True, threads=8)
    ang = _quat_angle(out[..., 3:7], ref_out[..., 3:7])
    assert np.nanmax(ang) <= 1e-4
    chain_len = float(np.sum(np.linalg.norm(rig.rest_local[:, 9:], axis=1)))

    def effector_origins(local):
        res = np.zeros((n, rig.n_pins, 3))
        for k in range(n):
            _, t = rigs.global_rest(rig.parent, local[k])
            res[k] = t[[p["bone"] for p in rig.pins]]
        return res

    resid = np.linalg.norm(effector_origins(loc) - T[..., 9:], axis=-1)
    resid_ref = np.linalg.norm(effector_origins(ref_loc) - T[..., 9:], axis=-1)
    assert np.max(np.abs(resid - resid_ref)) <= 1e-5 * chain_len
    assert np.array_equal(st, ref_st)


def test_large_batch_properties_at_bench_size():
    """At the bench's full size (2^20 poses) the oracle is too slow; size-independent properties instead:
    (1) a strided sample of the big batch equals the oracle, (2) solving the batch in a permuted order gives the
    permuted result (no cross-pose coupling), (3) every output quaternion is unit, every status is 0."""
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 1 << 20
    T = np.concatenate([rigs.random_targets(rig, s, 1 << 16) for s in range(0, n, 1 << 16)])
    out, st = R.solve(T)
    idx = np.arange(0, n, 4099)
    ref_out, ref_st = O.solve_batch(rig, T[idx], threads=8)
    assert np.array_equal(out[idx], ref_out, equal_nan=True)
    perm = np.random.default_rng(3).permutation(n)[: 1 << 18]
    out_p, st_p = R.solve(T[perm])
    assert np.array_equal(out_p, out[perm], equal_nan=True)
    qn = np.linalg.norm(out[..., 3:7].astype(np.float64), axis=-1)
    assert np.all(np.abs(qn - 1.0) < 1e-5)
    assert not st.any()


@pytest.mark.parametrize("name", ["chain64", "quad80"])
def test_large_rig_properties_at_baseline_batch(name):
    """BASELINE configs 4 / 5 at their full 65536-pose batch: a strided sample equals the oracle, a permuted
    sub-batch gives the permuted result, statuses are all 0, and the 4096-pose batch of the same configs (which the
    library may run in the segment-parallel mapping) is the bitwise prefix of the big one."""
    rig = rigs.RIGS[name]()
    R = BatchedIKRig(rig)
    n = 1 << 16
    T = rigs.random_targets(rig, 0, n)
    out, st = R.solve(T)
    idx = np.arange(0, n, 2731)
    ref_out, ref_st = O.solve_batch(rig, T[idx], threads=8)
    assert np.array_equal(out[idx], ref_out, equal_nan=True)
    assert np.array_equal(st[idx], ref_st)
    perm = np.random.default_rng(5).permutation(n)[: 1 << 14]
    out_p, _ = R.solve(T[perm])
    assert np.array_equal(out_p, out[perm], equal_nan=True)
    out_s, _ = R.solve(T[:4096])
    assert np.array_equal(out_s, out[:4096], equal_nan=True)
    assert not st.any()


def test_guarded_sqrt_div_groups_are_correctly_rounded():
    """The kernel's grouped sqrt/division sequences (mbik_math.cuh) == __fsqrt_rn/__fdiv_rn bit for bit: sqrt
    exhaustively over [2^-80, 2^80), division over all 2^23 divisor mantissas x 4 rounds x 64 numerators,
    plus random vectors (zeros, denormals, inf/NaN included) through vnorm / q_normalized / sqrt_then_div."""
    import ctypes as C
    lib = _capi.load_library()
    n, bad = C.c_uint64(0), C.c_uint64(0)
    rc = lib.mbik_selftest(0, 4, C.byref(n), C.byref(bad))
    assert rc == 0
    assert n.value > 3_000_000_000, n.value
    assert bad.value == 0, f"{bad.value} mismatches out of {n.value}"


@pytest.mark.parametrize("name", ["humanoid22", "chain64", "star_mixed_pins"])
def test_cpp_facade_solve_equals_c_abi_solve(tmp_path, name):
    """The reference-named C++ front end (ManyBoneIK3D setters / .tscn property paths ->
    process_modification_batch) produces the same bits as the direct C-ABI path and the oracle."""
    from test_host_facade_cpu import build_driver, run_driver
    cases = dict(rigs.RIGS)
    cases.update(rig_cases.EDGE_RIGS)
    rig = cases[name]()
    exe = build_driver(tmp_path)
    n = 200
    T = rigs.random_targets(rig, 0, n)
    tb, ob = os.path.join(str(tmp_path), "t.bin"), os.path.join(str(tmp_path), "o.bin")
    T.tofile(tb)
    rc, facts, log = run_driver(exe, tmp_path, rig, [tb, ob, str(n)])
    assert rc == 0 and int(facts["solve"]) == 0, log
    got = np.fromfile(ob, np.float32).reshape(n, rig.n_bones, 10)
    ref_out, ref_st = O.solve_batch(rig, T, threads=8)
    assert np.array_equal(got, ref_out, equal_nan=True)


@pytest.mark.parametrize("name", ["humanoid22", "quad80"])
def test_warm_start_stream_equals_per_frame_calls(name):
    """mbik_stream_*: device-resident frame-to-frame warm start == per-frame mbik_solve_batch with MBIK_LOCAL_RECOMPOSED and
    start_pose = previous out_local == the oracle re-seeded from the recomposed previous frame (5 frames, pipelined)."""
    from many_bone_ik_b200 import IKStream
    rig = rigs.RIGS[name]()
    R = BatchedIKRig(rig)
    n, frames = 300, 5
    S = IKStream(R, n, device=0)
    Ts = [rigs.random_targets(rig, 777 * f, n) for f in range(frames)]
    outs = [np.empty((n, rig.n_bones, 10), np.float32) for _ in range(frames)]
    sts = [np.empty(n, np.uint32) for _ in range(frames)]
    for f in range(frames):  # all frames in flight before the first sync
        S.submit(Ts[f], outs[f], sts[f])
    S.sync()
    assert S.frames == frames
    final_local = S.read_local()
    start = cuda_start = None
    for f in range(frames):
        ref_out, ref_loc, ref_st = O.solve_batch(rig, Ts[f], start_pose=start, want_local=True, threads=8)
        assert np.array_equal(outs[f], ref_out, equal_nan=True), f"frame {f}"
        assert np.array_equal(sts[f], ref_st)
        start = O.recompose_pose(rig, ref_out, start)
        c_out, c_loc, c_st = R.solve(Ts[f], start_pose=cuda_start, want_local=True, recomposed_local=True)
        assert np.array_equal(c_out, ref_out, equal_nan=True) and np.array_equal(c_loc, start, equal_nan=True), f"frame {f}"
        cuda_start = c_loc
    assert np.array_equal(final_local, start, equal_nan=True)
    # reset -> rest pose again; and a caller-supplied initial pose
    S.reset()
    S.submit(Ts[0], outs[1], None)
    S.sync()
    assert np.array_equal(outs[1], outs[0], equal_nan=True)
    sp = rig_cases.perturbed_start_pose(rig, n)
    S2 = IKStream(R, n, device=0, initial_pose=sp)
    S2.submit(Ts[1], outs[2], None)
    S2.sync()
    ref_out, _ = O.solve_batch(rig, Ts[1], start_pose=sp, threads=8)
    assert np.array_equal(outs[2], ref_out, equal_nan=True)


def test_non_finite_targets_propagate_like_the_reference():
    """NaN / Inf / huge targets: the kernel's guarded fast paths must fall back exactly where IEEE semantics matter."""
    rig = rigs.humanoid22()
    n = 64
    T = rigs.random_targets(rig, 0, n)
    T[0, 0, 9] = np.nan
    T[1, 1, :9] = np.inf
    T[2, 2, 9:] = 3.0e38
    T[3, 3, 9:] = 1.0e-30
    T[4, :, :] = 0.0
    T[5, 4, 0] = -np.inf
    T[6, 0, 9:] = (1e20, -1e20, 1e-20)
    ref = O.solve_batch(rig, T, want_local=True, threads=8)
    for sched in SCHEDS:
        _assert_same(rig, BatchedIKRig(rig).solve(T, want_local=True, sched=sched), ref)
    q = rigs.quad80()
    Tq = rigs.random_targets(q, 0, 32)
    Tq[::3, ::2, 9:] *= np.float32(1e19)
    Tq[1::3, 1::2, :9] = 0.0
    refq = O.solve_batch(q, Tq, want_local=True, threads=8)
    for sched in SCHEDS:
        _assert_same(q, BatchedIKRig(q).solve(Tq, want_local=True, sched=sched), refq)


def test_concurrent_host_calls_on_one_rig():
    """mbik_rig is shareable across threads: concurrent host-buffer solves on one (rig, device) take turns on the
    staging lanes; concurrent device-buffer solves on separate streams overlap.  Every result equals the serial one."""
    import threading
    import torch
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    n = 20000
    Ts = [rigs.random_targets(rig, 50000 * k, n) for k in range(4)]
    serial = [R.solve(T)[0] for T in Ts]
    got = [None] * 4

    def work(k):
        got[k] = R.solve(Ts[k])[0]

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for k in range(4):
        assert np.array_equal(got[k], serial[k], equal_nan=True), k
    # device path, one stream per thread
    dev = torch.device("cuda", 0)
    outs = [torch.empty((n, rig.n_bones, 10), dtype=torch.float32, device=dev) for _ in range(4)]
    tds = [torch.from_numpy(T).to(dev) for T in Ts]
    streams = [torch.cuda.Stream(device=dev) for _ in range(4)]

    def work_dev(k):
        R.solve_raw(n, tds[k], outs[k], device=0, flags=_capi.MBIK_IO_DEVICE, stream=streams[k].cuda_stream)

    th = [threading.Thread(target=work_dev, args=(k,)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    for k in range(4):
        assert np.array_equal(outs[k].cpu().numpy(), serial[k], equal_nan=True), k


def test_rig_and_stream_lifecycle_does_not_leak_device_memory():
    import torch
    from many_bone_ik_b200 import IKStream
    rig = rigs.quad80()
    T = rigs.random_targets(rig, 0, 2048)

    def cycle():
        R = BatchedIKRig(rig)
        R.solve(T)
        S = IKStream(R, 2048)
        S.submit(T, None, None)
        S.sync()
        S.close()
        R.close()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    for _ in range(20):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < 8 * 1024 * 1024, (free0, free1)


@pytest.mark.parametrize("seed", range(16))
def test_overflow_fuzz_random_rigs(seed):
    """Targets scaled by 1e5 ... 1e37 (and 1e-30) on random rigs, 12 iterations: poses overflow to Inf / NaN part-way.
    Every guarded fast path and finite-operand shortcut of the kernel must hand over to the literal IEEE formulation
    exactly where it matters: raw locals equal the oracle NaN-for-NaN, status words equal."""
    rig = rig_cases.random_rig(seed)
    rig.iterations = 12
    rng = np.random.default_rng(seed)
    n = 64
    T = rigs.random_targets(rig, 0, n)
    T[:, :, 9:] *= np.float32(10.0) ** rng.integers(5, 38, size=(n, 1, 1)).astype(np.float32)
    if seed % 3 == 0:
        T[:, :, :9] *= np.float32(10.0) ** rng.integers(0, 20, size=(n, 1, 1)).astype(np.float32)
    if seed % 4 == 0:
        T[::5] = np.float32(1e-30) * T[::5]
    R = BatchedIKRig(rig)
    ref = O.solve_batch(rig, T, want_local=True, threads=8)
    for sched in SCHEDS:
        _assert_same(rig, R.solve(T, want_local=True, sched=sched), ref)
