"""CPU tests of the product's HOST side (no compute calls, no GPU): the C-ABI library loads and exports every
symbol include/mbik.h declares, the rig flattener reproduces the reference's setup constants bit for bit
(checked against the oracle's object-graph restatement of ManyBoneIK3D::_bone_list_changed), error behaviour,
and the contiguous pose sharding (world_size-2 gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, MbikError, _capi, rigs, sharding
from oracle import oracle_py as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL_RIGS = dict(rigs.RIGS)
ALL_RIGS.update(rig_cases.EDGE_RIGS)


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "mbik.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mbik_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes found in include/mbik.h"
    assert declared == set(_capi.EXPORTED_SYMBOLS), declared ^ set(_capi.EXPORTED_SYMBOLS)
    lib = _capi.load_library()
    for s in declared:
        assert hasattr(lib, s), f"libmbik.so does not export {s}"
    # and nothing else leaks out of the library (hidden visibility): only mbik_* in the dynamic symbol table
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert defined >= declared
    assert all(d.startswith("mbik_") or d in ("_init", "_fini") for d in defined), sorted(defined - declared)[:10]


def test_library_has_sm100a_code_only():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_oracle_in_the_product():
    """The product package must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "many_bone_ik_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src.lower() or f in ("build.py", "mbik_math.cuh", "mbik_flatten.cu"), f
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                assert "libmbik_ref" not in src and "_ref/" not in src and "godot_shim" not in src, f
    ldd = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "mbik_ref" not in ldd


@pytest.mark.parametrize("name", sorted(ALL_RIGS))
def test_flattener_matches_reference_setup(name):
    """bone_list order, kept-segment count, per-step heading weights, bone-direction and twist frames, cone
    tangent geometry: flat host code (mbik_flatten.cu) == object-graph restatement (oracle), bit for bit."""
    rig = ALL_RIGS[name]()
    R = BatchedIKRig(rig)
    F = O.rig_facts(rig)
    assert np.array_equal(R.bone_order(), F["bone_order"])
    assert R.info["n_solved"] == len(F["bone_order"])
    assert R.info["n_segments"] == F["n_segments"]
    d, t = R.bone_frames()
    assert np.array_equal(d, F["dir_basis"], equal_nan=True)
    assert np.array_equal(t, F["twist_basis"], equal_nan=True)
    for s in range(R.info["n_steps"]):
        assert np.array_equal(R.step_weights(s), O.step_weights(rig, s)), f"step {s}"
    assert np.array_equal(R.cone_geometry(), O.cone_geometry(rig), equal_nan=True)
    R.close()


@pytest.mark.parametrize("seed", range(40))
def test_flattener_matches_reference_setup_random_rigs(seed):
    """Fuzz: random skeleton trees / pins / kusudama rows (tests/rig_cases.py::random_rig)."""
    rig = rig_cases.random_rig(seed)
    R = BatchedIKRig(rig)
    F = O.rig_facts(rig)
    assert np.array_equal(R.bone_order(), F["bone_order"])
    assert R.info["n_segments"] == F["n_segments"]
    d, t = R.bone_frames()
    assert np.array_equal(d, F["dir_basis"], equal_nan=True)
    assert np.array_equal(t, F["twist_basis"], equal_nan=True)
    for s in range(R.info["n_steps"]):
        assert np.array_equal(R.step_weights(s), O.step_weights(rig, s)), f"step {s}"
    assert np.array_equal(R.cone_geometry(), O.cone_geometry(rig), equal_nan=True)


@pytest.mark.parametrize("first", [0, 60, 120])
def test_flattener_matches_reference_setup_soak_rigs(first):
    """The same on the rigs of the GPU parity soaks (rig_cases.soak_rig: every 8th a dense tree of 64 ... 900 bones), 60 per case,
    plus the schedule's own consistency check against the variant that will run it (done by mbik_rig_create)."""
    for k in range(first, first + 60):
        rig = rig_cases.soak_rig(k)
        R = BatchedIKRig(rig)
        F = O.rig_facts(rig)
        assert np.array_equal(R.bone_order(), F["bone_order"]), k
        assert R.info["n_segments"] == F["n_segments"], k
        d, t = R.bone_frames()
        assert np.array_equal(d, F["dir_basis"], equal_nan=True) and np.array_equal(t, F["twist_basis"], equal_nan=True), k
        for s in range(0, R.info["n_steps"], 7 if R.info["n_steps"] > 64 else 1):
            assert np.array_equal(R.step_weights(s), O.step_weights(rig, s)), (k, s)
        assert np.array_equal(R.cone_geometry(), O.cone_geometry(rig), equal_nan=True), k
        R.close()


def test_schedule_facts_of_the_benchmark_rigs():
    """The segment structure SURVEY.md section 8(d) states for the canonical rigs."""
    h = BatchedIKRig(rigs.humanoid22())
    assert (h.info["n_bones"], h.info["n_solved"], h.info["n_segments"], h.info["n_effectors"]) == (22, 20, 7, 5)
    assert h.info["max_headings"] == 25 and h.info["n_cones"] == 32 and h.info["iterations"] == 10
    names = rigs.humanoid22().bone_names
    order = [names[b] for b in h.bone_order()]
    assert order[0] == "Head" and order[-1] == "Hips"
    c = BatchedIKRig(rigs.chain64())
    assert (c.info["n_solved"], c.info["n_segments"], c.info["n_effectors"], c.info["max_headings"]) == (64, 9, 9, 45)
    assert c.info["iterations"] == 30
    q = BatchedIKRig(rigs.quad80())
    assert (q.info["n_bones"], q.info["n_solved"], q.info["n_segments"], q.info["n_effectors"]) == (80, 63, 9, 9)
    for r in (h, c, q):
        assert r.info["rig_blob_bytes"] % 16 == 0 and r.info["rig_blob_bytes"] < 200 * 1024
        assert r.info["flops_per_solve"] > 0


def test_segment_parallel_schedule_facts():
    """Sibling segments are independent in segment_solver's post-order recursion (reference
    src/ik_bone_segment_3d.cpp:210-225).  humanoid22: phase 0 = {head, 2 arms, 2 legs} on 5 warps, phase 1 = spine,
    phase 2 = hips; a chain has nothing to run concurrently.  (mbik_rig_create also range- and order-checks the
    schedule: validate_schedule.)"""
    h = BatchedIKRig(rigs.humanoid22()).info
    assert (h["sp_roles"], h["sp_phases"]) == (5, 3) and h["sp_gain"] > 1.8
    c = BatchedIKRig(rigs.chain64()).info
    assert c["sp_roles"] == 1 and c["sp_phases"] == 9 and c["sp_gain"] == 1.0
    q = BatchedIKRig(rigs.quad80()).info
    assert 2 <= q["sp_roles"] <= 8 and q["sp_gain"] > 1.5
    t = BatchedIKRig(rig_cases.EDGE_RIGS["two_roots"]()).info  # independent skeleton roots run side by side
    assert (t["sp_roles"], t["sp_phases"]) == (2, 1)
    for seed in range(24):
        i = BatchedIKRig(rig_cases.random_rig(seed)).info
        assert 1 <= i["sp_roles"] <= 8 and 1 <= i["sp_phases"] <= i["n_segments"] and i["sp_gain"] >= 1.0


def test_segment_parallel_teams():
    """Warps a phase leaves idle join its one busy multi-effector segment as heading helpers: humanoid22's spine
    (3 effectors: head, hands) gets a team of 3 and its hips (5 effectors) a team of 5; phases with several busy
    segments and single-effector segments keep one warp per segment; stabilisation / constraint mode never team up."""
    rows = BatchedIKRig(rigs.humanoid22()).schedule()
    by_phase = {ph: rows[rows[:, 0] == ph] for ph in range(3)}
    assert (by_phase[0][:, 4] == 1).all() and len(by_phase[0]) == 5
    assert sorted(by_phase[1][:, 5]) == [0, 1, 2] and (by_phase[1][:, 4] == 3).all() and len(set(map(tuple, by_phase[1][:, 2:4]))) == 1
    assert sorted(by_phase[2][:, 5]) == [0, 1, 2, 3, 4] and (by_phase[2][:, 4] == 5).all()
    q = BatchedIKRig(rigs.quad80()).schedule()
    assert (q[q[:, 0] == 1][:, 4] == 1).all()      # spine and tail run side by side: no helpers
    assert (q[q[:, 0] == 2][:, 4] == 6).all()      # the pelvis step alone: all six warps
    for name in ("humanoid_stabilized", "humanoid_constraint_mode"):
        assert (BatchedIKRig(rig_cases.EDGE_RIGS[name]()).schedule()[:, 4] == 1).all()
    # every step is owned exactly once per iteration
    for f in list(rigs.RIGS.values()) + [rig_cases.EDGE_RIGS[n] for n in sorted(rig_cases.EDGE_RIGS)]:
        R = BatchedIKRig(f())
        owned = np.zeros(R.info["n_steps"], int)
        for ph, w, s0, s1, team, member in R.schedule():
            if member == 0:
                owned[s0:s1] += 1
        assert (owned == 1).all(), R.rig.name


def test_flops_floor_matches_survey_formula():
    """humanoid22 floor = 30.5 kflop/iteration -> ~305 kflop per 10-iteration solve."""
    h = BatchedIKRig(rigs.humanoid22())
    assert abs(h.info["flops_per_solve"] - 305e3) / 305e3 < 0.02


def test_rig_create_argument_errors():
    lib = _capi.load_library()
    h = C.c_void_p()
    assert lib.mbik_rig_create(None, C.byref(h)) == -1
    rig = rigs.humanoid22()
    desc, keep = _capi.rig_to_desc(rig)
    assert lib.mbik_rig_create(C.byref(desc), None) == -1
    # parent index out of range / not topologically ordered -> invalid argument, never a crash
    bad = rigs.humanoid22()
    bad.parent = bad.parent.copy()
    bad.parent[3] = 99
    with pytest.raises(MbikError) as ei:
        BatchedIKRig(bad)
    assert ei.value.code == -1
    bad2 = rigs.humanoid22()
    bad2.pins = [dict(bone=500, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2))]
    # a pin naming a bone the skeleton does not have never matches, as in the reference (pins are looked up by
    # bone name, src/ik_bone_3d.cpp:209-222): no effector, nothing solved, still a valid rig
    R2 = BatchedIKRig(bad2)
    assert R2.info["n_effectors"] == 0 and R2.info["n_pins"] == 1
    assert lib.mbik_strerror(-4).decode().startswith("no CUDA device")
    assert lib.mbik_rig_destroy(None) == 0


def test_rigs_beyond_every_compiled_capacity_select_the_unbounded_variant():
    """The reference has no bone limit (src/ik_bone_segment_3d.cpp:352-427).  300 solved bones are past the largest compiled
    variant {256, 256, 32}: the rig is accepted and runs the unbounded variant (state in a global workspace), whose
    capacity is that of the schedule's index types."""
    r = rig_cases.chain300()
    R = BatchedIKRig(r)
    assert R.info["n_solved"] == 300 and R.info["kernel_capacity"] == 16383
    ref = O.rig_facts(r)
    assert np.array_equal(R.bone_order(), ref["bone_order"])


@pytest.mark.parametrize("name", sorted(rig_cases.LARGE_RIGS))
def test_rigs_up_to_256_solved_bones_are_accepted(name):
    """129..256 solved bones select the {256, 256, 32} kernel variant; the flattener's bone order and weights still
    equal the oracle's setup."""
    rig = rig_cases.LARGE_RIGS[name]()
    R = BatchedIKRig(rig)
    F = O.rig_facts(rig)
    assert 128 < R.info["n_solved"] <= 256
    assert R.info["kernel_capacity"] == 256
    assert np.array_equal(R.bone_order(), F["bone_order"])
    assert R.info["n_segments"] == F["n_segments"]
    d, t = R.bone_frames()
    assert np.array_equal(d, F["dir_basis"], equal_nan=True)
    assert np.array_equal(t, F["twist_basis"], equal_nan=True)
    for s in range(R.info["n_steps"]):
        assert np.array_equal(R.step_weights(s), O.step_weights(rig, s)), f"step {s}"
    assert np.array_equal(R.cone_geometry(), O.cone_geometry(rig), equal_nan=True)
    R.close()


def test_solve_without_a_gpu_fails_loudly():
    """There is no CPU fallback: on a box without a CUDA device the solve entry point returns
    MBIK_ERR_NO_DEVICE / MBIK_ERR_CUDA and never produces numbers."""
    from many_bone_ik_b200 import device_count
    if device_count() > 0:
        pytest.skip("a CUDA device is present")
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    T = rigs.random_targets(rig, 0, 4)
    with pytest.raises(MbikError) as ei:
        R.solve(T)
    assert ei.value.code in (-4, -2)
    with pytest.raises(MbikError):
        R.solve(T, devices=[0, 1])


def test_solve_argument_validation():
    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    with pytest.raises(ValueError):
        R.solve(np.zeros((4, 3, 12), np.float32))
    with pytest.raises(ValueError):
        R.solve(np.zeros((4, 5, 12), np.float32), start_pose=np.zeros((4, 3, 12), np.float32))
    lib = _capi.load_library()
    p = _capi.SolveParams(-1, -1, 0, None)
    assert lib.mbik_solve_batch(None, C.byref(p), 1, None, None, None, None, None) == -1
    assert lib.mbik_solve_batch(R.handle, C.byref(p), 4, None, None, None, None, None) == -1


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 4096, (1 << 20) + 3):
        for w in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_target_slices_regenerate_independently():
    """Any shard regenerates its slice of the batch from the counter-based RNG (SURVEY 8(d) inputs)."""
    rig = rigs.quad80()
    full = rigs.random_targets(rig, 0, 100)
    a, b = sharding.shard_range(100, 1, 3)
    assert np.array_equal(rigs.random_targets(rig, a, b - a), full[a:b])


_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from many_bone_ik_b200 import rigs, sharding
from oracle import oracle_py as O
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rig = rigs.humanoid22()
N = 24
lo, hi = sharding.shard_range(N, rank, world)
T = rigs.random_targets(rig, lo, hi - lo)           # this rank's slice only
out, st = O.solve_batch(rig, T)                      # stand-in for the per-rank solve: the data path has NO collective
# bench-style bookkeeping only: max-over-ranks time and total count
t = torch.tensor([float(rank + 1)], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
c = torch.tensor([hi - lo], dtype=torch.int64); dist.all_reduce(c)
gathered = [None] * world
dist.all_gather_object(gathered, (lo, hi, out))
if rank == 0:
    full_T = rigs.random_targets(rig, 0, N)
    ref, _ = O.solve_batch(rig, full_T)
    got = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
    assert int(c.item()) == N and t.item() == world
    assert np.array_equal(got, ref), "result depends on the shard"
    print("GLOO_OK")
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding(tmp_path):
    """world_size-2 (gloo, CPU): each rank regenerates and solves only its contiguous slice; the concatenation
    equals the unsharded result bit for bit.  (The per-rank solve is the oracle here -- no GPU in this test; the
    GPU equivalent is tests/test_parity_gpu.py::test_multi_device_shard_invariance.)"""
    w = tmp_path / "worker.py"
    w.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(w), ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK" in r.stdout


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints one JSON line with the contract keys (bounded, CPU only)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--budget-s", "2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ik_skeleton_solves_per_sec" and line["unit"] == "solves/s"
    assert line["value"] > 0 and line["higher_is_better"] is True
    # "reference" = oracle/_ref (the reference module's own sources) when that library is built, else the restatement
    from oracle import reference_py as Rf
    assert line["cpu_baseline"]["kind"] == ("reference" if Rf.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    if line["cpu_baseline"]["kind"] == "reference":
        assert line["cpu_baseline"]["port"]["kind"] == "port" and line["cpu_baseline"]["port"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_numa_helpers_degrade_to_a_no_op():
    """many_bone_ik_b200/numa.py (bench / multi-GPU host glue): list parsing, and binding to the NUMA node of a GPU that
    does not exist or exposes no topology must be a reported no-op, never an exception."""
    from many_bone_ik_b200 import numa
    assert numa._parse_list("0-3,8,10-11") == [0, 1, 2, 3, 8, 10, 11] and numa._parse_list("") == [] and numa._parse_list(None) == []
    before = os.sched_getaffinity(0)
    done = numa.bind_to_gpu_node(0)
    assert "node" in done and ("skipped" in done or done["node"] is not None)
    if "skipped" in done:
        assert os.sched_getaffinity(0) == before


def test_stabilised_rigs_have_no_effector_list_limit():
    """Round 1 rejected stabilisation passes with more than 32 effectors in a list; the reference has no such limit."""
    for n_arms, arm_len, cap in ((40, 1, 64), (70, 2, 256), (90, 3, 16383)):
        R = BatchedIKRig(rig_cases.star_stabilized(n_arms, arm_len))
        assert R.info["n_effectors"] == n_arms and R.info["kernel_capacity"] == cap
