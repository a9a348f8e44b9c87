"""Edge-case rigs for parity tests (beyond the three benchmark rigs)."""
import numpy as np

from many_bone_ik_b200 import rigs
from many_bone_ik_b200.rigs import DEG, Rig, _add_constraints, _axis_angle, _xf


def _chain(n, seg=0.1, curl=2.0):
    parent = np.arange(-1, n - 1, dtype=np.int32)
    rest = np.zeros((n, 12))
    for i in range(n):
        rest[i] = _xf(_axis_angle((0.2, 0.1, 1.0), (curl if i else 0.0) * DEG), (0.0, seg if i else 0.3, 0.01 * (i % 3)))
    return parent, rest.astype(np.float32)


def chain_multibone_root(n=12, iterations=3):
    """Unpinned root: the translating root segment spans several bones.  The reference
    diverges on this within ~7 iterations; 3 iterations stay finite."""
    parent, rest = _chain(n)
    r = Rig("chain_multibone_root", [f"b{i}" for i in range(n)], parent, rest, iterations=iterations, config_id=11)
    r.pins = [dict(bone=5, weight=0.5, mpf=1.0, priorities=(0.2, 0.0, 0.2)), dict(bone=n - 1, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2))]
    _add_constraints(r, {b: (1, 40, 0, -30, 60) for b in range(2, n, 3)})
    return r


def chain_diverging(n=12):
    """Same rig, enough iterations to blow up to non-finite values: NaN propagation and the write-back reset
    (reference src/ik_bone_3d.cpp:174-176) must match too."""
    r = chain_multibone_root(n, iterations=60)
    r.name = "chain_diverging"
    r.config_id = 12
    return r


def two_roots():
    """Two parentless bones: only the LAST root keeps an IKNode3D parent (ik_origin is re-instantiated per root,
    reference src/many_bone_ik_3d.cpp:1022), so the first root never rotates."""
    names = ["A0", "A1", "A2", "B0", "B1", "B2", "B3"]
    parent = np.array([-1, 0, 1, -1, 3, 4, 5], np.int32)
    rest = np.zeros((7, 12))
    for i in range(7):
        rest[i] = _xf(_axis_angle((1, 0.3, 0.2), (5.0 * (i % 3)) * DEG), (0.4 if i == 3 else 0.02 * i, 0.0 if parent[i] < 0 else 0.2, 0.0))
    r = Rig("two_roots", names, parent, rest.astype(np.float32), iterations=6, config_id=13)
    r.pins = [dict(bone=2, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2)), dict(bone=6, weight=0.7, mpf=1.0, priorities=(0.1, 0.3, 0.0))]
    _add_constraints(r, {1: (2, 40, 25, -20, 50), 5: (1, 30, 0, -45, 90)})
    return r


def star_mixed_pins():
    """Branching rig exercising: a pin with all priorities 0 that is alone in its list (QCP single-heading branch,
    reference src/math/qcp.cpp:59-78), a 3-axis pin (7 headings), mpf = 0 cut-off, weight 0 pin, a constraint row
    with zero cones, a constraint row on an unsolved bone, bone_damp shorter than the skeleton."""
    names = ["Root", "S1", "S2", "L1", "L2", "L3", "R1", "R2", "R3", "T1", "T2", "T3", "T4", "Leaf"]
    par = [-1, 0, 1, 2, 3, 4, 2, 6, 7, 0, 9, 10, 11, 8]
    parent = np.array(par, np.int32)
    rest = np.zeros((len(names), 12))
    offs = [(0, 1, 0), (0, .2, 0), (0, .2, .02), (.15, .05, 0), (.2, 0, 0), (.2, 0, .01), (-.15, .05, 0), (-.2, 0, 0), (-.2, 0, 0),
            (0, -.1, -.2), (0, 0, -.2), (0, .01, -.2), (0, 0, -.2), (-.05, 0, 0)]
    for i in range(len(names)):
        rest[i] = _xf(_axis_angle(((i % 3 == 0) * 1.0, (i % 3 == 1) * 1.0, (i % 3 == 2) * 1.0), ((i * 29) % 9 - 4) * DEG), offs[i])
    r = Rig("star_mixed_pins", names, parent, rest.astype(np.float32), iterations=8, config_id=14)
    r.pins = [
        dict(bone=5, weight=1.0, mpf=1.0, priorities=(0.0, 0.0, 0.0)),   # L3: translation only -> 1 heading, alone in the L chain's list
        dict(bone=8, weight=0.8, mpf=1.0, priorities=(0.3, 0.2, 0.3)),   # R3: 7 headings (has an unsolved child Leaf)
        dict(bone=10, weight=0.6, mpf=0.0, priorities=(0.2, 0.0, 0.2)),  # T2: mpf 0 cuts T4 off from Root's list
        dict(bone=12, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2)),  # T4
        dict(bone=2, weight=0.0, mpf=0.5, priorities=(0.2, 0.0, 0.2)),   # S2: weight 0 (reference default)
    ]
    _add_constraints(r, {1: (1, 35, 0, -30, 60), 3: (2, 40, 30, -45, 90), 4: (3, 30, 25, 0, 45), 7: (4, 25, 20, -10, 20), 11: (1, 20, 0, -90, 180)})
    r.constraints.append(dict(bone=9, twist_from=-0.3, twist_range=1.0, cones=[]))        # zero cones
    r.constraints.append(dict(bone=13, twist_from=0.0, twist_range=1.0, cones=[(0, 1, 0, 0.5)]))  # unsolved bone: ignored
    r.bone_damp = np.array([0.05, 0.02, 0.2, 0.01, 0.03], np.float32)  # indexed by bone id, shorter than the skeleton
    return r


def humanoid_no_constraints():
    r = rigs.humanoid22()
    r.constraints = []
    r.name = "humanoid_no_constraints"
    r.config_id = 15
    return r


def humanoid_constraint_mode():
    r = rigs.humanoid22()
    r.constraint_mode = True
    r.name = "humanoid_constraint_mode"
    r.config_id = 16
    return r


def humanoid_stabilized():
    """stabilization_passes = 2 (reference src/ik_bone_segment_3d.cpp:163-176): MSD accept / revert on the root segment."""
    r = rigs.humanoid22()
    r.stabilization_passes = 2
    r.name = "humanoid_stabilized"
    r.config_id = 17
    return r


def chain64_stabilized():
    """Root segment of 8 translating bones with a 9-effector list, one stabilisation pass, fewer iterations."""
    r = rigs.chain64()
    r.stabilization_passes = 1
    r.iterations = 6
    r.name = "chain64_stabilized"
    r.config_id = 18
    return r


def star_stabilized_constraint_mode():
    """constraint_mode + stabilisation: no QCP, snaps only, MSD still evaluated against the step's target headings."""
    r = star_mixed_pins()
    r.stabilization_passes = 3
    r.constraint_mode = True
    r.name = "star_stabilized_constraint_mode"
    r.config_id = 19
    return r


def chain_multibone_root_stabilized():
    """Multi-bone translating root segment where stabilisation actually reverts steps."""
    r = chain_multibone_root(12, iterations=5)
    r.stabilization_passes = 2
    r.name = "chain_multibone_root_stabilized"
    r.config_id = 21
    return r


def two_roots_stabilized():
    r = two_roots()
    r.stabilization_passes = 1
    r.name = "two_roots_stabilized"
    r.config_id = 20
    return r


EDGE_RIGS = {f.__name__: f for f in [chain_multibone_root, chain_diverging, two_roots, star_mixed_pins, humanoid_no_constraints, humanoid_constraint_mode,
                                     humanoid_stabilized, chain64_stabilized, star_stabilized_constraint_mode, two_roots_stabilized,
                                     chain_multibone_root_stabilized]}


def perturbed_start_pose(rig, n, seed=7, angle_deg=12.0, offset=0.02):
    """[n, n_bones, 12] start poses = rest pose with small random local rotations/offsets."""
    rng = np.random.default_rng(seed)
    out = np.repeat(rig.rest_local[None], n, axis=0).astype(np.float32).copy()
    for k in range(n):
        for b in range(rig.n_bones):
            ax = rng.normal(size=3)
            R = _axis_angle(ax, rng.uniform(0, angle_deg) * DEG) @ rig.rest_local[b, :9].astype(np.float64).reshape(3, 3)
            out[k, b, :9] = R.reshape(9).astype(np.float32)
            out[k, b, 9:] += rng.uniform(-offset, offset, 3).astype(np.float32)
    return out


def random_rig(seed, n_bones=None, n_pins=None):
    """Random skeleton tree with random pins / kusudama rows / damping: fuzzes the flattener (segment building,
    dropped segments, effector lists, mpf cut-offs, weights) and every kernel stage against the oracle.
    n_bones: override the random size (3..40) -- larger rigs land on the 64-bone-and-up kernel variants; n_pins: override the
    random pin count (1..7) -- many pins on a large tree = many solved bones."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(3, 41))
    if n_bones is not None:
        n = int(n_bones)
    parent = np.full(n, -1, np.int32)
    for b in range(1, n):
        # mostly chains with occasional branching; a second root now and then
        r = rng.random()
        if r < 0.04:
            parent[b] = -1
        elif r < 0.7:
            parent[b] = b - 1
        else:
            parent[b] = int(rng.integers(0, b))
    rest = np.zeros((n, 12))
    for b in range(n):
        ax = rng.normal(size=3)
        ang = rng.uniform(0, 40) * DEG if rng.random() < 0.8 else 0.0
        off = rng.normal(size=3) * 0.15 if parent[b] >= 0 else rng.normal(size=3) * 0.5
        if rng.random() < 0.1:
            off[:] = 0.0  # coincident joints
        rest[b] = _xf(_axis_angle(ax, ang), off)
    r = Rig(f"random{seed}", [f"b{i}" for i in range(n)], parent, rest.astype(np.float32), iterations=int(rng.integers(1, 6)), config_id=100 + seed)
    n_pins_drawn = int(rng.integers(1, min(n, 7) + 1))
    n_pins = n_pins_drawn if n_pins is None else min(int(n_pins), n)
    bones = rng.choice(n, size=n_pins, replace=False)
    for b in bones:
        pr = tuple(float(x) for x in np.where(rng.random(3) < 0.6, rng.uniform(0.05, 0.6, 3), 0.0))
        r.pins.append(dict(bone=int(b), weight=float(rng.choice([0.0, 0.3, 1.0, 0.75])) if rng.random() < 0.9 else 0.0,
                           mpf=float(rng.choice([0.0, 0.5, 1.0, 1.0])), priorities=pr))
    for b in range(n):
        if rng.random() < 0.6:
            nc = int(rng.integers(0, 5))
            cones = []
            for _ in range(nc):
                c = rng.normal(size=3)
                c = c / np.linalg.norm(c) if rng.random() < 0.95 else np.zeros(3)
                cones.append((float(c[0]), float(c[1]), float(c[2]), float(rng.uniform(0.05, 1.4))))
            r.constraints.append(dict(bone=b, twist_from=float(rng.uniform(-2, 2)), twist_range=float(rng.uniform(0.05, 6.0)), cones=cones))
    if rng.random() < 0.5:
        r.bone_damp = rng.uniform(0.01, 0.3, size=int(rng.integers(1, n + 1))).astype(np.float32)
    r.default_damp = float(np.float32(rng.uniform(0.02, 0.5)))
    if rng.random() < 0.3:
        r.stabilization_passes = int(rng.integers(1, 3))
    if rng.random() < 0.1:
        r.constraint_mode = True
    return r


def soak_rig(k):
    """Rig k of the parity soak (profiles/run_fuzz_soak.py): random_rig(1000 + k); every 8th one is a dense large tree (64 ... 520
    bones, one pin per 5 ... 11 bones), every 40th a 600 ... 900-bone one (past 256 solved bones: the unbounded kernel variant)."""
    seed = 1000 + k
    rng = np.random.default_rng(seed)
    if k % 8 == 7:
        nb = int(rng.integers(64, 521)) if k % 40 != 39 else int(rng.integers(600, 901))
        return random_rig(seed, n_bones=nb, n_pins=max(2, nb // int(rng.integers(5, 12))))
    return random_rig(seed)


def star_stabilized(n_arms=40, arm_len=2):
    """A hub with n_arms arms of arm_len bones, every arm tip pinned, stabilisation passes on: one root-segment effector list of
    n_arms entries (the reference has no limit on it; the stabilisation variants keep one pre-step tip origin per entry)."""
    n = 1 + arm_len * n_arms
    parent = np.full(n, -1, np.int32)
    rest = np.zeros((n, 12))
    rest[0] = _xf(np.eye(3), np.zeros(3))
    for a in range(n_arms):
        ang = 2 * np.pi * a / n_arms
        d = np.array([np.cos(ang), 0.3 * np.sin(3 * ang), np.sin(ang)])
        for j in range(arm_len):
            b = 1 + arm_len * a + j
            parent[b] = 0 if j == 0 else b - 1
            rest[b] = _xf(_axis_angle(np.array([0.0, 1.0, 0.0]), ang * 0.1) if j == 0 else np.eye(3), (0.2 + 0.05 * j) * d)
    r = Rig(f"star_stabilized{n_arms}x{arm_len}", [f"b{i}" for i in range(n)], parent, rest.astype(np.float32), iterations=4, config_id=300 + n_arms)
    for a in range(n_arms):
        r.pins.append(dict(bone=arm_len * (a + 1), weight=1.0 if a % 3 else 0.5, mpf=1.0 if a % 2 else 0.6,
                           priorities=(0.2, 0.0, 0.3) if a % 4 == 0 else (0.0, 0.0, 0.0)))
    for a in range(0, n_arms, 3):
        r.constraints.append(dict(bone=1 + arm_len * a, twist_from=0.1, twist_range=2.0, cones=[(0.0, 1.0, 0.0, 0.8)]))
    r.default_damp = float(np.float32(0.2))
    r.stabilization_passes = 2
    return r


def no_pins():
    """No pins at all: nothing is solved; every bone passes through (reference early-out, many_bone_ik_3d.cpp:649)."""
    r = rigs.humanoid22()
    r.pins = []
    r.name = "no_pins"
    r.config_id = 30
    return r


def scaled_bones():
    """Non-unit and negative bone scales in the rest pose (det < 0 flips in get_rotation_quaternion, get_scale sign)."""
    n = 8
    parent = np.arange(-1, n - 1, dtype=np.int32)
    rest = np.zeros((n, 12))
    scales = [(1, 1, 1), (1.5, 1.5, 1.5), (1, 2, 0.5), (-1, 1, 1), (1, 1, 1), (0.7, 0.7, 0.7), (1, -1, 1), (1, 1, 1)]
    for i in range(n):
        R = _axis_angle((0.3, 1.0, 0.2), (7.0 * i) * DEG) @ np.diag(scales[i])
        rest[i] = _xf(R, (0.02 * i, 0.0 if i == 0 else 0.2, 0.0))
    r = Rig("scaled_bones", [f"b{i}" for i in range(n)], parent, rest.astype(np.float32), iterations=4, config_id=31)
    r.pins = [dict(bone=0, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2)), dict(bone=4, weight=0.6, mpf=1.0, priorities=(0.2, 0.1, 0.2)),
              dict(bone=7, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2))]
    _add_constraints(r, {2: (2, 40, 25, -20, 50), 5: (1, 30, 0, -45, 90), 6: (3, 25, 20, -10, 40)})
    return r


EDGE_RIGS.update({f.__name__: f for f in [no_pins, scaled_bones]})


def chain100():
    """100-bone single chain, one pin at the tip: ONE 100-bone translating root segment (largest kernel variant)."""
    n = 100
    parent, rest = _chain(n, seg=0.03, curl=1.0)
    r = Rig("chain100", [f"b{i}" for i in range(n)], parent, rest, iterations=2, config_id=32)
    r.pins = [dict(bone=n - 1, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2))]
    _add_constraints(r, {b: (1 + b % 3, 35, 20, -40, 80) for b in range(1, n, 2)})
    return r


def big_tree120():
    """120-bone branching rig with 12 pins (deep effector walks, many segments): largest kernel variant."""
    rng = np.random.default_rng(77)
    n = 120
    parent = np.full(n, -1, np.int32)
    for b in range(1, n):
        parent[b] = b - 1 if rng.random() < 0.8 else int(rng.integers(max(0, b - 30), b))
    rest = np.zeros((n, 12))
    for b in range(n):
        rest[b] = _xf(_axis_angle(rng.normal(size=3), rng.uniform(0, 25) * DEG), rng.normal(size=3) * 0.08 + (0, 0.1, 0))
    r = Rig("big_tree120", [f"b{i}" for i in range(n)], parent, rest.astype(np.float32), iterations=3, config_id=33)
    for b in sorted(rng.choice(np.arange(5, n), size=12, replace=False)):
        r.pins.append(dict(bone=int(b), weight=float(rng.uniform(0.3, 1.0)), mpf=float(rng.choice([0.5, 1.0])), priorities=(0.2, 0.0, 0.2)))
    _add_constraints(r, {int(b): (int(rng.integers(1, 4)), 40, 25, -30, 60) for b in range(1, n, 3)})
    return r


EDGE_RIGS.update({f.__name__: f for f in [chain100, big_tree120]})


def chain200():
    """200-bone single chain, pins at bones 99 and 199: a 100-bone translating root segment under a 100-bone child
    segment -- needs the {256, 256, 32} kernel variant, and its walk list (quadratic in chain depth, ~160 KB) puts the
    rig constants past the shared-memory budget: tail layout, walk list read from global memory."""
    n = 200
    parent, rest = _chain(n, seg=0.02, curl=0.6)
    r = Rig("chain200", [f"b{i}" for i in range(n)], parent, rest, iterations=2, config_id=34)
    r.pins = [dict(bone=99, weight=0.5, mpf=1.0, priorities=(0.2, 0.0, 0.2)), dict(bone=n - 1, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2))]
    _add_constraints(r, {b: (1 + b % 2, 35, 20, -40, 80) for b in range(1, n, 3)})
    return r


def chain150():
    """150-bone chain, one pin: {256, 256, 32} variant with the ordinary (fully shared-memory resident) layout."""
    n = 150
    parent, rest = _chain(n, seg=0.02, curl=0.8)
    r = Rig("chain150", [f"b{i}" for i in range(n)], parent, rest, iterations=2, config_id=36)
    r.pins = [dict(bone=n - 1, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2))]
    _add_constraints(r, {b: (1, 35, 20, -40, 80) for b in range(1, n, 5)})
    return r


def big_tree240():
    """240-bone branching rig, every leaf pinned (mixed priorities, deep walks, walk stack beyond 16 is allowed): {256, 256, 32}."""
    rng = np.random.default_rng(78)
    n = 240
    parent = np.full(n, -1, np.int32)
    for b in range(1, n):
        parent[b] = b - 1 if rng.random() < 0.85 else int(rng.integers(max(0, b - 60), b))
    rest = np.zeros((n, 12))
    for b in range(n):
        rest[b] = _xf(_axis_angle(rng.normal(size=3), rng.uniform(0, 20) * DEG), rng.normal(size=3) * 0.05 + (0, 0.08, 0))
    r = Rig("big_tree240", [f"b{i}" for i in range(n)], parent, rest.astype(np.float32), iterations=2, config_id=35)
    leaves = [b for b in range(n) if not np.any(parent == b)]
    inner = [int(b) for b in rng.choice(np.arange(5, n), size=6, replace=False) if int(b) not in leaves]
    for b in sorted(set(leaves + inner)):  # every leaf pinned: (almost) every bone is solved
        r.pins.append(dict(bone=int(b), weight=float(rng.uniform(0.3, 1.0)), mpf=float(rng.choice([0.5, 1.0])),
                           priorities=(0.2, float(rng.choice([0.0, 0.1])), 0.2)))
    _add_constraints(r, {int(b): (int(rng.integers(1, 4)), 40, 25, -30, 60) for b in range(1, n, 4)})
    return r


# rigs beyond 128 solved bones: their own dict (not part of the committed golden fixtures); compared live with the oracle
# and, where it travelled, with the reference module's own code
def chain300():
    """300 solved bones in one chain with three pins: beyond the {256, 256, 32} kernel variant -> the unbounded variant."""
    n = 300
    parent = np.arange(-1, n - 1, dtype=np.int32)
    rest = np.zeros((n, 12), np.float32)
    rest[:, 0] = rest[:, 4] = rest[:, 8] = 1.0
    rest[1:, 10] = 0.05
    r = rigs.Rig("chain300", [f"b{i}" for i in range(n)], parent, rest, iterations=2)
    r.pins = [dict(bone=0, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2)), dict(bone=140, weight=0.5, mpf=0.5, priorities=(0.2, 0.0, 0.2)),
              dict(bone=n - 1, weight=1.0, mpf=1.0, priorities=(0.2, 0.1, 0.2))]
    rigs._add_constraints(r, {b: (1 + b % 3, 20, 0, -20, 40) for b in range(10, n, 10)})
    return r


LARGE_RIGS = {f.__name__: f for f in [chain150, chain200, big_tree240]}
UNBOUNDED_RIGS = {f.__name__: f for f in [chain300]}
