"""Canonical benchmark rigs and seeded synthetic targets.

A rig is everything ``ManyBoneIK3D::_bone_list_changed`` reads (reference
src/many_bone_ik_3d.cpp:1011-1068): skeleton topology + rest pose, the pins table
(IKEffectorTemplate3D rows, src/ik_effector_template_3d.h:40-45) and the constraint tables
(constraint_names / joint_twist / kusudama_open_cones, src/many_bone_ik_3d.h:53-59).

Generators are pure numpy so that any shard can regenerate its slice of a batch independently.
Nothing here touches the solver; it only produces inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DEG = np.pi / 180.0


@dataclass
class Rig:
    name: str
    bone_names: list
    parent: np.ndarray  # int32 [n_bones]
    rest_local: np.ndarray  # float32 [n_bones, 12]  basis rows + origin
    pins: list = field(default_factory=list)  # dicts: bone, weight, mpf, priorities
    constraints: list = field(default_factory=list)  # dicts: bone, twist_from, twist_range, cones [(cx,cy,cz,r)]
    bone_damp: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    default_damp: float = float(np.float32(5.0) * np.float32(np.pi / 180.0))
    iterations: int = 15
    stabilization_passes: int = 0
    constraint_mode: bool = False
    config_id: int = 0

    @property
    def n_bones(self):
        return int(self.parent.shape[0])

    @property
    def n_pins(self):
        return len(self.pins)


# ---------------------------------------------------------------------------------------------
# small float64 math helpers (generation only)
# ---------------------------------------------------------------------------------------------
def _axis_angle(axis, angle):
    axis = np.asarray(axis, np.float64)
    n = np.linalg.norm(axis)
    if n == 0 or angle == 0:
        return np.eye(3)
    a = axis / n
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def _xf(R, t):
    out = np.zeros(12, np.float64)
    out[:9] = np.asarray(R, np.float64).reshape(9)
    out[9:] = t
    return out


def global_rest(parent, rest_local):
    """FK of the rest pose in float64: returns (R[n,3,3], t[n,3])."""
    n = len(parent)
    R = np.zeros((n, 3, 3))
    t = np.zeros((n, 3))
    for b in range(n):
        Rl = rest_local[b, :9].astype(np.float64).reshape(3, 3)
        tl = rest_local[b, 9:].astype(np.float64)
        p = parent[b]
        if p < 0:
            R[b], t[b] = Rl, tl
        else:
            assert p < b, "bones must be topologically ordered (parent index < child index)"
            R[b] = R[p] @ Rl
            t[b] = R[p] @ tl + t[p]
    return R, t


def _children(parent):
    ch = [[] for _ in parent]
    for b, p in enumerate(parent):
        if p >= 0:
            ch[p].append(b)
    return ch


def bone_direction_globals(rig_or_parent, rest_local=None, pinned=None):
    """Approximate (float64) global +Y axis of every bone's bone-direction frame after
    IKBone3D::update_default_bone_direction_transform (reference src/ik_bone_3d.cpp:57-93),
    including its argument-order quirk.  Used only to place cones and
    targets sensibly; the solver computes its own (float32) version."""
    if isinstance(rig_or_parent, Rig):
        parent, rest_local = rig_or_parent.parent, rig_or_parent.rest_local
        pinned = {p["bone"] for p in rig_or_parent.pins}
    else:
        parent = rig_or_parent
    n = len(parent)
    R, t = global_rest(parent, rest_local)
    ch = _children(parent)
    has_pin_below = np.zeros(n, bool)
    for b in reversed(range(n)):
        has_pin_below[b] = (b in pinned) or any(has_pin_below[c] for c in ch[b])
    ydir = np.zeros((n, 3))
    for b in range(n):
        y = R[b][:, 1]
        y = y / np.linalg.norm(y)
        ydir[b] = y
        if not ch[b]:
            continue
        is_tip = len(ch[b]) > 1 or (b in pinned)
        if is_tip:
            offs = [(t[c] - t[b]) if has_pin_below[c] else np.zeros(3) for c in ch[b]]
        else:
            offs = [t[ch[b][0]] - t[b]]
        c = np.mean(offs, axis=0)
        if np.dot(c, c) < 1e-5:
            p = parent[b]
            c = (R[p][:, 1] if p >= 0 else R[b][:, 1]).copy()
        if np.dot(c, c) < 1e-5:
            continue
        c = c / np.linalg.norm(c)
        ydir[b] = 2.0 * np.dot(c, y) * y - c  # R(c->y) applied to y
    return R, t, ydir


def _cone_centres_around(d, n_cones, spread):
    """n_cones unit vectors: the first is d, the others tilt away from d by `spread` radians in
    different azimuths (gives the tangent-path code something to do)."""
    d = d / np.linalg.norm(d)
    ref = np.array([0.0, 0.0, 1.0]) if abs(d[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
    u = np.cross(d, ref)
    u /= np.linalg.norm(u)
    out = [d]
    for k in range(1, n_cones):
        az = 2.4 * k
        axis = np.cos(az) * u + np.sin(az) * np.cross(d, u)
        out.append(_axis_angle(axis, spread) @ d)
    return out


def _add_constraints(rig: Rig, spec):
    """spec: {bone_index: (n_cones, radius_deg, spread_deg, twist_from_deg, twist_range_deg)}.
    Cone centres are expressed in the constraint-orientation frame = the parent's aligned frame."""
    R, _, ydir = bone_direction_globals(rig)
    for b, (nc, rad, spread, tf, tr) in spec.items():
        p = rig.parent[b]
        Rp = R[p] if p >= 0 else np.eye(3)
        d_local = Rp.T @ ydir[b]
        cones = []
        for k, c in enumerate(_cone_centres_around(d_local, nc, spread * DEG)):
            r = rad * DEG * (1.0 if k == 0 else 0.75)
            cones.append((float(c[0]), float(c[1]), float(c[2]), float(r)))
        rig.constraints.append(dict(bone=int(b), twist_from=float(tf * DEG), twist_range=float(tr * DEG), cones=cones))


# ---------------------------------------------------------------------------------------------
# humanoid22 (configs 1-3)
# ---------------------------------------------------------------------------------------------
def humanoid22() -> Rig:
    names = ["Hips", "Spine", "Chest", "UpperChest", "Neck", "Head",
             "LShoulder", "LUpperArm", "LLowerArm", "LHand",
             "RShoulder", "RUpperArm", "RLowerArm", "RHand",
             "LUpperLeg", "LLowerLeg", "LFoot", "LToes",
             "RUpperLeg", "RLowerLeg", "RFoot", "RToes"]
    idx = {n: i for i, n in enumerate(names)}
    par = {"Hips": None, "Spine": "Hips", "Chest": "Spine", "UpperChest": "Chest", "Neck": "UpperChest", "Head": "Neck",
           "LShoulder": "UpperChest", "LUpperArm": "LShoulder", "LLowerArm": "LUpperArm", "LHand": "LLowerArm",
           "RShoulder": "UpperChest", "RUpperArm": "RShoulder", "RLowerArm": "RUpperArm", "RHand": "RLowerArm",
           "LUpperLeg": "Hips", "LLowerLeg": "LUpperLeg", "LFoot": "LLowerLeg", "LToes": "LFoot",
           "RUpperLeg": "Hips", "RLowerLeg": "RUpperLeg", "RFoot": "RLowerLeg", "RToes": "RFoot"}
    # T-pose, 1.7 m tall; offsets in the parent's frame; a few non-identity rest rotations so no
    # basis is trivially the identity
    off = {"Hips": (0, 0.95, 0), "Spine": (0, 0.10, 0), "Chest": (0, 0.12, 0.01), "UpperChest": (0, 0.12, 0),
           "Neck": (0, 0.15, 0.01), "Head": (0, 0.10, 0),
           "LShoulder": (0.05, 0.10, 0), "LUpperArm": (0.12, 0, 0), "LLowerArm": (0.28, 0, 0), "LHand": (0.26, 0, 0),
           "RShoulder": (-0.05, 0.10, 0), "RUpperArm": (-0.12, 0, 0), "RLowerArm": (-0.28, 0, 0), "RHand": (-0.26, 0, 0),
           "LUpperLeg": (0.09, -0.05, 0), "LLowerLeg": (0, -0.42, 0), "LFoot": (0, -0.42, 0), "LToes": (0, -0.06, 0.12),
           "RUpperLeg": (-0.09, -0.05, 0), "RLowerLeg": (0, -0.42, 0), "RFoot": (0, -0.42, 0), "RToes": (0, -0.06, 0.12)}
    rot = {"Spine": ((1, 0, 0), 3), "Chest": ((1, 0, 0), -2), "Neck": ((1, 0, 0), 8), "Head": ((1, 0, 0), -6),
           "LShoulder": ((0, 0, 1), 4), "RShoulder": ((0, 0, 1), -4), "LUpperArm": ((0, 1, 0), 5), "RUpperArm": ((0, 1, 0), -5),
           "LLowerArm": ((0, 1, 0), -8), "RLowerArm": ((0, 1, 0), 8), "LUpperLeg": ((1, 0, 0), -4), "RUpperLeg": ((1, 0, 0), -4),
           "LLowerLeg": ((1, 0, 0), 7), "RLowerLeg": ((1, 0, 0), 7), "LFoot": ((1, 0, 0), -3), "RFoot": ((1, 0, 0), -3)}
    n = len(names)
    parent = np.array([-1 if par[nm] is None else idx[par[nm]] for nm in names], np.int32)
    rest = np.zeros((n, 12), np.float64)
    for i, nm in enumerate(names):
        ax, deg = rot.get(nm, ((0, 1, 0), 0))
        rest[i] = _xf(_axis_angle(ax, deg * DEG), off[nm])
    rig = Rig("humanoid22", names, parent, rest.astype(np.float32), iterations=10, config_id=1)
    for nm in ["Head", "LHand", "RHand", "LFoot", "RFoot"]:
        rig.pins.append(dict(bone=idx[nm], weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2)))
    spec = {}
    for nm in ["LUpperArm", "RUpperArm", "LUpperLeg", "RUpperLeg"]:
        spec[idx[nm]] = (3, 60, 40, -45, 90)
    for nm in ["Spine", "Chest", "UpperChest", "Neck", "Head"]:
        spec[idx[nm]] = (2, 30, 20, -30, 60)
    for nm, rad in [("LShoulder", 25), ("RShoulder", 25), ("LLowerArm", 50), ("RLowerArm", 50), ("LHand", 40), ("RHand", 40),
                    ("LLowerLeg", 50), ("RLowerLeg", 50), ("LFoot", 35), ("RFoot", 35)]:
        spec[idx[nm]] = (1, rad, 0, -45, 90)
    _add_constraints(rig, dict(sorted(spec.items())))
    return rig


# ---------------------------------------------------------------------------------------------
# chain64 (config 4)
# ---------------------------------------------------------------------------------------------
def chain64() -> Rig:
    n = 64
    names = [f"Seg{i:02d}" for i in range(n)]
    parent = np.arange(-1, n - 1, dtype=np.int32)
    rest = np.zeros((n, 12), np.float64)
    for i in range(n):
        # gentle helical rest curl so neither positions nor bases are degenerate
        R = _axis_angle((0.3, 0.1, 1.0), (1.5 if i else 0.0) * DEG)
        rest[i] = _xf(R, (0.0, 0.05 if i else 0.0, 0.0))
    rig = Rig("chain64", names, parent, rest.astype(np.float32), iterations=30, config_id=4)
    # The root is pinned so that the translating root segment is the single bone Seg00: with the reference's
    # heading asymmetry a multi-bone translating root segment diverges to NaN
    # within ~7 iterations, which would make the config meaningless (tests/ keep such a rig as an edge case).
    inter = [0, 8, 16, 24, 32, 40, 48, 56]
    for k, b in enumerate(inter):
        rig.pins.append(dict(bone=b, weight=float(0.25 + 0.5 * k / (len(inter) - 1)), mpf=1.0, priorities=(0.2, 0.0, 0.2)))
    rig.pins.append(dict(bone=63, weight=1.0, mpf=1.0, priorities=(0.2, 0.0, 0.2)))
    # Kusudama on every 4th bone only: in float32 the reference's twist snap (which folds the parent's
    # accumulated non-orthonormality into the child's local basis, src/ik_kusudama_3d.cpp:130) is unstable on
    # long runs of consecutively constrained bones -- a fully constrained 64-chain is all-NaN by iteration 7.
    spec = {b: (1, 15, 0, -20, 40) for b in range(4, n, 4)}
    _add_constraints(rig, spec)
    # "stiffness+damping": stiffness is inert in the reference; the effective damp is
    # min(bone_damp[bone_id], default_damp) (reference src/ik_bone_segment_3d.cpp:229-237)
    rig.bone_damp = np.array([(2.0 + 3.0 * ((i * 7) % 11) / 10.0) * DEG for i in range(n)], np.float32)
    return rig


# ---------------------------------------------------------------------------------------------
# quad80 (config 5)
# ---------------------------------------------------------------------------------------------
def quad80() -> Rig:
    names, par, off = [], [], []

    def add(name, parent_name, offset):
        names.append(name)
        par.append(parent_name)
        off.append(offset)

    add("Pelvis", None, (0, 0.9, -0.6))
    for i in range(1, 9):
        add(f"Spine{i}", "Pelvis" if i == 1 else f"Spine{i-1}", (0, 0.01, 0.15))
    for i in range(1, 8):
        add(f"Neck{i}", "Spine8" if i == 1 else f"Neck{i-1}", (0, 0.07, 0.05))
    add("Head", "Neck7", (0, 0.06, 0.08))
    add("Jaw", "Head", (0, -0.05, 0.10))
    for s in "LR":
        add(f"Ear{s}1", "Head", (0.05 if s == "L" else -0.05, 0.08, 0))
        add(f"Ear{s}2", f"Ear{s}1", (0, 0.06, 0))
    for s in "LR":
        sx = 1 if s == "L" else -1
        chain = [("Scapula", (0.12 * sx, -0.05, 0.0)), ("Humerus", (0.02 * sx, -0.22, 0.04)), ("Radius", (0, -0.25, -0.03)),
                 ("Carpus", (0, -0.20, 0.0)), ("Metacarpus", (0, -0.08, 0.01)), ("Paw", (0, -0.05, 0.03))]
        prev = "Spine8"
        for nm, o in chain:
            add(f"Fore{s}{nm}", prev, o)
            prev = f"Fore{s}{nm}"
        for k in range(3):
            add(f"Fore{s}Toe{k}", prev, (0.02 * (k - 1), -0.02, 0.04))
    for s in "LR":
        sx = 1 if s == "L" else -1
        chain = [("Femur", (0.11 * sx, -0.06, 0.0)), ("Tibia", (0, -0.28, 0.06)), ("Tarsus", (0, -0.27, -0.08)),
                 ("Metatarsus", (0, -0.18, 0.02)), ("Paw", (0, -0.06, 0.03))]
        prev = "Pelvis"
        for nm, o in chain:
            add(f"Hind{s}{nm}", prev, o)
            prev = f"Hind{s}{nm}"
        for k in range(3):
            add(f"Hind{s}Toe{k}", prev, (0.02 * (k - 1), -0.02, 0.04))
    for i in range(24):
        add(f"Tail{i}", "Pelvis" if i == 0 else f"Tail{i-1}", (0, 0.005, -0.05))
    assert len(names) == 80, len(names)
    idx = {n: i for i, n in enumerate(names)}
    parent = np.array([-1 if p is None else idx[p] for p in par], np.int32)
    rest = np.zeros((80, 12), np.float64)
    for i, nm in enumerate(names):
        ang = ((i * 37) % 13 - 6) * 1.0  # deterministic few-degree rest rotations
        ax = ((i % 3 == 0) * 1.0, (i % 3 == 1) * 1.0, (i % 3 == 2) * 1.0)
        rest[i] = _xf(_axis_angle(ax, ang * DEG if i else 0.0), off[i])
    rig = Rig("quad80", names, parent, rest.astype(np.float32), iterations=15, config_id=5)
    pr = (0.2, 0.0, 0.2)
    rig.pins = [
        dict(bone=idx["Pelvis"], weight=1.0, mpf=1.0, priorities=pr),
        dict(bone=idx["Spine8"], weight=0.5, mpf=0.5, priorities=pr),
        dict(bone=idx["Head"], weight=1.0, mpf=1.0, priorities=(0.3, 0.2, 0.3)),  # 3 priority axes: 7 headings
        dict(bone=idx["ForeLPaw"], weight=1.0, mpf=1.0, priorities=pr),
        dict(bone=idx["ForeRPaw"], weight=1.0, mpf=1.0, priorities=pr),
        dict(bone=idx["HindLPaw"], weight=1.0, mpf=1.0, priorities=pr),
        dict(bone=idx["HindRPaw"], weight=1.0, mpf=1.0, priorities=(0.0, 0.0, 0.0)),  # translation-only pin: 1 heading
        dict(bone=idx["Tail11"], weight=0.3, mpf=0.0, priorities=pr),  # mpf = 0: cuts Tail12..23 off from ancestors
        dict(bone=idx["Tail23"], weight=1.0, mpf=1.0, priorities=pr),
    ]
    spec = {}
    solved_candidates = [n for n in names if not (n.startswith("Ear") or n == "Jaw" or "Toe" in n)]
    for nm in solved_candidates:
        b = idx[nm]
        if parent[b] < 0:
            continue
        # long chains are constrained sparsely (see chain64): spine/neck every 2nd bone, tail every 4th
        if (nm.startswith("Spine") or nm.startswith("Neck")) and int(nm[5 if nm.startswith("Spine") else 4:]) % 2 == 1:
            continue
        if nm.startswith("Tail") and int(nm[4:]) % 4 != 3:
            continue
        k = (b * 5) % 7
        ncones = [1, 2, 0, 3, 1, 4, 2][k]  # mixed cone counts 0-4
        spec[b] = (ncones, 35 + 5 * (b % 4), 25, -40 + (b % 3) * 10, 80)
    _add_constraints(rig, spec)
    return rig


RIGS = {"humanoid22": humanoid22, "chain64": chain64, "quad80": quad80}


# ---------------------------------------------------------------------------------------------
# seeded synthetic targets (counter based: any shard regenerates its slice independently)
# ---------------------------------------------------------------------------------------------
def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform(seed, counters, stream):
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03))
        z = _splitmix64(counters.astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D) + key)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def effector_rest_frames(rig: Rig):
    """float64 global rest frame per pin: effector bone's aligned basis, bone origin."""
    R, t = global_rest(rig.parent, rig.rest_local)
    return [(R[p["bone"]], t[p["bone"]]) for p in rig.pins]


def random_targets(rig: Rig, first_pose: int, n_poses: int, pos_range=0.3, max_angle_deg=45.0, scale=1.0):
    """[n_poses, n_pins, 12] float32 targets = effector rest frame o random perturbation
    (origin offset U[-pos_range,pos_range]^3 * scale, rotation about a uniform axis by U(0,max_angle))."""
    seed = 0x5EED0000 + rig.config_id
    E = rig.n_pins
    frames = effector_rest_frames(rig)
    pose = np.arange(first_pose, first_pose + n_poses, dtype=np.uint64)
    out = np.zeros((n_poses, E, 12), np.float32)
    for e in range(E):
        ctr = pose * np.uint64(E) + np.uint64(e)
        u = [_uniform(seed, ctr, s) for s in range(7)]
        dpos = (np.stack(u[0:3], 1) * 2.0 - 1.0) * pos_range * scale
        zc = u[3] * 2.0 - 1.0
        ph = u[4] * 2.0 * np.pi
        rxy = np.sqrt(np.maximum(0.0, 1.0 - zc * zc))
        axis = np.stack([rxy * np.cos(ph), rxy * np.sin(ph), zc], 1)
        ang = u[5] * max_angle_deg * DEG
        c, s = np.cos(ang)[:, None, None], np.sin(ang)[:, None, None]
        K = np.zeros((n_poses, 3, 3))
        K[:, 0, 1], K[:, 0, 2] = -axis[:, 2], axis[:, 1]
        K[:, 1, 0], K[:, 1, 2] = axis[:, 2], -axis[:, 0]
        K[:, 2, 0], K[:, 2, 1] = -axis[:, 1], axis[:, 0]
        Rp = np.eye(3)[None] + s * K + (1 - c) * (K @ K)
        R0, t0 = frames[e]
        Rt = Rp @ R0[None]
        out[:, e, :9] = Rt.reshape(n_poses, 9).astype(np.float32)
        out[:, e, 9:] = (t0[None] + dpos).astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------------
# text forms consumed by the C++ host facade (many_bone_ik_b200/host/many_bone_ik_host.hpp)
# ---------------------------------------------------------------------------------------------
def _f(x):
    return repr(float(np.float32(x)))  # shortest decimal that round-trips the float32 value via double


def to_skeleton_text(rig: Rig) -> str:
    """`n_bones`, then one line per bone: name parent 12 floats (Transform3D memory layout)."""
    lines = [str(rig.n_bones)]
    for b in range(rig.n_bones):
        lines.append(" ".join([rig.bone_names[b], str(int(rig.parent[b]))] + [_f(v) for v in rig.rest_local[b]]))
    return "\n".join(lines) + "\n"


def to_property_text(rig: Rig) -> str:
    """The rig's tables as the `key = value` property lines a Godot .tscn stores for a ManyBoneIK3D node
    (property names of ManyBoneIK3D::_set, reference src/many_bone_ik_3d.cpp:296-375)."""
    out = ['[node name="ManyBoneIK3D" type="ManyBoneIK3D" parent="."]',
           f"iterations_per_frame = {float(rig.iterations)}", f"default_damp = {_f(rig.default_damp)}",
           f"constraint_mode = {'true' if rig.constraint_mode else 'false'}", f"stabilization_passes = {int(rig.stabilization_passes)}",
           f"pin_count = {len(rig.pins)}"]
    for i, p in enumerate(rig.pins):
        name = rig.bone_names[p["bone"]] if 0 <= p["bone"] < rig.n_bones else "NoSuchBone"
        out += [f'pins/{i}/bone_name = &"{name}"', f'pins/{i}/target_node = NodePath("../Targets/{name}")',
                f"pins/{i}/motion_propagation_factor = {_f(p['mpf'])}", f"pins/{i}/weight = {_f(p['weight'])}",
                "pins/%d/direction_priorities = Vector3(%s, %s, %s)" % ((i,) + tuple(_f(v) for v in p["priorities"]))]
    out.append(f"constraint_count = {len(rig.constraints)}")
    for i, c in enumerate(rig.constraints):
        out += [f'constraints/{i}/bone_name = &"{rig.bone_names[c["bone"]]}"', f"constraints/{i}/twist_from = {_f(c['twist_from'])}",
                f"constraints/{i}/twist_range = {_f(c['twist_range'])}", f"constraints/{i}/kusudama_open_cone_count = {len(c['cones'])}"]
        for j, (cx, cy, cz, r) in enumerate(c["cones"]):
            out += [f"constraints/{i}/kusudama_open_cone/{j}/center = Vector3({_f(cx)}, {_f(cy)}, {_f(cz)})",
                    f"constraints/{i}/kusudama_open_cone/{j}/radius = {_f(r)}"]
        out.append(f"constraints/{i}/kusudama_twist = Transform3D(1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0)")
    if len(rig.bone_damp):
        out.append(f"bone_count = {len(rig.bone_damp)}")
        out += [f"bone_damp/{i} = {_f(v)}" for i, v in enumerate(rig.bone_damp)]
    return "\n".join(out) + "\n"
