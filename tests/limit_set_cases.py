"""Alternative fills of a rig's constraint tables (same rows, other values) for the limit-set tests."""
import copy

import numpy as np


def variants(rig, n_sets=4, seed=5):
    """[n_sets] constraint lists: set 0 = the rig's own; the others scale radii / twist ranges and tilt cone centres."""
    rng = np.random.default_rng(seed)
    sets = [copy.deepcopy(rig.constraints)]
    for s in range(1, n_sets):
        cs = copy.deepcopy(rig.constraints)
        for c in cs:
            c["twist_from"] = float(np.float32(c["twist_from"] + rng.uniform(-0.4, 0.4)))
            c["twist_range"] = float(np.float32(max(0.05, c["twist_range"] * rng.uniform(0.4, 1.6))))
            cones = []
            for (cx, cy, cz, r) in c["cones"]:
                d = np.array([cx, cy, cz], np.float64) + rng.normal(size=3) * 0.15
                cones.append((float(np.float32(d[0])), float(np.float32(d[1])), float(np.float32(d[2])), float(np.float32(max(0.02, r * rng.uniform(0.5, 1.5))))))
            c["cones"] = cones
        sets.append(cs)
    return sets


def rig_with(rig, constraints):
    r = copy.deepcopy(rig)
    r.constraints = constraints
    return r
