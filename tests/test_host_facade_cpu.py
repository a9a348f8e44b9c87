"""C++ host facade (many_bone_ik_b200/host/many_bone_ik_host.hpp): the reference-named ManyBoneIK3D front end, its
property paths (`_set`/`_get`) and the .tscn-style loader.  CPU part: the facade builds the same rig as the direct
C-ABI path for every test rig.  The solve through the facade is covered by tests/test_parity_gpu.py."""
import os
import subprocess

import numpy as np
import pytest

import rig_cases
from many_bone_ik_b200 import BatchedIKRig, rigs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL_RIGS = dict(rigs.RIGS)
ALL_RIGS.update(rig_cases.EDGE_RIGS)


def build_driver(tmp):
    exe = os.path.join(str(tmp), "facade_driver")
    pkg = os.path.join(ROOT, "many_bone_ik_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "cpp", "facade_driver.cpp"),
                           "-L" + pkg, "-l:libmbik.so", "-Wl,-rpath," + pkg, "-ldl", "-lpthread", "-lrt"])
    return exe


def run_driver(exe, tmp, rig, extra=()):
    sk, pr = os.path.join(str(tmp), rig.name + ".skel"), os.path.join(str(tmp), rig.name + ".props")
    open(sk, "w").write(rigs.to_skeleton_text(rig))
    open(pr, "w").write(rigs.to_property_text(rig))
    r = subprocess.run([exe, sk, pr] + list(extra), capture_output=True, text=True)
    facts = {}
    for ln in r.stdout.splitlines():
        k, _, v = ln.partition(":")
        facts[k.strip()] = v.strip()
    return r.returncode, facts, r.stdout + r.stderr


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    return build_driver(tmp_path_factory.mktemp("facade"))


@pytest.mark.parametrize("name", sorted(ALL_RIGS))
def test_facade_builds_the_same_rig(driver, tmp_path, name):
    rig = ALL_RIGS[name]()
    rc, facts, log = run_driver(driver, tmp_path, rig)
    assert rc == 0, log
    R = BatchedIKRig(rig)
    assert int(facts["pin_count"]) == len(rig.pins) and int(facts["constraint_count"]) == len(rig.constraints)
    for k in ("n_bones", "n_solved", "n_segments", "n_effectors", "max_headings", "n_cones", "iterations"):
        assert int(facts[k]) == R.info[k], (k, facts[k], R.info[k])
    got = [int(x) for x in facts["bone_list"].split()]
    assert got == list(R.bone_order())
    if rig.pins:
        assert np.float32(float(facts["pins/0/weight"])) == np.float32(rig.pins[0]["weight"])
    if rig.constraints and rig.constraints[0]["cones"]:
        assert np.float32(float(facts["constraints/0/kusudama_open_cone/0/radius"])) == np.float32(rig.constraints[0]["cones"][0][3])


def test_facade_solve_without_gpu_reports_no_device(driver, tmp_path):
    from many_bone_ik_b200 import device_count
    if device_count() > 0:
        pytest.skip("a CUDA device is present")
    rig = rigs.humanoid22()
    T = rigs.random_targets(rig, 0, 8)
    tb = os.path.join(str(tmp_path), "t.bin")
    T.tofile(tb)
    rc, facts, log = run_driver(driver, tmp_path, rig, [tb, os.path.join(str(tmp_path), "o.bin"), "8"])
    assert rc == 3 and int(facts["solve"]) in (-4, -2), log
    assert "no CUDA device" in log or "CUDA" in log
