"""A/B of experimental libmbik builds (profiles/build_variant.py): for every many_bone_ik_b200/_variants/libmbik_<tag>.so (or the
tags given) run device-resident launches of one rig in a subprocess (own `timeout`: a hung kernel must not take the box
down), report ms per launch and whether the outputs are bit-identical to the stock library's.

    python profiles/run_variants.py [--rig humanoid22] [--poses 1048576] [--launches 5] [tags ...]"""
import argparse
import glob
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--rig", default="humanoid22")
ap.add_argument("--poses", type=int, default=1048576)
ap.add_argument("--launches", type=int, default=5)
ap.add_argument("--child", action="store_true")
ap.add_argument("--timeout", type=int, default=240)
ap.add_argument("tags", nargs="*")
a = ap.parse_args()

if a.child:
    sys.path.insert(0, ROOT)
    import torch
    from many_bone_ik_b200 import BatchedIKRig, rigs
    from many_bone_ik_b200._capi import MBIK_IO_DEVICE
    if a.rig in rigs.RIGS:
        rig = rigs.RIGS[a.rig]()
    else:  # the large / edge rigs of the tests
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import rig_cases
        rig = getattr(rig_cases, a.rig)()
    R = BatchedIKRig(rig)
    T = torch.from_numpy(rigs.random_targets(rig, 0, a.poses)).cuda()
    O = torch.empty((a.poses, rig.n_bones, 10), dtype=torch.float32, device="cuda")
    ms = []
    for i in range(a.launches):
        R.solve_raw(a.poses, T, O, device=0, flags=MBIK_IO_DEVICE, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ms.append(R.last_kernel_ms(0))
    h = hashlib.sha256(O.cpu().numpy().tobytes()).hexdigest()
    best = min(ms[1:]) if len(ms) > 1 else ms[0]
    print(json.dumps({"ms": ms, "best_ms": best, "M_solves_per_s": a.poses / best / 1e3, "sha256": h}))
    sys.exit(0)

vdir = os.path.join(ROOT, "many_bone_ik_b200", "_variants")
tags = a.tags or sorted(os.path.basename(p)[len("libmbik_"):-3] for p in glob.glob(os.path.join(vdir, "libmbik_*.so")))
results = {}
ref_hash = None
for tag in ["stock"] + tags:
    env = dict(os.environ)
    if tag != "stock":
        env["MBIK_LIB"] = os.path.join(vdir, f"libmbik_{tag}.so")
    cmd = ["timeout", str(a.timeout), sys.executable, os.path.abspath(__file__), "--child", "--rig", a.rig, "--poses", str(a.poses), "--launches", str(a.launches)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        results[tag] = {"error": r.returncode, "stderr": r.stderr[-400:]}
        print(tag, "FAILED rc", r.returncode, r.stderr[-300:], flush=True)
        continue
    d = json.loads(r.stdout.strip().splitlines()[-1])
    if tag == "stock":
        ref_hash = d["sha256"]
    d["bit_identical_to_stock"] = d["sha256"] == ref_hash
    results[tag] = d
    print(f"{a.rig:11s} {tag:14s} {d['best_ms']:9.3f} ms  {d['M_solves_per_s']:8.3f} M solves/s  identical={d['bit_identical_to_stock']}", flush=True)
print(json.dumps({"rig": a.rig, "poses": a.poses, "results": results}))
