"""Residency sweep of the thread-per-pose kernel on the large rigs: CTA size (MBIK_THREADS) x CTAs per SM (capped through
MBIK_SMEM_FLOOR) -> ms per launch.  Each configuration runs in its own process (the knobs are read once).
    python profiles/run_residency.py chain64 75776"""
import os
import subprocess
import sys

rig, poses = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
for threads, floor_kb in ((512, 0), (128, 0), (128, 57), (128, 76), (128, 114), (32, 0), (32, 28), (32, 45), (32, 57), (32, 76)):
    env = dict(os.environ, MBIK_THREADS=str(threads), MBIK_SMEM_FLOOR=str(floor_kb * 1024), MBIK_SCHED="throughput")
    r = subprocess.run([sys.executable, os.path.join(here, "run_kernel.py"), "--rig", rig, "--poses", poses, "--launches", "3"], env=env, capture_output=True, text=True)
    last = [ln for ln in r.stdout.splitlines() if ln.startswith("launch")]
    per_sm = "all that fit" if floor_kb == 0 else f"<= {227 // floor_kb} CTAs/SM"
    print(f"{rig} threads/CTA={threads:4d} smem floor={floor_kb:3d} KiB ({per_sm}): {last[-1] if last else r.stderr[-300:]}", flush=True)
