#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched ManyBoneIK solve loop (see the contract in DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], which also covers N=1): the `humanoid22` rig (22 bones, 5 pinned effectors,
kusudama open-cone + twist limits on the 19 non-root solved bones, 10 iterations), 2^20 independent poses PER GPU
with seeded random effector targets; rank g of N solves the contiguous pose range [g*2^20, (g+1)*2^20) of an
N*2^20-pose batch (weak scaling, no collective in the solve).  One "step" = one pass of the hot path over the
rank's shard.

  value : skeleton-solves/s with inputs already resident in HBM (device-resident I/O, CUDA-event timed)
  e2e   : the same through the C ABI with pinned HOST buffers, H2D + D2H inside the timed region
  roofline / cpu_baseline / latency_p50_ms_4096 : see DESIGN.md "Measurement"

`--impl reference` times the reference's own CPU implementation (oracle/_ref: /root/reference/src compiled unmodified
over the engine stand-in oracle/godot_shim; falls back to the line-cited restatement under oracle/ when oracle/_ref is
absent) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POSES_PER_GPU = 1 << 20
LATENCY_BATCH = 4096
METRIC = "ik_skeleton_solves_per_sec"
UNIT = "solves/s"


def workload_config(n_gpus, per_gpu=POSES_PER_GPU):
    return {
        "workload": "humanoid22 (22 bones, 5 effectors, 32 kusudama cones + twist, 10 iterations), 2^20 random-target poses per GPU "
                    "(BASELINE configs[2]'s 1M-pose batch on every GPU: weak scaling)",
        "rig": "humanoid22", "iterations": 10, "total_poses": per_gpu * n_gpus, "poses_per_gpu": per_gpu,
        "parallelism": f"dp{n_gpus} (contiguous pose shards, no collective)",
        "cache": "per GPU the inputs (252 MB targets) and outputs (923 MB) are larger than the 126 MB L2",
    }


# --------------------------------------------------------------------------------------------------
# CPU reference timing -- the ONLY place bench.py touches oracle/
#   kind "reference": oracle/_ref/libmbik_ref.so = the reference module's own sources compiled unmodified over the
#                     engine stand-in (built in the build container from /root/reference, travels prebuilt);
#   kind "port"     : the line-cited restatement oracle/ewbik_oracle.cpp (bit-identical to the above, leaner
#                     containers), used when oracle/_ref is absent and reported next to it otherwise.
# --------------------------------------------------------------------------------------------------
def _time_cpu_solver(solve, cores, rig, budget_s, steps, warmup):
    from many_bone_ik_b200 import rigs
    probe = max(cores * 8, 64)
    T = rigs.random_targets(rig, 0, probe)
    t0 = time.perf_counter()
    solve(rig, T, threads=cores)
    rate = probe / max(time.perf_counter() - t0, 1e-6)
    n = int(max(probe, min(rate * budget_s / max(steps + warmup, 1), 1 << 18)))
    T = rigs.random_targets(rig, 0, n)
    for _ in range(warmup):
        solve(rig, T, threads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        solve(rig, T, threads=cores)
    dt = time.perf_counter() - t0
    # single-thread figure (SURVEY 8(d) CPU reference timing (i)): one pose stream on one core, ~2 s
    n1 = int(max(64, min(rate / max(cores, 1) * 2.0, 1 << 14)))
    t1 = time.perf_counter()
    solve(rig, T[:n1], threads=1)
    dt1 = time.perf_counter() - t1
    return n, dt, n1, dt1


def _host_supports_x86_64_v3():
    try:
        flags = set()
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                flags = set(ln.split(":", 1)[1].split())
                break
        return {"avx2", "bmi2", "fma", "movbe", "f16c"} <= flags
    except Exception:
        return False


def time_cpu_reference(rig, budget_s=12.0, steps=1, warmup=0, with_port=True):
    from oracle import oracle_py as O
    from oracle import reference_py as Rf
    O.build()
    cores = O.hardware_threads()
    have_ref = os.path.exists(Rf.LIB) or Rf.available()
    port = None
    if with_port or not have_ref:
        port_budget = budget_s * (0.35 if have_ref else 1.0)
        n, dt, n1, dt1 = _time_cpu_solver(O.solve_batch, cores, rig, port_budget, steps, warmup)
        port = {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n} poses x {steps} pass(es) of the same workload, oracle restatement, all {cores} host threads, {dt:.1f} s",
                "single_thread_value": n1 / dt1, "single_thread_us_per_solve": dt1 / n1 * 1e6}
    if not have_ref:
        return port, dt / steps * 1e3
    # oracle/_ref holds two builds of the same unmodified sources: -O2 (baseline x86-64) and -O2 -march=x86-64-v3.  The v3 build
    # is the one timed when this host's CPU has the ISA (it is what SURVEY 8(d) asked for short of -march=native, which cannot
    # travel between hosts); contraction is off in both, so they return the same bits.
    v3_lib = os.path.join(os.path.dirname(Rf.LIB), "libmbik_ref_v3.so")
    build_name = "-O2"
    if os.path.exists(v3_lib) and _host_supports_x86_64_v3() and not os.environ.get("MBIK_REF_BASELINE_ISA"):
        Rf.LIB = v3_lib
        Rf._lib = None
        build_name = "-O2 -march=x86-64-v3 -ffp-contract=off"
    n, dt, n1, dt1 = _time_cpu_solver(Rf.solve_batch, cores, rig, budget_s, steps, warmup)
    ref = {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": f"{n} poses x {steps} pass(es) of the same workload, the reference module's own sources (oracle/_ref: /root/reference/src "
                     f"compiled unmodified {build_name} over the engine stand-in oracle/godot_shim -- reference sources on a stand-in engine, not Godot "
                     f"headless), one long-lived ManyBoneIK3D per thread, all {cores} host threads, {dt:.1f} s",
           "build": build_name,
           "single_thread_value": n1 / dt1, "single_thread_us_per_solve": dt1 / n1 * 1e6}
    if port:
        ref["port"] = port
    return ref, dt / steps * 1e3


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference_arm(args, rank, world, emit):
    if rank != 0:
        return
    from many_bone_ik_b200 import rigs
    rig = rigs.humanoid22()
    cb, ms = time_cpu_reference(rig, budget_s=args.budget_s, steps=max(args.steps, 1), warmup=max(min(args.warmup, 1), 0))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus), "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("reference = the reference module's own C++ sources (ManyBoneIK3D::_process_modification and everything under it) compiled unmodified "
                 "into oracle/_ref over a stand-in of the Godot engine headers; the engine itself is not in the tree" if cb["kind"] == "reference" else
                 "reference = line-cited CPU restatement of the reference solver (oracle/): oracle/_ref was not built on this box"),
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mbik", choices=["mbik", "reference"])
    ap.add_argument("--poses", type=int, default=POSES_PER_GPU, help="poses per GPU (default 2^20); smaller values are for debugging only")
    ap.add_argument("--budget-s", type=float, default=20.0, help="CPU seconds the reference arm may spend per run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: anything libraries print (NCCL's version banner, ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr

    def emit(line):
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()

    if args.impl == "reference":
        run_reference_arm(args, rank, world, emit)
        return 0

    import torch
    import torch.distributed as dist

    from many_bone_ik_b200 import BatchedIKRig, IKStream, rigs, sharding
    from many_bone_ik_b200._capi import MBIK_IO_DEVICE, MBIK_IO_HOST, MBIK_OUT_SOLVED_ONLY, MBIK_SCHED_THROUGHPUT

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: many_bone_ik_b200 has no CPU fallback")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # multi-GPU boxes: keep this rank's threads and pinned buffers on the NUMA node its GPU hangs off (what
    # `numactl --cpunodebind --membind` does); a no-op where the platform exposes no topology (single-node VMs)
    numa_binding = None
    if world > 1:
        try:
            from many_bone_ik_b200 import numa
            numa_binding = numa.bind_to_gpu_node(local_rank)
        except Exception as e:
            numa_binding = {"error": str(e)}

    rig = rigs.humanoid22()
    R = BatchedIKRig(rig)
    total = args.poses * world
    lo, hi = sharding.shard_range(total, rank, world)
    n = hi - lo
    nb, npins, nsolved = rig.n_bones, rig.n_pins, R.info["n_solved"]

    # synthetic inputs: this rank's slice, regenerated independently from the counter-based RNG
    t_host = torch.empty((n, npins, 12), dtype=torch.float32, pin_memory=True)
    CH = 1 << 16
    for s in range(0, n, CH):
        e = min(n, s + CH)
        t_host[s:e] = torch.from_numpy(rigs.random_targets(rig, lo + s, e - s))
    o_host = torch.empty((n, nb, 10), dtype=torch.float32, pin_memory=True)
    t_dev = t_host.to(dev)
    o_dev = torch.empty((n, nb, 10), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        R.solve_raw(n, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)

    # The reference-facing call with HOST buffers.  Output layout of the headline e2e leg: MBIK_OUT_SOLVED_ONLY -- the
    # position / rotation / scale of the bones of bone_list, exactly what _update_skeleton_bones_transform hands to the
    # skeleton (src/many_bone_ik_3d.cpp:104-116; 20 of humanoid22's 22 bones: 800 B per pose).  The full-skeleton layout
    # (all 22 bones, pass-through bones included: 880 B) is timed next to it.
    o_host_c = o_host.view(-1)[: n * nsolved * 10].view(n, nsolved, 10)

    def step_host():
        R.solve_raw(n, t_host.numpy(), o_host_c.numpy(), device=local_rank, flags=MBIK_IO_HOST | MBIK_OUT_SOLVED_ONLY)

    def step_host_full():
        R.solve_raw(n, t_host.numpy(), o_host.numpy(), device=local_rank, flags=MBIK_IO_HOST)

    # ---- device-resident throughput (value) ----
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    local_dev_ms = dev_ms
    barrier()
    # cross-check: the library's own per-launch event pair on a few extra launches (same stream, same inputs)
    for _ in range(min(args.steps, 5)):
        step_device()
        torch.cuda.synchronize()
        kernel_ms.append(R.last_kernel_ms(local_rank))
    dev_ms = max_over_ranks(dev_ms)
    value = total * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers (e2e) ----
    def time_host(step, steps):
        for _ in range(max(1, min(args.warmup, 2))):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        s_ = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return s_
    e2e_s = time_host(step_host, args.steps)
    e2e_value = total * args.steps / e2e_s
    clocks = sampler.stop() if rank == 0 else None
    full_steps = max(1, min(args.steps, 4))
    e2e_full_s = time_host(step_host_full, full_steps)
    # correctness guard on a small sample of what was just computed (device path == host path)
    chk = min(n, 1024)
    same = bool(torch.equal(o_dev[:chk].cpu(), o_host[:chk]))

    # ---- steady state of a crowd that lives on the device: mbik_stream_* (frame f+1 starts from frame f's solution as the
    # skeleton holds it; only targets go up, only the solved bones' poses come down -- or nothing at all) ----
    stream_e2e = None
    frames = 16  # enough frames that the pipeline's fill (first upload) and drain (last download) do not dominate
    stream_local = [float("inf"), float("inf")]  # seconds for `frames` frames with / without the pose download (this rank)
    stream_err = None
    try:
        S = IKStream(R, n, device=local_rank, solved_only=True)
        for k_, download in enumerate((True, False)):
            for _ in range(2):
                S.submit(t_host.numpy(), o_host_c.numpy() if download else None, None)
            S.sync()
            t0 = time.perf_counter()
            for _ in range(frames):
                S.submit(t_host.numpy(), o_host_c.numpy() if download else None, None)
            S.sync()
            stream_local[k_] = time.perf_counter() - t0
        S.close()
        del S
    except Exception as e:  # no collective inside the try: a rank that fails must not leave the others waiting
        stream_err = str(e)
    # the per-frame kernel of a warm-started crowd (start = the previous frame's recomposed poses, same targets): its per-pose
    # work differs from the cold, rest-pose start the headline times (other snaps trigger), so it is timed on its own
    warm_ms = None
    try:
        from many_bone_ik_b200._capi import MBIK_LOCAL_RECOMPOSED
        loc_a = torch.empty((n, nb, 12), dtype=torch.float32, device=dev)
        loc_b = torch.empty((n, nb, 12), dtype=torch.float32, device=dev)
        R.solve_raw(n, t_dev, o_dev, out_local=loc_a, device=local_rank, flags=MBIK_IO_DEVICE | MBIK_LOCAL_RECOMPOSED, stream=stream)
        for _ in range(3):
            R.solve_raw(n, t_dev, o_dev, start_pose=loc_a, out_local=loc_b, device=local_rank, flags=MBIK_IO_DEVICE | MBIK_LOCAL_RECOMPOSED, stream=stream)
            loc_a, loc_b = loc_b, loc_a
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(4):
            R.solve_raw(n, t_dev, o_dev, start_pose=loc_a, out_local=loc_b, device=local_rank, flags=MBIK_IO_DEVICE | MBIK_LOCAL_RECOMPOSED, stream=stream)
            loc_a, loc_b = loc_b, loc_a
        w1.record()
        torch.cuda.synchronize()
        warm_ms = w0.elapsed_time(w1) / 4
        del loc_a, loc_b
    except Exception as e:
        stream_err = (stream_err or "") + f" warm kernel: {e}"
    barrier()
    st_dl, st_no = max_over_ranks(stream_local[0]), max_over_ranks(stream_local[1])
    stream_e2e = {"value_with_pose_download": total * frames / st_dl if np.isfinite(st_dl) else None,
                  "value_targets_only": total * frames / st_no if np.isfinite(st_no) else None, "unit": UNIT, "frames": frames,
                  "h2d_bytes_per_frame": int(n * npins * 48) * world, "d2h_bytes_per_frame_with_download": int(n * nsolved * 40) * world,
                  "error": stream_err, "kernel_ms_per_warm_started_frame_rank0": warm_ms,
                  "note": "mbik_stream_submit x frames, then mbik_stream_sync: warm-started frames of the same 2^20-pose crowd per GPU, pinned host targets; "
                          "'targets_only' leaves the poses on the device (a renderer that skins on the GPU); ranks are not re-synchronised between frames"}

    # ---- side measurement: what the host link gives this rank while ALL ranks copy at once (plain pinned-memory copies of the
    # same buffers, no solve): the ceiling of e2e at N GPUs is min(kernel rate, link rate / bytes per solve) ----
    host_link = None
    link_local = [float("inf"), float("inf")]
    try:
        o_dev_c = o_dev.view(-1)[: n * nsolved * 10].view(n, nsolved, 10)
        o_host_c.copy_(o_dev_c, non_blocking=True)
        t_dev.copy_(t_host, non_blocking=True)
        torch.cuda.synchronize()
        h0, h1, h2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 3
        h0.record()
        for _ in range(reps):
            o_host_c.copy_(o_dev_c, non_blocking=True)
        h1.record()
        for _ in range(reps):
            t_dev.copy_(t_host, non_blocking=True)
        h2.record()
        torch.cuda.synchronize()
        link_local = [h0.elapsed_time(h1) / reps, h1.elapsed_time(h2) / reps]
    except Exception as e:  # a side measurement must never break the headline line (and holds no collective)
        host_link = {"error": str(e)}
    barrier()
    d2h_ms, h2d_ms = max_over_ranks(link_local[0]), max_over_ranks(link_local[1])
    if np.isfinite(d2h_ms) and np.isfinite(h2d_ms):
        d2h_gbs, h2d_gbs = n * nsolved * 40 / (d2h_ms * 1e-3) / 1e9, t_host.numel() * 4 / (h2d_ms * 1e-3) / 1e9
        per_solve_s = max((nsolved * 40) / (d2h_gbs * 1e9), (npins * 48) / (h2d_gbs * 1e9))  # separate copy engines: the slower direction bounds
        host_link = {"d2h_gbs_per_gpu_all_ranks_copying": d2h_gbs, "h2d_gbs_per_gpu_all_ranks_copying": h2d_gbs,
                     "e2e_ceiling_solves_per_s": world / per_solve_s,
                     "e2e_ceiling_full_layout_solves_per_s": world / max((nb * 40) / (d2h_gbs * 1e9), (npins * 48) / (h2d_gbs * 1e9)),
                     "note": f"slowest rank, pinned host buffers of the e2e leg, every rank copying at about the same time, one direction at a time; ceiling = N / max({nsolved * 40} B / d2h, {npins * 48} B / h2d)"}

    # ---- BASELINE configs[2] read literally: ONE 2^20-pose batch sharded over the N GPUs (strong scaling; each rank solves
    # the first 2^20/N poses of its buffer), device-resident.  The launch picks a wave-balanced CTA size (mbik_kernel.cu) ----
    strong = None
    ns_ = min(n, (1 << 20) // world)
    for _ in range(2):
        R.solve_raw(ns_, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        R.solve_raw(ns_, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
    s1.record()
    torch.cuda.synchronize()
    strong_ms = max_over_ranks(s0.elapsed_time(s1))
    barrier()
    strong = {"total_poses": ns_ * world, "poses_per_gpu": ns_, "value": ns_ * world * args.steps / (strong_ms * 1e-3), "unit": UNIT,
              "ms_per_step": strong_ms / args.steps, "scaling": "strong",
              "note": "one 2^20-pose batch split contiguously over the N GPUs (BASELINE configs[2] literally); compare with N x the N=1 `value`"}

    # ---- side measurement (N > 1): the same host-buffer workload from ONE process -- mbik_solve_batch_multi, one worker thread
    # per device, one pinned arena -- against the N-process arrangement above.  Rank 0 runs it while the other ranks wait. ----
    single_process = None
    if world > 1:
        barrier()
        # the other ranks wait on the CPU (a key of the rendezvous store), not inside a NCCL kernel that would share their GPU
        # with rank 0's launches
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                per = min(n, 1 << 19)
                tot = per * world
                th_ = torch.empty((tot, npins, 12), dtype=torch.float32, pin_memory=True)
                for g in range(world):
                    th_[g * per:(g + 1) * per] = t_host[:per]
                oh_ = torch.empty((tot, nsolved, 10), dtype=torch.float32, pin_memory=True)
                devs = list(range(world))
                for _ in range(2):
                    R.solve_raw(tot, th_.numpy(), oh_.numpy(), flags=MBIK_IO_HOST | MBIK_OUT_SOLVED_ONLY, devices=devs)
                reps = max(2, min(args.steps, 5))
                t0 = time.perf_counter()
                for _ in range(reps):
                    R.solve_raw(tot, th_.numpy(), oh_.numpy(), flags=MBIK_IO_HOST | MBIK_OUT_SOLVED_ONLY, devices=devs)
                dt_ = time.perf_counter() - t0
                single_process = {"value": tot * reps / dt_, "unit": UNIT, "total_poses": tot, "devices": world,
                                  "note": "mbik_solve_batch_multi from rank 0's process (host buffers, solved-only layout), the other ranks idle"}
                del th_, oh_
            except Exception as e:
                single_process = {"error": str(e)}
            store.set("mbik_single_process_done", "1")
        else:
            store.wait(["mbik_single_process_done"])
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- p50 latency of a 4096-pose batch, device-resident (BASELINE configs[1]) ----
    lt = t_dev[:LATENCY_BATCH].contiguous()
    lo_ = torch.empty((LATENCY_BATCH, nb, 10), dtype=torch.float32, device=dev)
    # the library picks the segment-parallel mapping for a batch this small (32 poses per CTA, one warp per concurrently
    # solvable segment); the one-thread-per-pose mapping is timed next to it for comparison
    def _p50(extra_flags):
        lat = []
        for i in range(220):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            R.solve_raw(LATENCY_BATCH, lt, lo_, device=local_rank, flags=MBIK_IO_DEVICE | extra_flags, stream=stream)
            b.record()
            torch.cuda.synchronize()
            if i >= 20:
                lat.append(a.elapsed_time(b))
        return float(np.median(lat))
    p50 = _p50(0)
    p50_thread_per_pose = _p50(MBIK_SCHED_THROUGHPUT)

    # small batches, for scale (BASELINE configs[0] is ONE pose on the CPU: 0.41 ms in the reference): p50 of a device-resident
    # call with 1 and 32 poses (one 32-pose group either way), and of a HOST-buffer call with 1 / 64 / 4096 poses (copies
    # included) -- 64 poses is what a CrowdBinding flush of 64 nodes costs, against 64 one-pose calls of the plain binding
    def _p50_n(n_small, host=False):
        lat = []
        th, oh = t_host[:n_small].contiguous().pin_memory(), torch.empty((n_small, nb, 10), dtype=torch.float32).pin_memory()
        for i in range(120):
            if host:
                t0_ = time.perf_counter()
                R.solve_raw(n_small, th.numpy(), oh.numpy(), device=local_rank, flags=MBIK_IO_HOST)
                dt_ = (time.perf_counter() - t0_) * 1e3
            else:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                R.solve_raw(n_small, lt, lo_, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
                b.record()
                torch.cuda.synchronize()
                dt_ = a.elapsed_time(b)
            if i >= 20:
                lat.append(dt_)
        return float(np.median(lat))
    small = {"p50_ms_1_pose_device": _p50_n(1), "p50_ms_32_poses_device": _p50_n(32), "p50_ms_1_pose_host_buffers": _p50_n(1, host=True),
             "p50_ms_64_poses_host_buffers": _p50_n(64, host=True), "p50_ms_4096_poses_host_buffers": _p50_n(LATENCY_BATCH, host=True)}

    # ---- roofline of the one kernel (FP32 CUDA cores; HBM shown as the sanity figure) ----
    import ctypes as C
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf = C.c_double(0)
    R.lib.mbik_measure_fp32_tflops(local_rank, 5, C.byref(tf))
    # average launch duration of the solve kernel over the TIMED region: K back-to-back launches between two CUDA events
    # on the launching stream (this rank's own figure); the library's per-launch event pair, read on a few extra
    # launches afterwards, is reported next to it as a cross-check
    k_ms = float(local_dev_ms / args.steps)
    k_ms_single = float(np.mean(kernel_ms))
    # DRAM traffic of the kernel from the committed ncu --set full capture (bytes per pose x poses of one launch)
    traffic_table = {}
    try:
        traffic_table = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        pass
    traffic, traffic_src = None, None
    if "dram_bytes" in traffic_table:
        traffic = float(traffic_table["dram_bytes"]) / float(traffic_table["poses"]) * n
        traffic_src = traffic_table.get("source")
    flops = R.info["flops_per_solve"]
    bytes_per_solve = npins * 48 + nb * 40
    ach_tf = flops * n / (k_ms * 1e-3) / 1e12
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {
        "bound": "fp32", "kernel": "mbik_solve_kernel<20,4,2,512> (one launch = all iterations of one batch)", "achieved": ach_tf, "peak": float(tf.value), "unit": "TFLOP/s",
        "frac": ach_tf / float(tf.value) if tf.value else None, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": "FP32 FMA micro-benchmark run in this process (mbik_measure_fp32_tflops)",
        "flops_per_solve": flops, "kernel_ms": k_ms, "kernel_ms_single_launch_events": k_ms_single,
        "note": "algorithmic flop floor of SURVEY 8(d); the kernel issues separately rounded FMUL/FADD (bit-exact parity), so 50% of the FMA peak is its structural ceiling",
        "hbm": {"bound": "hbm", "achieved": bytes_per_solve * n / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_per_solve * n / (k_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_solve": bytes_per_solve,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"},
    }

    # ---- BASELINE configs[3] (chain64, 30 iterations) and configs[4] (quad80, 15 iterations): the same measurements as the
    # headline on a 65 536-pose batch -- device-resident rate, e2e through host buffers, p50 of a 4096-pose
    # batch, the reference CPU path on a bounded sample beside it, and the roofline that binds these rigs (HBM: the
    # per-pose state does not fit on chip) with the FP32 figure next to it ----
    configs = {}
    for name in ("chain64", "quad80"):
        try:
            rg = rigs.RIGS[name]()
            Rg = BatchedIKRig(rg)
            m = 65536
            ns_g = Rg.info["n_solved"]
            th_g = torch.from_numpy(rigs.random_targets(rg, 0, m)).pin_memory()
            tg = th_g.to(dev)
            og = torch.empty((m, rg.n_bones, 10), dtype=torch.float32, device=dev)
            oh_g = torch.empty((m, ns_g, 10), dtype=torch.float32, pin_memory=True)
            for _ in range(2):
                Rg.solve_raw(m, tg, og, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps_g = 3
            a_.record()
            for _ in range(reps_g):
                Rg.solve_raw(m, tg, og, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
            b_.record()
            torch.cuda.synchronize()
            ms = a_.elapsed_time(b_) / reps_g
            Rg.solve_raw(m, th_g.numpy(), oh_g.numpy(), device=local_rank, flags=MBIK_IO_HOST | MBIK_OUT_SOLVED_ONLY)
            t0_ = time.perf_counter()
            for _ in range(reps_g):
                Rg.solve_raw(m, th_g.numpy(), oh_g.numpy(), device=local_rank, flags=MBIK_IO_HOST | MBIK_OUT_SOLVED_ONLY)
            e2e_g = m * reps_g / (time.perf_counter() - t0_)
            lat_g = []
            for i in range(40):  # p50 of a 4096-pose batch (the mapping is the library's choice), 30 calls after 10 warm-ups
                c_, d_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c_.record()
                Rg.solve_raw(LATENCY_BATCH, tg, og, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
                d_.record()
                torch.cuda.synchronize()
                if i >= 10:
                    lat_g.append(c_.elapsed_time(d_))
            fl = Rg.info["flops_per_solve"]
            alg_bytes = rg.n_pins * 48 + rg.n_bones * 40
            fp32 = {"achieved": fl * m / (ms * 1e-3) / 1e12, "peak": float(tf.value), "unit": "TFLOP/s",
                    "frac": fl * m / (ms * 1e-3) / 1e12 / float(tf.value) if tf.value else None, "flops_per_solve": fl}
            roof = {"bound": "hbm", "kernel": traffic_table.get("other_rigs", {}).get(name, {}).get("kernel", "mbik_solve_kernel_glw<64,*>"), "unit": "GB/s", "peak": hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6.65 TB/s",
                    "algorithmic_bytes_per_solve": alg_bytes, "algorithmic_gbs": alg_bytes * m / (ms * 1e-3) / 1e9, "fp32": fp32}
            tr = traffic_table.get("other_rigs", {}).get(name)
            if tr:  # DRAM bytes per pose of this rig's kernel from the committed ncu --set full capture x the poses of this launch
                t_bytes = float(tr["dram_bytes"]) / float(tr["poses"]) * m
                roof.update({"achieved": t_bytes / (ms * 1e-3) / 1e9, "frac": t_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": t_bytes,
                             "traffic_source": tr["source"],
                             "note": "achieved = measured DRAM traffic of the launch (committed ncu capture of the same kernel) / its duration measured here: the "
                                     "effector walks re-read the per-pose local transforms, which do not fit L2 for a resident batch.  With thread-local state (round 1) "
                                     "these rigs waited on HBM (chain64: 4.3 TB/s, stall_long_sb 50 %); the streamed-walk instantiation (cp.async ring + L2 policies) "
                                     "cut the traffic by a third and hid its latency, so the kernel is now limited by dependent-instruction latency like humanoid22 -- "
                                     "the FP32 fraction is under `fp32`"})
            cfg = {"workload": f"{name}: {rg.n_bones} bones, {rg.n_pins} effectors, {rg.iterations} iterations, {m} random-target poses",
                   "value": m / (ms * 1e-3), "unit": UNIT, "poses": m, "iterations": rg.iterations, "ms_per_launch": ms,
                   "e2e": {"value": e2e_g, "unit": UNIT, "h2d_bytes_per_step": int(m * rg.n_pins * 48), "d2h_bytes_per_step": int(m * ns_g * 40)},
                   "latency_p50_ms_4096": float(np.median(lat_g)), "segment_parallel_warps": Rg.info["sp_roles"], "roofline": roof}
            if not args.no_cpu_baseline and world == 1:
                cb_g, _ = time_cpu_reference(rg, budget_s=6.0, with_port=False)
                cfg["cpu_baseline"] = cb_g
                cfg["e2e_over_cpu_baseline"] = e2e_g / cb_g["value"] if cb_g.get("value") else None
                cfg["cpu_baseline"]["latency_ms_4096"] = LATENCY_BATCH / cb_g["value"] * 1e3 if cb_g.get("value") else None
            configs[name] = cfg
            del Rg, tg, og, th_g, oh_g
        except Exception as e:  # never let the side measurements break the headline line
            configs[name] = {"error": str(e)}

    # ---- side measurement: per-pose limit sets: alternative fills of the constraint tables authored on a
    # pool of host threads (sets per second), then pose k solved with set k % 4 ----
    limit_sets = None
    try:
        import copy

        def variant(s_i):
            cset = copy.deepcopy(rig.constraints)
            for c in cset:
                c["twist_range"] = float(np.float32(c["twist_range"] * (1.0 - 0.15 * (s_i % 4))))
                c["cones"] = [(cx, cy, cz, float(np.float32(r * (1.0 - 0.1 * (s_i % 4) - 1e-4 * (s_i // 4))))) for (cx, cy, cz, r) in c["cones"]]
            return cset
        many = [variant(k) for k in range(2048)]
        hm = R.create_limit_sets(many)
        info_m = R.limit_sets_info(hm)
        R.destroy_limit_sets(hm)
        os.environ["MBIK_AUTHOR_THREADS"] = "1"  # the same authoring on one host thread, for the scaling of the pool
        h1_ = R.create_limit_sets(many[:512])
        info_1 = R.limit_sets_info(h1_)
        R.destroy_limit_sets(h1_)
        del os.environ["MBIK_AUTHOR_THREADS"]
        hs = R.create_limit_sets(many[:4])
        idx_dev = (torch.arange(n, device=dev, dtype=torch.int32) % 4).contiguous()
        for _ in range(2):
            R.solve_with_limits_raw(hs, n, idx_dev, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(3):
            R.solve_with_limits_raw(hs, n, idx_dev, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
        l1.record()
        torch.cuda.synchronize()
        lms = l0.elapsed_time(l1) / 3
        lat_l = []
        for i in range(60):  # a 4096-pose crowd with per-pose limits: the segment-parallel mapping reads the limit-set record too
            c_, d_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c_.record()
            R.solve_with_limits_raw(hs, LATENCY_BATCH, idx_dev, lt, lo_, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
            d_.record()
            torch.cuda.synchronize()
            if i >= 10:
                lat_l.append(c_.elapsed_time(d_))
        limit_sets = {"sets": 4, "poses": n, "ms_per_launch": lms, "solves_per_s": n / (lms * 1e-3), "latency_p50_ms_4096": float(np.median(lat_l)),
                      "authoring": {"sets": info_m["n_sets"], "seconds": info_m["author_seconds"], "host_threads": info_m["author_threads"],
                                    "sets_per_s": info_m["n_sets"] / info_m["author_seconds"] if info_m["author_seconds"] > 0 else None,
                                    "table_bytes": info_m["table_bytes"], "bytes_per_set": info_m["bytes_per_set"],
                                    "sets_per_s_one_thread": info_1["n_sets"] / info_1["author_seconds"] if info_1["author_seconds"] > 0 else None},
                      "note": "mbik_limit_sets_create (host authoring on all host threads) / mbik_solve_batch_limits, device-resident: kusudama data read per pose "
                              "from a device table instead of the rig blob"}
        R.destroy_limit_sets(hs)
    except Exception as e:
        limit_sets = {"error": str(e)}

    # ---- side measurement: stabilisation passes: the headline rig with stabilization_passes = 1 -- every
    # bone-step walks its effectors a second time for the accept / revert test (src/ik_bone_segment_3d.cpp:163-176) ----
    stabilisation = None
    try:
        import copy
        rig_s = copy.deepcopy(rig)
        rig_s.stabilization_passes = 1
        rig_s.name = rig.name + "_stabilized"
        Rs_ = BatchedIKRig(rig_s)
        for _ in range(2):
            Rs_.solve_raw(n, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
        s0_, s1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0_.record()
        for _ in range(3):
            Rs_.solve_raw(n, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
        s1_.record()
        torch.cuda.synchronize()
        sms = s0_.elapsed_time(s1_) / 3
        lat_s = []
        for i in range(60):
            c_, d_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c_.record()
            Rs_.solve_raw(LATENCY_BATCH, lt, lo_, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
            d_.record()
            torch.cuda.synchronize()
            if i >= 10:
                lat_s.append(c_.elapsed_time(d_))
        stabilisation = {"workload": f"{rig.name} with stabilization_passes = 1, {n} poses, device-resident", "poses": n, "ms_per_launch": sms,
                         "solves_per_s": n / (sms * 1e-3), "relative_to_plain": (dev_ms / args.steps) / sms, "latency_p50_ms_4096": float(np.median(lat_s)),
                         "flops_per_solve": Rs_.info["flops_per_solve"]}
        if not args.no_cpu_baseline and world == 1:
            cb_s, _ = time_cpu_reference(rig_s, budget_s=4.0, with_port=False)
            stabilisation["cpu_baseline"] = cb_s
            stabilisation["over_cpu_baseline"] = stabilisation["solves_per_s"] / cb_s["value"] if cb_s.get("value") else None
        del Rs_
    except Exception as e:
        stabilisation = {"error": str(e)}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cpu_baseline, _ = time_cpu_reference(rig, budget_s=12.0)
        if cpu_baseline and cpu_baseline.get("value"):
            # the latency metric on the same CPU arm: one 4096-pose batch at the all-thread rate
            cpu_baseline["latency_ms_4096"] = LATENCY_BATCH / cpu_baseline["value"] * 1e3

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world, args.poses),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * npins * 48) * world, "d2h_bytes_per_step": int(n * nsolved * 40) * world,
                "ms_per_step": e2e_s / args.steps * 1e3,
                "output_layout": f"MBIK_OUT_SOLVED_ONLY: position / rotation / scale of the {nsolved} bones of bone_list (what the reference writes to the skeleton)",
                "full_skeleton_layout": {"value": total * full_steps / e2e_full_s, "unit": UNIT, "d2h_bytes_per_step": int(n * nb * 40) * world, "steps": full_steps}},
        "gpu_launches": args.steps * world,
        "strong_scaling_1M_batch": strong,
        "latency_p50_ms_4096": p50,
        "latency_p50_ms_4096_thread_per_pose_mapping": p50_thread_per_pose,
        "latency_small_batches": small,
        "latency_mapping": "segment-parallel: 32 poses per CTA, %d warps (one per concurrently solvable segment), %d phases per iteration" % (R.info["sp_roles"], R.info["sp_phases"]),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "configs": configs,
        "stream_e2e": stream_e2e,
        "single_process_multi_gpu_e2e": single_process,
        "host_link": host_link,
        "numa_binding_rank0": numa_binding,
        "limit_sets_device_resident": limit_sets,
        "stabilisation_device_resident": stabilisation,
        "clocks": clocks,
        "device_equals_host_path": same,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
