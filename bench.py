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


def time_cpu_reference(rig, budget_s=12.0, steps=1, warmup=0):
    from oracle import oracle_py as O
    from oracle import reference_py as Rf
    O.build()
    cores = O.hardware_threads()
    have_ref = os.path.exists(Rf.LIB) or Rf.available()
    port_budget = budget_s * (0.35 if have_ref else 1.0)
    n, dt, n1, dt1 = _time_cpu_solver(O.solve_batch, cores, rig, port_budget, steps, warmup)
    port = {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} poses x {steps} pass(es) of the same workload, oracle restatement, all {cores} host threads, {dt:.1f} s",
            "single_thread_value": n1 / dt1, "single_thread_us_per_solve": dt1 / n1 * 1e6}
    if not have_ref:
        return port, dt / steps * 1e3
    n, dt, n1, dt1 = _time_cpu_solver(Rf.solve_batch, cores, rig, budget_s, steps, warmup)
    ref = {"value": n * steps / dt, "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": f"{n} poses x {steps} pass(es) of the same workload, the reference module's own sources (oracle/_ref: /root/reference/src "
                     f"compiled unmodified -O2 over the engine stand-in oracle/godot_shim), one long-lived ManyBoneIK3D per thread, all {cores} host threads, {dt:.1f} s",
           "single_thread_value": n1 / dt1, "single_thread_us_per_solve": dt1 / n1 * 1e6,
           "port": port}
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

    from many_bone_ik_b200 import BatchedIKRig, rigs, sharding
    from many_bone_ik_b200._capi import MBIK_IO_DEVICE, MBIK_IO_HOST, MBIK_SCHED_THROUGHPUT

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
    nb, npins = rig.n_bones, rig.n_pins

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

    def step_host():
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
        # per-launch kernel time from the library's own event pair on the launching stream (read after sync below)
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    local_dev_ms = dev_ms
    barrier()
    # average kernel duration: time K more launches individually (same stream, same inputs)
    for _ in range(min(args.steps, 5)):
        step_device()
        torch.cuda.synchronize()
        kernel_ms.append(R.last_kernel_ms(local_rank))
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    value = total * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers (e2e) ----
    for _ in range(max(1, min(args.warmup, 2))):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = total * args.steps / e2e_s

    # ---- side measurement: what the host link gives this rank while ALL ranks copy at once (plain pinned-memory copies of the
    # same buffers, no solve): the ceiling of e2e at N GPUs is min(kernel rate, link rate / 1120 B per solve) ----
    host_link = None
    try:
        for _ in range(1):
            o_host.copy_(o_dev, non_blocking=True)
            t_dev.copy_(t_host, non_blocking=True)
        barrier()
        h0, h1, h2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 3
        h0.record()
        for _ in range(reps):
            o_host.copy_(o_dev, non_blocking=True)
        h1.record()
        for _ in range(reps):
            t_dev.copy_(t_host, non_blocking=True)
        h2.record()
        torch.cuda.synchronize()
        d2h_ms, h2d_ms = max_over_ranks(h0.elapsed_time(h1)) / reps, max_over_ranks(h1.elapsed_time(h2)) / reps
        barrier()
        d2h_gbs, h2d_gbs = o_host.numel() * 4 / (d2h_ms * 1e-3) / 1e9, t_host.numel() * 4 / (h2d_ms * 1e-3) / 1e9
        per_solve_s = (nb * 40) / (d2h_gbs * 1e9)  # D2H and H2D run on separate copy engines: the slower direction bounds
        per_solve_s = max(per_solve_s, (npins * 48) / (h2d_gbs * 1e9))
        host_link = {"d2h_gbs_per_gpu_all_ranks_copying": d2h_gbs, "h2d_gbs_per_gpu_all_ranks_copying": h2d_gbs,
                     "e2e_ceiling_solves_per_s": world / per_solve_s,
                     "note": "slowest rank, pinned host buffers of the e2e leg, every rank copying simultaneously; ceiling = N / max(880 B / d2h, 240 B / h2d)"}
    except Exception as e:  # a side measurement must never break the headline line
        host_link = {"error": str(e)}

    # ---- side measurement: BASELINE configs[2] read literally = ONE 2^20-pose batch sharded over the N GPUs (strong
    # scaling; each rank solves the first 2^20/N poses of its buffer), device-resident ----
    strong = None
    if world > 1:
        ns = min(n, (1 << 20) // world)
        for _ in range(2):
            R.solve_raw(ns, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            R.solve_raw(ns, t_dev, o_dev, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
        s1.record()
        torch.cuda.synchronize()
        strong_ms = max_over_ranks(s0.elapsed_time(s1))
        barrier()
        strong = {"total_poses": ns * world, "poses_per_gpu": ns, "value": ns * world * args.steps / (strong_ms * 1e-3), "unit": UNIT,
                  "ms_per_step": strong_ms / args.steps}

    # correctness guard on a small sample of what was just computed (device path == host path, finite)
    chk = min(n, 1024)
    same = bool(torch.equal(o_dev[:chk].cpu(), o_host[:chk]))

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
    # call with 1 and 32 poses (one 32-pose group either way), and of a HOST-buffer call with 4096 poses (copies included)
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
    small = {"p50_ms_1_pose_device": _p50_n(1), "p50_ms_32_poses_device": _p50_n(32), "p50_ms_4096_poses_host_buffers": _p50_n(LATENCY_BATCH, host=True)}

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
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = float(tj["dram_bytes"]) / float(tj["poses"]) * n
        traffic_src = tj.get("source")
    except Exception:
        pass
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

    # ---- the other BASELINE configs (4: chain64, 30 iterations; 5: quad80, 15 iterations), device-resident, brief ----
    other = {}
    for name in ("chain64", "quad80"):
        try:
            rg = rigs.RIGS[name]()
            Rg = BatchedIKRig(rg)
            m = 148 * 512
            tg = torch.from_numpy(rigs.random_targets(rg, 0, m)).to(dev)
            og = torch.empty((m, rg.n_bones, 10), dtype=torch.float32, device=dev)
            for _ in range(2):
                Rg.solve_raw(m, tg, og, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(3):
                Rg.solve_raw(m, tg, og, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
            b_.record()
            torch.cuda.synchronize()
            ms = a_.elapsed_time(b_) / 3
            lat_g = []
            for i in range(40):  # p50 of a 4096-pose batch (the mapping is the library's choice), 30 calls after 10 warm-ups
                c_, d_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c_.record()
                Rg.solve_raw(LATENCY_BATCH, tg, og, device=local_rank, flags=MBIK_IO_DEVICE, stream=stream)
                d_.record()
                torch.cuda.synchronize()
                if i >= 10:
                    lat_g.append(c_.elapsed_time(d_))
            other[name] = {"solves_per_s": m / (ms * 1e-3), "poses": m, "iterations": rg.iterations, "ms_per_launch": ms,
                           "latency_p50_ms_4096": float(np.median(lat_g)), "segment_parallel_warps": Rg.info["sp_roles"],
                           "flops_per_solve": Rg.info["flops_per_solve"],
                           "fp32_roofline_frac": Rg.info["flops_per_solve"] * m / (ms * 1e-3) / 1e12 / float(tf.value) if tf.value else None}
            try:  # DRAM traffic of this rig's kernel from the committed ncu capture -> fraction of the measured HBM peak
                tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["other_rigs"][name]
                gbs = float(tr["dram_bytes"]) / float(tr["poses"]) * m / (ms * 1e-3) / 1e9
                other[name]["hbm"] = {"traffic_bytes_per_launch": float(tr["dram_bytes"]) / float(tr["poses"]) * m, "achieved_gbs": gbs,
                                      "frac_of_measured_peak": gbs / hbm_peak, "source": tr["source"]}
            except Exception:
                pass
            del Rg, tg, og
        except Exception as e:  # never let the side measurements break the headline line
            other[name] = {"error": str(e)}

    # ---- side measurement: per-pose limit sets: 4 alternative fills of the constraint tables, pose k uses set k % 4 ----
    limit_sets = None
    try:
        import copy
        cs = [copy.deepcopy(rig.constraints) for _ in range(4)]
        for s_i, cset in enumerate(cs):
            for c in cset:
                c["twist_range"] = float(np.float32(c["twist_range"] * (1.0 - 0.15 * s_i)))
                c["cones"] = [(cx, cy, cz, float(np.float32(r * (1.0 - 0.1 * s_i)))) for (cx, cy, cz, r) in c["cones"]]
        hs = R.create_limit_sets(cs)
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
        limit_sets = {"sets": 4, "poses": n, "ms_per_launch": lms, "solves_per_s": n / (lms * 1e-3),
                      "note": "mbik_solve_batch_limits, device-resident: kusudama data read per pose from a device table instead of the rig blob"}
        R.destroy_limit_sets(hs)
    except Exception as e:
        limit_sets = {"error": str(e)}

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
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * npins * 48) * world, "d2h_bytes_per_step": int(n * nb * 40) * world,
                "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": args.steps * world,
        "latency_p50_ms_4096": p50,
        "latency_p50_ms_4096_thread_per_pose_mapping": p50_thread_per_pose,
        "latency_small_batches": small,
        "latency_mapping": "segment-parallel: 32 poses per CTA, %d warps (one per concurrently solvable segment), %d phases per iteration" % (R.info["sp_roles"], R.info["sp_phases"]),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "other_rigs_device_resident": other,
        "strong_scaling_1M_batch": strong,
        "host_link": host_link,
        "numa_binding_rank0": numa_binding,
        "limit_sets_device_resident": limit_sets,
        "clocks": clocks,
        "device_equals_host_path": same,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
