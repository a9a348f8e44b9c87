"""API soak on the parity soak's rigs (tests/rig_cases.soak_rig): beyond the plain solve, per rig
  * per-pose limit sets (3 sets; thread-per-pose and whatever the library picks) == the oracle on a rig with that set's values,
  * a device-resident stream over 3 frames (full and solved-only layout) == the oracle re-seeded from the recomposed previous frame,
  * device buffers + MBIK_OUT_SOLVED_ONLY == the solved rows of the host-buffer result,
  * a 20 000-pose batch in the throughput mapping (wave-balanced CTAs / streamed walk / several launches of the unbounded variant):
    its head and tail == the same poses solved as a small batch.
      python profiles/run_api_soak.py [--rigs 160] [--first 0]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import limit_set_cases as LS  # noqa: E402
import rig_cases  # noqa: E402
import torch  # noqa: E402
from many_bone_ik_b200 import BatchedIKRig, IKStream, MbikError, _capi, rigs  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rigs", type=int, default=160)
ap.add_argument("--first", type=int, default=0)
ap.add_argument("--big", type=int, default=20000)
a = ap.parse_args()


def same(x, y):
    return x.shape == y.shape and bool(np.all((x == y) | (np.isnan(x) & np.isnan(y))))


t0 = time.time()
fails = []
counts = dict(rigs=0, limit_sets=0, streams=0, device_io=0, big=0)
for k in range(a.first, a.first + a.rigs):
    rig = rig_cases.soak_rig(k)
    try:
        R = BatchedIKRig(rig)
    except MbikError as e:
        fails.append((k, "rejected: %s" % e))
        continue
    counts["rigs"] += 1
    n = 40
    T = rigs.random_targets(rig, 5000 + k, n)
    ref = O.solve_batch(rig, T, want_local=True, threads=8)

    def check(tag, ok):
        if not ok:
            fails.append((k, tag))
            print("MISMATCH rig", k, rig.name, "bones", rig.n_bones, "solved", R.info["n_solved"], "capacity", R.info["kernel_capacity"], tag, flush=True)

    # --- limit sets ---
    if rig.constraints and R.info["n_solved"] > 0:
        sets = LS.variants(rig, 3, seed=k)
        h = R.create_limit_sets(sets, asynchronous=bool(k & 1))
        idx = (np.arange(n) % 3).astype(np.int32)
        want = [O.solve_batch(LS.rig_with(rig, s), T, threads=8) for s in sets]
        for sched in ("throughput", "auto"):
            out, st = R.solve_with_limits(h, idx, T, sched=sched)
            ok = all(same(out[idx == s], want[s][0][idx == s]) and np.array_equal(st[idx == s], want[s][1][idx == s]) for s in range(3))
            check("limit sets / " + sched, ok)
        R.destroy_limit_sets(h)
        counts["limit_sets"] += 1
    # --- stream over frames ---
    frames = 3
    Ts = [rigs.random_targets(rig, 6000 + 7 * k + f, n) for f in range(frames)]
    order = R.bone_order()
    for solved_only in (False, True):
        S = IKStream(R, n, device=0, solved_only=solved_only)
        rows = len(order) if solved_only else rig.n_bones
        outs = [np.empty((n, rows, 10), np.float32) for _ in range(frames)]
        sts = [np.empty(n, np.uint32) for _ in range(frames)]
        for f in range(frames):
            S.submit(Ts[f], outs[f], sts[f])
        S.sync()
        start = None
        ok = True
        for f in range(frames):
            r_out, r_loc, r_st = O.solve_batch(rig, Ts[f], start_pose=start, want_local=True, threads=8)
            ok = ok and same(outs[f], r_out[:, order] if solved_only else r_out) and np.array_equal(sts[f], r_st)
            start = O.recompose_pose(rig, r_out, start)
        ok = ok and same(S.read_local(), start)
        S.close()
        check("stream solved_only=%s" % solved_only, ok)
    counts["streams"] += 1
    # --- device buffers, compact layout ---
    if len(order) > 0:
        t_dev = torch.from_numpy(T).cuda()
        o_dev = torch.full((n, len(order), 10), float("nan"), dtype=torch.float32, device="cuda")
        s_dev = torch.zeros(n, dtype=torch.int32, device="cuda")
        R.solve_raw(n, t_dev, o_dev, out_status=s_dev, device=0, flags=_capi.MBIK_IO_DEVICE | _capi.MBIK_OUT_SOLVED_ONLY,
                    stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        check("device io / solved only", same(o_dev.cpu().numpy(), ref[0][:, order]) and np.array_equal(s_dev.cpu().numpy().view(np.uint32), ref[2]))
        counts["device_io"] += 1
    # --- a large batch in the throughput mapping: head and tail against small batches ---
    if k % 4 == 3:
        nb = a.big if R.info["n_solved"] <= 256 else 3000
        Tb = rigs.random_targets(rig, 9000 + k, nb)
        big = R.solve(Tb, want_local=True, sched="throughput")
        head = R.solve(Tb[:64], want_local=True)
        tail = R.solve(Tb[-64:], want_local=True, sched="throughput")
        ok = all(same(b[:64], h_) and same(b[-64:], t_) for b, h_, t_ in zip(big, head, tail))
        h_ref = O.solve_batch(rig, Tb[:16], want_local=True, threads=8)
        ok = ok and all(same(b[:16], r) for b, r in zip(big, h_ref))
        check("big batch", ok)
        counts["big"] += 1
print(f"api soak: rigs {a.first}..{a.first + a.rigs - 1}: {counts}; {len(fails)} failures {fails[:10]}; {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
