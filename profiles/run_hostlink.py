"""Host-link diagnostic for the e2e leg at N GPUs (run under torchrun, one rank per GPU):
prints, per rank, the GPU's PCI bus id and NUMA node, the CPUs / memory nodes the process may use, and the pinned-memory
copy bandwidth of this rank (a) alone and (b) with every rank copying at once, for buffers allocated (1) with the default
memory policy and (2) after binding the process (CPU affinity + memory policy) to the GPU's NUMA node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 profiles/run_hostlink.py
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from many_bone_ik_b200 import numa  # noqa: E402


def bw(host, dev, d2h, reps=3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    (host.copy_(dev, non_blocking=True) if d2h else dev.copy_(host, non_blocking=True))
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        (host.copy_(dev, non_blocking=True) if d2h else dev.copy_(host, non_blocking=True))
    b.record()
    torch.cuda.synchronize()
    return host.numel() * 4 * reps / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    info = numa.describe(lr)
    n = 256 << 20  # 1 GiB of float32
    d = torch.empty(n, dtype=torch.float32, device=dev)
    res = {"rank": rank, **info}
    for label in ("default", "bound"):
        if label == "bound":
            res["bind"] = numa.bind_to_gpu_node(lr)
        h = torch.empty(n, dtype=torch.float32, pin_memory=True)
        h.fill_(1.0)
        for r in range(world):  # alone, rank by rank
            barrier()
            if r == rank:
                res[f"{label}_alone_d2h"] = bw(h, d, True)
                res[f"{label}_alone_h2d"] = bw(h, d, False)
        barrier()
        res[f"{label}_all_d2h"] = bw(h, d, True)
        barrier()
        res[f"{label}_all_h2d"] = bw(h, d, False)
        barrier()
        del h
    out = [None] * world
    if world > 1:
        dist.all_gather_object(out, res)
    else:
        out = [res]
    if rank == 0:
        for r in out:
            print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
