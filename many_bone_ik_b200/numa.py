"""NUMA placement helpers for the host side of the e2e path (bench / test glue, Linux only, no extra dependency).

A multi-GPU box has the GPUs spread over the CPU sockets; pinned host buffers that live on the other socket's memory
make every H2D / D2H copy cross the inter-socket link.  `bind_to_gpu_node(device)` does for one process what
`numactl --cpunodebind=N --membind=N` does: CPU affinity and memory policy to the NUMA node the GPU hangs off, taken
from /sys/bus/pci/devices/<bdf>/numa_node.  Every step degrades to a no-op (with the reason returned) when the
platform does not expose the topology or does not allow the binding (containers with a restricted cpuset)."""
from __future__ import annotations

import ctypes
import os

_SYS_SET_MEMPOLICY = 238  # x86_64
_MPOL_DEFAULT, _MPOL_PREFERRED, _MPOL_BIND = 0, 1, 2


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_list(text):
    out = []
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def gpu_pci_bus_id(device):
    """'dddd:bb:dd.0' of CUDA device `device`, or None (no CUDA device, or the runtime does not say)."""
    try:
        import torch
        if not torch.cuda.is_available() or device >= torch.cuda.device_count():
            return None
        p = torch.cuda.get_device_properties(device)
        if hasattr(p, "pci_bus_id") and hasattr(p, "pci_device_id"):
            return "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None
    try:
        import subprocess
        r = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)], capture_output=True, text=True, timeout=10)
        bdf = r.stdout.strip().lower()
        return bdf[-12:] if len(bdf) >= 12 else None  # nvidia-smi prints an 8-digit domain
    except Exception:
        return None


def gpu_numa_node(device):
    bdf = gpu_pci_bus_id(device)
    if not bdf:
        return None
    v = _read(f"/sys/bus/pci/devices/{bdf}/numa_node")
    try:
        return int(v)
    except (TypeError, ValueError):
        return None


def describe(device):
    nodes = _parse_list(_read("/sys/devices/system/node/online"))
    status = _read("/proc/self/status") or ""
    mems = [ln.split(":")[1].strip() for ln in status.splitlines() if ln.startswith("Mems_allowed_list")]
    return {"pci": gpu_pci_bus_id(device), "gpu_numa_node": gpu_numa_node(device), "nodes_online": nodes,
            "cpus_allowed": len(os.sched_getaffinity(0)), "cpus_allowed_list": sorted(os.sched_getaffinity(0))[:4] + ["..."],
            "mems_allowed": mems[0] if mems else None,
            "node_cpulists": {n: _read(f"/sys/devices/system/node/node{n}/cpulist") for n in nodes}}


def bind_to_gpu_node(device, strict=False):
    """CPU affinity + memory policy of THIS process to the GPU's NUMA node.  Returns a dict saying what was done."""
    done = {"node": None, "cpus": None, "mempolicy": None}
    node = gpu_numa_node(device)
    if node is None or node < 0:
        done["skipped"] = "GPU NUMA node unknown (numa_node = %r)" % (node,)
        return done
    done["node"] = node
    cpus = set(_parse_list(_read(f"/sys/devices/system/node/node{node}/cpulist"))) & os.sched_getaffinity(0)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
            done["cpus"] = len(cpus)
        except OSError as e:
            done["cpus"] = f"failed: {e}"
    else:
        done["cpus"] = "none of the node's CPUs are allowed for this process"
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        mode = _MPOL_BIND if strict else _MPOL_PREFERRED
        rc = libc.syscall(_SYS_SET_MEMPOLICY, mode, ctypes.byref(mask), 16 * 64 + 1)
        done["mempolicy"] = "bind" if (rc == 0 and strict) else ("preferred" if rc == 0 else f"failed: errno {ctypes.get_errno()}")
    except Exception as e:
        done["mempolicy"] = f"failed: {e}"
    return done
