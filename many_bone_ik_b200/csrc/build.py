"""Build many_bone_ik_b200/libmbik.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python -m many_bone_ik_b200.csrc.build [--force]

-fmad=false / -ffp-contract=off are belt and braces: every arithmetic op of the solve already goes through
individually rounded intrinsics (mbik_math.cuh), which is what makes the kernel bit-identical to the
reference arithmetic.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libmbik.so")
SOURCES = ["mbik_capi.cu", "mbik_flatten.cu", "mbik_kernel.cu", "mbik_kernel_v0.cu", "mbik_kernel_v1.cu", "mbik_kernel_v2.cu", "mbik_kernel_v3.cu",
           "mbik_kernel_v4.cu", "mbik_kernel_v5.cu", "mbik_kernel_v6.cu", "mbik_kernel_l0.cu", "mbik_kernel_l1.cu", "mbik_kernel_l2.cu", "mbik_kernel_l3.cu", "mbik_kernel_l4.cu", "mbik_kernel_l5.cu", "mbik_kernel_sp0.cu", "mbik_kernel_sp1.cu", "mbik_kernel_sp3.cu", "mbik_kernel_sp4.cu", "mbik_peak.cu", "mbik_selftest.cu"]
HEADERS = ["mbik_blob.h", "mbik_flatten.h", "mbik_kernel.h", "mbik_kernel_body.cuh", "mbik_math.cuh", os.path.join("..", "..", "include", "mbik.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden", "-shared", "-cudart", "static",
]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


OBJ_DIR = os.path.join(HERE, "_obj")
COMPILE_FLAGS = [f for f in NVCC_FLAGS if f not in ("-shared",)]


def _deps_mtime():
    return max(os.path.getmtime(os.path.join(HERE, h)) for h in HEADERS + [os.path.basename(__file__)])


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile each .cu to an object (only the stale ones, in parallel) and link libmbik.so."""
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _deps_mtime()
    procs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(HERE, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), hdr_t):
            continue
        cmd = [nvcc_path()] + [f for f in COMPILE_FLAGS if f != "-cudart" and f != "static"] + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, sp]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        failed = failed or pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libmbik.so")
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-o", OUT] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed linking libmbik.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
