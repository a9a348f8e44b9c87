"""Build an experimental variant of libmbik.so: recompile some translation units with extra -D flags (or a replacement
source) and link them with the objects of the stock build.

    python profiles/build_variant.py <tag> [-DNAME=VALUE ...] [--tu mbik_kernel_v0.cu ...] [--src mbik_kernel_v0.cu=/path/other.cu]

Output: many_bone_ik_b200/_variants/libmbik_<tag>.so (git-ignored, travels with gpurun); use it with MBIK_LIB=<path>."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from many_bone_ik_b200.csrc import build as B

tag = sys.argv[1]
defs = [a for a in sys.argv[2:] if a.startswith("-D") or a.startswith("-maxrregcount")]
tus, repl = [], {}
args = sys.argv[2:]
for i, a in enumerate(args):
    if a == "--tu":
        tus.append(args[i + 1])
    if a == "--src":
        k, v = args[i + 1].split("=")
        repl[k] = v
        tus.append(k)
if not tus:
    tus = ["mbik_kernel_v0.cu"]
B.build()
out_dir = os.path.join(B.PKG, "_variants")
obj_dir = os.path.join(out_dir, "obj_" + tag)
os.makedirs(obj_dir, exist_ok=True)
objs, procs = [], []
for src in B.SOURCES:
    if src in tus:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [B.nvcc_path()] + [f for f in B.COMPILE_FLAGS if f not in ("-cudart", "static")] + defs + ["-I", B.HERE, "-Xptxas", "-v", "-c", "-o", obj,
                                                                                                         repl.get(src, os.path.join(B.HERE, src))]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    else:
        obj = os.path.join(B.OBJ_DIR, src.replace(".cu", ".o"))
    objs.append(obj)
for src, pr in procs:
    out, _ = pr.communicate()
    lines = out.splitlines()
    for i, ln in enumerate(lines):
        if "Used" in ln and "registers" in ln:
            fn = [x for x in lines[:i] if "Compiling entry" in x][-1].split("'")[1]
            sp = [x for x in lines[:i] if "spill" in x][-1].strip()
            print(f"  {fn[:70]:70s} {ln.split(':')[1].split(',')[0].strip():22s} {sp}")
    if pr.returncode != 0:
        sys.stderr.write(out)
        sys.exit(1)
so = os.path.join(out_dir, f"libmbik_{tag}.so")
r = subprocess.run([B.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-o", so] + objs,
                   capture_output=True, text=True)
if r.returncode != 0:
    sys.stderr.write(r.stdout + r.stderr)
    sys.exit(1)
print(so)
