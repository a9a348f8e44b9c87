"""Per-function / per-statement profile of the solve kernel from an ncu report.

Joins the ncu source page (per-SASS-instruction executed counts and stall samples) with `nvdisasm -gi` line info
(which carries the inline chain) of the SAME libmbik.so build.

    python profiles/line_profile.py <report.ncu-rep> <libmbik.so> <mangled-kernel-substring> [--top N]

Output: (1) share of executed warp-instructions and stall samples per device function (innermost function of
mbik_kernel.cu / mbik_math.cuh on the inline chain, and the outermost one below the kernel body = "stage"),
(2) per statement of the kernel body.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "many_bone_ik_b200", "csrc")


def function_table(path):
    """[(first_line, name)] of function definitions in a source file (crude, good enough for these files)."""
    out = []
    for i, ln in enumerate(open(path), 1):
        if not re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:MBIK_HD|__device__|__global__)", ln):
            continue
        sig = ln.split("{")[0]
        names = [n for n in re.findall(r"(\w+)\s*\(", sig) if n not in ("__launch_bounds__", "defined")]
        if names and not sig.rstrip().endswith(";"):
            out.append((i, names[-1]))
    return out


BODY = "mbik_kernel_body.cuh"
FT = {f: function_table(os.path.join(CSRC, f)) for f in (BODY, "mbik_math.cuh")}


def func_of(fname, line):
    name = "?"
    for start, n in FT.get(fname, []):
        if start <= line:
            name = n
        else:
            break
    return name


# ---- address -> inline chain from nvdisasm -gi
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = ""
for f in sorted(os.listdir(tmp)):
    if "mbik_kernel" in f and f.endswith(".cubin"):
        d = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in d:
            dis = d
            break
addr_chain = {}
chain = []
in_kernel = False
pending = []
for ln in dis.splitlines():
    if ln.startswith(".text."):
        in_kernel = kname in ln
        continue
    if not in_kernel:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*)", ln)
    if m:
        if pending:
            chain = pending
            pending = []
        addr_chain[int(m.group(1), 16)] = chain

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def addr(r):
    a = r[idx["Address"]]
    return int(a, 16) if a.startswith("0x") else int(a)


base = addr(data[0])
scalar_caller = collections.Counter()  # scalar r_mul / r_add / r_sub instructions by the function that calls them
SCALAR = ("r_mul", "r_add", "r_sub")
inner = collections.Counter()
stage = collections.Counter()
stmt = collections.Counter()
inner_s = collections.Counter()
stage_s = collections.Counter()
stmt_s = collections.Counter()
tot = tots = 0
kernel_start = [s for s, n in FT[BODY] if n == "solve_body"][0]
kernel_end = min(s for s, n in FT[BODY] if s > kernel_start)
for r in data:
    try:
        ie = int(r[idx["Instructions Executed"]])
        sm = int(r[idx["# Samples"]])
    except Exception:
        continue
    ch = addr_chain.get(addr(r) - base, [])
    tot += ie
    tots += sm
    ours = [(f, l) for f, l in ch if f in FT]
    if not ours:
        inner["(no line info)"] += ie
        inner_s["(no line info)"] += sm
        continue
    fi = func_of(*ours[0])
    if fi in SCALAR:
        callers = [func_of(*x) for x in ours[1:]]
        callers = [c for c in callers if c not in SCALAR]
        scalar_caller[(callers[0] if callers else "?") + (" < " + callers[1] if len(callers) > 1 else "")] += ie
    inner[fi] += ie
    inner_s[fi] += sm
    # the chain is innermost first; the last entry is the kernel body statement
    body = [(f, l) for f, l in ours if f == BODY and kernel_start <= l < kernel_end]
    if body:
        stmt[body[-1][1]] += ie
        stmt_s[body[-1][1]] += sm
    below = [x for x in ours if x not in body]
    st = func_of(*below[-1]) if below else "(kernel body)"
    stage[st] += ie
    stage_s[st] += sm

print(f"total warp-instructions {tot}, stall samples {tots}")
print("\nby stage (outermost device function called from the kernel body): share of executed instructions / of stall samples")
for k, v in stage.most_common(top):
    print(f"  {k:28s} {100.0 * v / tot:6.2f}%  {100.0 * stage_s[k] / max(tots, 1):6.2f}%")
if "--scalar-callers" in sys.argv:
    print("\nscalar FMUL / FADD (r_mul, r_add, r_sub) by calling function < its caller: share of ALL executed instructions")
    for k, v in scalar_caller.most_common(top):
        print(f"  {k:60s} {100.0 * v / tot:6.2f}%")
print("\nby innermost function:")
for k, v in inner.most_common(top):
    print(f"  {k:28s} {100.0 * v / tot:6.2f}%  {100.0 * inner_s[k] / max(tots, 1):6.2f}%")
print("\nby kernel-body statement (mbik_kernel_body.cuh line):")
src = open(os.path.join(CSRC, BODY)).read().splitlines()
for k, v in stmt.most_common(top):
    print(f"  {k:5d} {100.0 * v / tot:6.2f}%  {100.0 * stmt_s[k] / max(tots, 1):6.2f}%  {src[k - 1].strip()[:90]}")
