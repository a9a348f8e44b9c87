"""Join an ncu source-page CSV (per-SASS-instruction counts) with nvdisasm -g line info to get a per-source-line
(and per-inlined-function) profile.
    python profiles/line_profile.py <report.ncu-rep> <kernel.disasm from `nvdisasm -g -c cubin`> <kernel-name-substring>
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, disasm, kname = sys.argv[1], sys.argv[2], sys.argv[3]
# ---- address -> (file, line, inline chain) from nvdisasm
addr_line = {}
cur = None
in_kernel = False
for ln in open(disasm, errors="replace"):
    if ln.startswith(".text."):
        in_kernel = kname in ln
        continue
    if not in_kernel:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)), m.group(3).strip())
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*)", ln)
    if m and cur:
        addr_line[int(m.group(1), 16)] = cur
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][idx["Address"]], 16) if data[0][idx["Address"]].startswith("0x") else int(data[0][idx["Address"]])
per_line = collections.Counter()
per_line_samples = collections.Counter()
tot = tots = 0
for r in data:
    a = r[idx["Address"]]
    a = int(a, 16) if a.startswith("0x") else int(a)
    off = a - base
    try:
        ie = int(r[idx["Instructions Executed"]])
        sm = int(r[idx["# Samples"]])
    except Exception:
        continue
    key = addr_line.get(off, ("?", 0, ""))
    per_line[(key[0], key[1])] += ie
    per_line_samples[(key[0], key[1])] += sm
    tot += ie
    tots += sm
print(f"total warp-instructions {tot}, samples {tots}, mapped lines {len(per_line)}")
print("top source lines by executed warp-instructions (share exec, share samples):")
for k, v in per_line.most_common(45):
    print(f"  {k[0]}:{k[1]:<5d} {100.0 * v / tot:5.2f}%  {100.0 * per_line_samples[k] / max(tots, 1):5.2f}%")
