"""Summarise an .ncu-rep (full set, --import-source on) into a small text file that can be committed.
    python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/r1_xxx.txt "title"
"""
import collections
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")


def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


raw = page("raw")
hdr, units, vals = raw[0], raw[1], raw[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
lines = [f"# {title}", f"# source: {rep} (ncu --set full --clock-control none --import-source on)", ""]
for k in keys:
    if k in m:
        lines.append(f"{k:72s} {m[k][0]} {m[k][1]}")
src = page("source")
sh = src[1]
idx = {h: i for i, h in enumerate(sh)}
data = src[2:]
stall_cols = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
n_samples = inst = 0
ops = collections.Counter()
for r in data:
    try:
        n_samples += int(r[idx["# Samples"]])
        ie = int(r[idx["Instructions Executed"]])
    except Exception:
        continue
    inst += ie
    for h in stall_cols:
        try:
            tot[h] += int(r[idx[h]])
        except Exception:
            pass
    s = r[idx["Source"]].split()
    if s:
        op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
        ops[op.split(".")[0]] += ie
lines += ["", f"SASS instructions (static): {len(data)}   warp-instructions executed: {inst}   stall samples: {n_samples}", "", "warp stall sampling (all samples):"]
for h, v in tot.most_common():
    if v:
        lines.append(f"  {h:28s} {100.0 * v / max(n_samples, 1):5.1f}%")
lines += ["", "opcode mix (share of executed warp-instructions):"]
for op, v in ops.most_common(16):
    lines.append(f"  {op:10s} {100.0 * v / max(inst, 1):5.1f}%")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
