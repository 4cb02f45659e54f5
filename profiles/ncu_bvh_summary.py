"""Text summary of an `ncu --set full --import-source on` capture of a hierarchy-traversal kernel (pathtrace_kernel<1,0,1,BIN>):
key counters, lanes per instruction, the larger runs of instructions with a common execution count (node visit, leaf test, stack
handling) with their lane counts, opcode shares and warp-state samples. Reads the report with `ncu -i`; no GPU needed.
Usage: python profiles/ncu_bvh_summary.py <capture.ncu-rep> <out.txt> ["header line" ...]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, out, notes = sys.argv[1], sys.argv[2], sys.argv[3:]


def page(name):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], check=True, capture_output=True, text=True).stdout)))


raw = page("raw")
hdr, units, row = raw[0], raw[1], raw[2]
lines = list(notes)
lines.append("kernel: " + row[hdr.index("Kernel Name")])
for m in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
          "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
          "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct"]:
    if m in hdr:
        i = hdr.index(m)
        lines.append("%s [%s] = %s" % (m, units[i], row[i]))

src = page("source")
sh, data = src[1] if src[0][0] != "Address" and "Source" in src[1] else src[0], None
for k, r in enumerate(src):
    if "Source" in r and "Instructions Executed" in r:
        sh, data = r, src[k + 1:]
        break
isrc, ie, it, isamp = sh.index("Source"), sh.index("Instructions Executed"), sh.index("Thread Instructions Executed"), sh.index("# Samples")


def op(r):
    t = r[isrc].split()
    t = t[1] if t[0].startswith("@") else t[0]
    return t.split(".")[0]


data = [r for r in data if len(r) > it and r[ie].isdigit()]
total = sum(int(r[ie]) for r in data)
lines.append("")
lines.append("the larger runs of consecutive instructions with a common execution count (source page):")
runs, k = [], 0
while k < len(data):
    j = k
    while j + 1 < len(data) and data[j + 1][ie] == data[k][ie]:
        j += 1
    n, ex = j - k + 1, int(data[k][ie])
    if ex and n >= 4 and n * ex >= 0.025 * total:
        th = sum(int(r[it]) for r in data[k:j + 1])
        ops = Counter(op(r) for r in data[k:j + 1]).most_common(5)
        runs.append("  %4d instructions  %5.1f %% of warp instructions  %5.1f lanes   %s" % (n, 100.0 * n * ex / total, th / (n * ex), " ".join("%sx%d" % o for o in ops)))
    k = j + 1
lines += runs
ops = Counter()
for r in data:
    ops[op(r)] += int(r[ie])
lines.append("")
lines.append("opcode share of warp instructions: " + ", ".join("%s %.1f %%" % (o, 100.0 * n / total) for o, n in ops.most_common(14)))
stall_cols = [(h, k) for k, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
if stall_cols:
    tot = Counter()
    for r in data:
        for h, k in stall_cols:
            if k < len(r) and r[k].isdigit():
                tot[h] += int(r[k])
    s = sum(tot.values())
    if s:
        lines.append("warp-state samples: " + ", ".join("%s %.1f %%" % (h, 100.0 * n / s) for h, n in tot.most_common(8)))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
