"""Generic text summary of an `ncu --set full --import-source on` capture: key counters of the raw page and, from the source page, the SASS
in blocks of 50 instructions that hold at least 1 % of the executed warp instructions or of the warp-state samples (share, lanes per
instruction, dominant opcodes, dominant stall reasons).
Usage: python profiles/ncu_region_summary.py <capture.ncu-rep> <out.txt> [units, e.g. rays of the launch] [title]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
title = sys.argv[4] if len(sys.argv) > 4 else ""


def page(which):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], check=True, capture_output=True, text=True).stdout)))


raw = page("raw")
hdr, un, vals = raw[0], raw[1], raw[2]
col = {h: i for i, h in enumerate(hdr)}
lines = [title, f"capture: {rep.split('/')[-1]}; kernel: {vals[col['Kernel Name']]}", ""]
for m in ("gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if m in col:
        lines.append(f"{m:70s} = {vals[col[m]]} {un[col[m]]}")
lines.append("")
lines.append("warps stalled per issue-active cycle, by reason:")
for m in sorted((h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h),
                key=lambda h: -float(vals[col[h]].replace(",", "") or 0))[:8]:
    lines.append(f"  {m[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {float(vals[col[m]].replace(',', '')):.2f}")
src = page("source")
shdr, data = src[1], src[2:]
ix = {h: i for i, h in enumerate(shdr)}


def op(r):
    t = r[ix["Source"]].split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
ts = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
lines += ["", f"SASS: {len(data)} instructions, {tot:.4g} executed warp instructions" + (f" = {tot / units * 32:.0f} per 32 units" if units else ""), ""]
for lo in range(0, len(data), 50):
    sub = data[lo:lo + 50]
    w = sum(int(r[ix["Instructions Executed"]]) for r in sub)
    s = sum(int(r[ix["# Samples"]]) for r in sub)
    th = sum(int(r[ix["Thread Instructions Executed"]]) for r in sub)
    if w / tot < 0.01 and s / max(ts, 1) < 0.01:
        continue
    ops, st = {}, {}
    for r in sub:
        ops[op(r)] = ops.get(op(r), 0) + int(r[ix["Instructions Executed"]])
        for h in stalls:
            st[h] = st.get(h, 0) + int(r[ix[h]] or 0)
    lines.append(f"  SASS {lo:5d}-{lo + len(sub) - 1:5d}  {100 * w / tot:5.1f} % of warp instructions  {100 * s / max(ts, 1):5.1f} % of samples  {th / max(w, 1):5.1f} lanes   "
                 + " ".join(k for k, _ in sorted(ops.items(), key=lambda kv: -kv[1])[:5]) + "   | "
                 + ", ".join(f"{k[6:]} {100 * v / max(s, 1):.0f} %" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
