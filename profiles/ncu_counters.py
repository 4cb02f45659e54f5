"""Extracts the per-launch counters bench.py reports from an ncu capture of the bench command and appends / replaces them in
profiles/ncu_counters.json (what `roofline.executed_*` and `roofline.traffic` in the bench line are read from).

  python profiles/ncu_counters.py <capture.ncu-rep> --kernel 'pathtrace_kernel<1,1,0>' --width 1200 --height 800 --spp 500 \
         --max-depth 50 --n-prims 484 [--n-gpus 1] --source 'profiles/r02a_ncu_pathtrace_c2.txt'

Needs the capture to hold the raw page (--set full plus the smsp__sass_thread_inst_executed_op_{ffma,fmul,fadd}_pred_on.sum
metrics) and the source page (--import-source on): FFMA2, the packed fp32x2 FMA, is not part of `op_ffma`; its thread-level
count comes from the per-opcode sums of the source page and is checked against `op_fp32`."""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_counters.json")


def page(rep, which):
    text = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], check=True, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(text)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", required=True)
    for name in ("width", "height", "spp", "max-depth", "n-prims"):
        ap.add_argument("--" + name, type=int, required=True)
    ap.add_argument("--n-gpus", type=int, default=1)
    ap.add_argument("--rays", type=int, default=0, help="ray segments of the captured launch (rt3_stats.rays): lets bench.py scale the counters to a rank's share at N > 1")
    ap.add_argument("--tests", type=int, default=0, help="ray-primitive tests the captured launch is credited with (rt3_stats.sphere_tests + triangle_tests); 0 = rays x primitives")
    ap.add_argument("--source", required=True, help="the committed summary this capture is described in")
    a = ap.parse_args()

    raw = page(a.rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    col = {h: i for i, h in enumerate(hdr)}

    def metric(name, scale_units=True):
        v = float(vals[col[name]].replace(",", ""))
        u = units[col[name]]
        if scale_units:
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u, 1.0)
        return v

    src = page(a.rep, "source")
    shdr, rows = src[1], src[2:]
    i_src, i_inst, i_thr = shdr.index("Source"), shdr.index("Instructions Executed"), shdr.index("Predicated-On Thread Instructions Executed")
    per_op = {}
    for r in rows:
        tok = r[i_src].split()
        op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
        w, t = per_op.get(op, (0, 0))
        per_op[op] = (w + int(r[i_inst]), t + int(r[i_thr]))
    total_warp = sum(w for w, _ in per_op.values())
    row = {
        "kernel": a.kernel, "source": a.source,
        "workload": {"width": a.width, "height": a.height, "spp": a.spp, "max_depth": a.max_depth, "n_prims": a.n_prims, "n_gpus": a.n_gpus},
        "duration_ms_under_ncu": metric("gpu__time_duration.sum"), "rays_per_launch": a.rays, "tests_per_launch": a.tests or a.rays * a.n_prims,
        "ffma_thread_inst": metric("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"),
        "fmul_thread_inst": metric("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"),
        "fadd_thread_inst": metric("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"),
        "fp32_thread_inst": metric("smsp__sass_thread_inst_executed_op_fp32_pred_on.sum"),
        "ffma2_thread_inst": per_op.get("FFMA2", (0, 0))[1],
        "ffma2_warp_inst": per_op.get("FFMA2", (0, 0))[0],
        "warp_inst": metric("smsp__inst_executed.sum"),
        "dram_bytes_read": metric("dram__bytes_read.sum"), "dram_bytes_write": metric("dram__bytes_write.sum"),
        "pipe_fma_cycles_active_pct": round(metric("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"), 2),
        "issue_active_pct": round(metric("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
        "registers_per_thread": int(metric("launch__registers_per_thread")),
        "opcode_share_of_warp_inst": {op: round(w / total_warp, 4) for op, (w, _) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:14]},
    }
    table = {"captures": []}
    if os.path.exists(OUT):
        table = json.load(open(OUT))
    table["captures"] = [r for r in table["captures"] if not (r["kernel"] == row["kernel"] and r["workload"] == row["workload"])] + [row]
    json.dump(table, open(OUT, "w"), indent=1)
    flop = 2 * row["ffma_thread_inst"] + 4 * row["ffma2_thread_inst"] + row["fmul_thread_inst"] + row["fadd_thread_inst"]  # FFMA2 = two FMAs
    print(json.dumps(row, indent=1))
    print(f"executed FP32: {flop:.4g} FLOP per launch = {flop / row['duration_ms_under_ncu'] / 1e9:.2f} TFLOP/s at the captured duration")


if __name__ == "__main__":
    main()
