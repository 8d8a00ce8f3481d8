#!/usr/bin/env python
"""Turn the ncu captures that gpurun brought back (gpurun_out/prof_<kernel>.ncu-rep, launches_*.csv) into the
small, tracked summaries under profiles/: per-kernel key metrics (CSV + markdown) and the launch list with each
kernel's share of the step. Usage: python profiles/summarise.py r1 [kernel ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("FPT_SUMM_OUT") or os.path.join(ROOT, "profiles")     # on the GPU box: a directory under gpurun_out/
BASE = os.path.join(ROOT, "profiles")                                       # committed records to merge into
SRC = os.path.join(ROOT, "gpurun_out")
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_tensor.sum",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
# capture name -> the kernel label bench.py's live profile uses
ALIAS = {"css_perm2": "css_perm", "css_perm3": "css_perm", "css_tridiag_reg": "css_tridiag", "css_mds_codes": "css_mds_large",
         "fet_score_tables": "fet_score_log", "css_k4_umma": "css_k4"}


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    kernels = sys.argv[2:] or [f[5:-8] for f in sorted(os.listdir(SRC)) if f.startswith("prof_") and f.endswith(".ncu-rep")]
    md = ["# ncu summaries, round %s" % tag, "",
          "One `ncu --set full --clock-control none --import-source on` capture per kernel (one launch each), taken by",
          "`gpurun` on a B200 (commands: `profiles/run_%s_final.sh` / `profiles/capture.sh`); raw metric tables in" % tag,
          "`%s_<kernel>_raw.csv`. Times under ncu are serialised and cold-cache: compare shares, not absolutes." % tag, ""]
    traffic = {}
    insts = {}
    for k in kernels:
        rep = os.path.join(SRC, "prof_%s.ncu-rep" % k)
        if not os.path.exists(rep):
            continue
        hdr, units, rows = raw(rep)
        with open(os.path.join(OUT, "%s_%s_raw.csv" % (tag, k)), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            for i, h in enumerate(hdr):
                w.writerow([h, units[i], rows[0][i]])
        name = rows[0][hdr.index("Kernel Name")] if "Kernel Name" in hdr else k
        md += ["## %s (`%s`)" % (k, name), "", "| metric | value | unit |", "|---|---|---|"]
        vals = {}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                vals[key] = (rows[0][i], units[i])
                md.append("| %s | %s | %s |" % (key, rows[0][i], units[i]))
        md.append("")
        try:
            rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
            traffic[ALIAS.get(k, k)] = float(rd[0].replace(",", "")) * UNIT.get(rd[1], 1.0) + float(wr[0].replace(",", "")) * UNIT.get(wr[1], 1.0)
        except Exception:
            pass
        if "smsp__inst_executed.sum" in vals:
            insts[ALIAS.get(k, k)] = float(vals["smsp__inst_executed.sum"][0].replace(",", ""))
    with open(os.path.join(OUT, "%s_ncu_summary.md" % tag), "w") as f:
        f.write("\n".join(md) + "\n")
    os.makedirs(OUT, exist_ok=True)
    tpath = os.path.join(BASE, "ncu_traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old.update(traffic)
    json.dump(old, open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1, sort_keys=True)
    ipath = os.path.join(BASE, "ncu_inst.json")         # warp instructions per launch, for the issue-slot roofline
    old = json.load(open(ipath)) if os.path.exists(ipath) else {}
    old.update(insts)
    json.dump(old, open(os.path.join(OUT, "ncu_inst.json"), "w"), indent=1, sort_keys=True)
    # launch list -> per-kernel totals and shares
    for fn in sorted(os.listdir(SRC)):
        if fn.startswith("launches_%s" % tag) and fn.endswith(".csv"):
            lines = [l for l in open(os.path.join(SRC, fn)) if not l.startswith("==")]
            rows = list(csv.DictReader(io.StringIO("".join(lines))))
            agg = {}
            for r in rows:
                if r.get("Metric Name") != "gpu__time_duration.sum":
                    continue
                v = float(r["Metric Value"].replace(",", ""))
                u = r["Metric Unit"]
                ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
                a = agg.setdefault(r["Kernel Name"], [0, 0.0])
                a[0] += 1
                a[1] += ns
            tot = sum(a[1] for a in agg.values()) or 1.0
            with open(os.path.join(OUT, fn.replace(".csv", "_by_kernel.csv")), "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(["kernel", "launches", "total_us", "share"])
                for kname, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    w.writerow([kname[:120], n, "%.1f" % (ns / 1e3), "%.4f" % (ns / tot)])
            with open(os.path.join(OUT, fn), "w") as f:
                f.write("".join(lines))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
