#!/usr/bin/env python3
"""Per-launch summary of an `ncu --set full` report as JSON (the figures DESIGN.md and bench.py's roofline.traffic quote).
Usage: python tools/ncu_summarize.py gpurun_out/prof_X.ncu-rep [more.ncu-rep ...] > profiles/ncu_summary_X.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    out = {}
    for rep in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, body = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        for r in body:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")
            e = {}
            for k in KEEP:
                if k in col:
                    e[k] = {"value": num(r[col[k]]), "unit": units[col[k]]}
            stalls = {}
            for h, i in col.items():
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    v = num(r[i])
                    if v is not None and v >= 0.2:
                        stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 2)
            e["stalls_per_issue"] = stalls
            rd, wr = e.get("dram__bytes_read.sum"), e.get("dram__bytes_write.sum")
            if rd and wr and rd["value"] is not None and wr["value"] is not None:
                e["dram_bytes_per_launch"] = rd["value"] * SCALE.get(rd["unit"], 1) + wr["value"] * SCALE.get(wr["unit"], 1)
            out.setdefault(name, []).append(e)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
