#!/usr/bin/env python3
"""Key figures of `ncu --page raw --csv` exports (one line per kernel launch): python tools/ncu_csv_summary.py a_raw.csv ..."""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed_op_shared_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in body:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")
        e = {"kernel": name}
        for k in KEEP:
            if k in col:
                v = num(r[col[k]])
                if v is not None:
                    e[k] = [v, units[col[k]]]
        stalls = {}
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                v = num(r[i])
                if v is not None and v >= 0.1:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(v, 2)
        e["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        rd, wr = e.get("dram__bytes_read.sum"), e.get("dram__bytes_write.sum")
        if rd and wr:
            e["dram_bytes_per_launch"] = rd[0] * SCALE.get(rd[1], 1) + wr[0] * SCALE.get(wr[1], 1)
        out.append(e)
    return out


if __name__ == "__main__":
    res = {p: summarize(p) for p in sys.argv[1:]}
    json.dump(res, sys.stdout, indent=1)
