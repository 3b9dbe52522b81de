#!/usr/bin/env python
"""Key counters of an `ncu --set full` report as a small JSON (what profiles/*.json hold):
    ncu -i rep.ncu-rep --page raw --csv > rep.raw.csv ; python tools/ncu_summary.py rep.raw.csv > profiles/r02_ncu_<kernel>.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
data = rows[2] if len(rows) > 2 else None
col = {h: i for i, h in enumerate(hdr)}


def get(name, scale=1.0):
    if name not in col or data is None:
        return None
    try:
        v = float(data[col[name]].replace(",", ""))
    except ValueError:
        return data[col[name]]
    u = units[col[name]]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
    return round(v * mult * scale, 3)


out = {
    "kernel": data[col["Kernel Name"]] if data else None,
    "grid_x_block": "%s x %s" % (data[col["Grid Size"]], data[col["Block Size"]]) if data else None,
    "gpu_time_us": get("gpu__time_duration.sum"),
    "dram_bytes_read": get("dram__bytes_read.sum"),
    "dram_bytes_write": get("dram__bytes_write.sum"),
    "lts_sector_hit_rate_pct": get("lts__t_sector_hit_rate.pct"),
    "sm_pipe_tensor_cycles_active_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    or get("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    "smsp_issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "dram_throughput_pct": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "sm_throughput_pct": get("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    "registers_per_thread": get("launch__registers_per_thread"),
    "shared_mem_per_block_bytes": get("launch__shared_mem_per_block_dynamic"),
    "achieved_occupancy_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"),
}
if out["dram_bytes_read"] is not None and out["dram_bytes_write"] is not None:
    out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    if out["gpu_time_us"]:
        out["dram_GBps"] = round(out["dram_bytes_per_launch"] / out["gpu_time_us"] / 1e3, 1)
print(json.dumps(out, indent=1))
