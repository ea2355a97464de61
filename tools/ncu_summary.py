#!/usr/bin/env python
"""Summarise one `ncu --set full` report (.ncu-rep) into the handful of numbers DESIGN.md and
bench.py quote: duration, DRAM bytes, instruction count, issue utilisation, pipe utilisation,
occupancy limits and the top stall reasons.  Usage: ncu_summary.py report.ncu-rep [out.json]"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "launch__occupancy_limit_registers": "occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    result = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                d[name] = "%s %s" % (r[i], units[i])
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        result.append(d)
    text = json.dumps(result, indent=1)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
