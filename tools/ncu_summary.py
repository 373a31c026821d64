#!/usr/bin/env python
"""Print the metrics we care about from an .ncu-rep (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more metric substrings...]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "lts__t_bytes.sum",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("---- %s  grid=%s block=%s" % (r[idx["Kernel Name"]][:70], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
        for h in hdr:
            if any(h.startswith(w) for w in WANT) or any(e in h for e in extra):
                print("  %-75s %s %s" % (h, r[idx[h]], units[idx[h]]))


if __name__ == "__main__":
    main()
