#!/usr/bin/env python
"""Summarise one `ncu --set full` report into a small text file for profiles/ (the .ncu-rep stays in gpurun_out/).

  python tools/profile_summary.py gpurun_out/x.ncu-rep k_canon_w2 [units-per-launch] > profiles/r01_x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_sample_count", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
]


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    iK = hdr.index("Kernel Name")
    print("# source: %s (ncu --set full --clock-control none)" % rep)
    for r in rows[2:]:
        if pat not in r[iK]:
            continue
        print("\n## %s" % r[iK])
        for k in KEYS:
            if k in hdr:
                print("%-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        if per and "smsp__inst_executed.sum" in hdr:
            print("%-70s %.1f" % ("warp instructions per unit (%g units/launch)" % per, float(r[hdr.index("smsp__inst_executed.sum")]) / per))
        if "dram__bytes_read.sum" in hdr:
            def val(k):
                v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            print("%-70s %.0f byte" % ("dram traffic (read + write) per launch", val("dram__bytes_read.sum") + val("dram__bytes_write.sum")))


if __name__ == "__main__":
    main()
