#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of lines that get committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# %s  (ncu --set full --clock-control none; per launch)" % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("\n## %s   grid %s block %s" % (d.get("Kernel Name", "?"), d.get("Grid Size", "?"), d.get("Block Size", "?")))
        for k in KEYS:
            if k in d and d[k] != "":
                print("%-78s %-12s %s" % (k, u.get(k, ""), d[k]))
        print("# warp stall reasons (warps per issue-active cycle)")
        st = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(d[h].replace(",", ""))
                except ValueError:
                    continue
                if v >= 0.05:
                    st.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        for v, n in sorted(st, reverse=True):
            print("    %-28s %.3f" % (n, v))


if __name__ == "__main__":
    main()
