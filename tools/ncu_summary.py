#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into the few numbers the design is argued from.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/rN_xx_summary.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA (FP64 tensor) pipe %"),
    ("sm__icc_request_hit_rate.pct", "instruction cache hit %"),
    ("gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "GPC instr-cache requests % of peak"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles / issued instr"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "  stall: no instruction"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "  stall: wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "  stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "  stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "  stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "  stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "  stall: barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "  stall: mio throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "  stall: lg throttle"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("== %s ==" % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("\n%s  grid %s block %s" % (d["Kernel Name"][:100], d.get("Grid Size", ""), d.get("Block Size", "")))
        for k, label in KEYS:
            if k in d and d[k] != "":
                print("    %-40s %s %s" % (label, d[k], u.get(k, "")))


if __name__ == "__main__":
    main()
