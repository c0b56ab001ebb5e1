#!/usr/bin/env python3
"""Extract the per-kernel counters bench.py quotes (profiles/ncu_counters.json) from an
`ncu --set full` report of `bench.py` at its default size.

    python tools/ncu_counters.py gpurun_out/x.ncu-rep key:match=units [key:match=units ...] > profiles/ncu_counters.json

units = work items (draws) the captured launch processed; the first launch whose name contains
`match` is used and reported under `key` (key alone: match = key).  DRAM bytes are reported per unit so that bench.py can scale
them to the launch it timed.
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
want = {}
for a in sys.argv[2:]:
    km, units = a.split("=")
    key, _, match = km.partition(":")
    want[key] = (match or key, units)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
res = {}


UNITS = dict(zip(hdr, rows[1]))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3,
         "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}


def num(d, k):
    """metric value; bytes in bytes, durations in ms"""
    try:
        return float(d[k].replace(",", "")) * SCALE.get(UNITS.get(k, ""), 1.0)
    except (KeyError, ValueError):
        return None


for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k, (match, units) in want.items():
        if match in d["Kernel Name"] and k not in res:
            units = float(units)
            rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
            res[k] = {
                "source": rep.split("/")[-1] + " (ncu --set full --clock-control none)",
                "units_in_captured_launch": units,
                "duration_ms": num(d, "gpu__time_duration.sum"),
                "dram_bytes_per_unit": (rd + wr) / units,
                "issue_slots_busy_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fp64_pipe_pct": num(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": num(d, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "alu_pipe_pct": num(d, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                "xu_pipe_pct": num(d, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "lsu_pipe_pct": num(d, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                "achieved_occupancy_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "icache_hit_pct": num(d, "sm__icc_request_hit_rate.pct"),
                "active_threads_per_warp_instr": num(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                "warp_instr_per_unit": num(d, "smsp__inst_executed.sum") / units,
                "registers_per_thread": num(d, "launch__registers_per_thread"),
            }
json.dump(res, sys.stdout, indent=1)
print()
