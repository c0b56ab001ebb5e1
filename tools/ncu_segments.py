#!/usr/bin/env python3
"""Per-function split of a kernel's executed instructions from an ncu report (--set full,
--import-source on): the SASS page is cut at RET instructions (entry first, then the out-of-line
device functions in layout order -- compare with tools/sass_size.py for their names).

    python tools/ncu_segments.py report.ncu-rep <kernel regex> [launch index]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Address" in r)
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[rows.index(hdr) + 1:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        if "Address" in r and data:
            break          # next kernel
        continue
    data.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]),
                 int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])))
tot = sum(d[1] for d in data)
seg, cur = [], []
for d in data:
    cur.append(d)
    if d[0].startswith("RET"):
        seg.append(cur)
        cur = []
if cur:
    seg.append(cur)
print("%d SASS instructions, %d warp instructions executed" % (len(data), tot))
print("%8s %8s %8s %10s %10s" % ("offset", "size", "exec %", "thr/instr", "entries"))
off = 0
for s in seg:
    ie = sum(d[1] for d in s)
    te = sum(d[2] for d in s)
    print("%8d %8d %8.1f %10.1f %10d" % (off, len(s), 100.0 * ie / max(tot, 1), te / max(ie, 1), s[0][1]))
    off += len(s)
