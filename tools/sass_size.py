#!/usr/bin/env python3
"""Code footprint of the engine's kernels: instructions per kernel and per device function inside it.

    python tools/sass_size.py [file.o | lib.so] [kernel name filter]

A kernel's footprint (entry + the out-of-line device functions it calls) is what has to live in
the SM's 32 KB instruction cache; the samplers are sized against that.
"""
import glob
import os
import re
import subprocess
import sys
import tempfile

path = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "bayeslogit_b200/lib/libbayeslogit_b200.so")
flt = sys.argv[2] if len(sys.argv) > 2 else ""
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", path], cwd=tmp, capture_output=True)
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        dis = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
        kernel, fn, sizes = None, None, {}
        for line in dis.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                kernel, fn = m.group(1), "(entry)"
                sizes[kernel] = {fn: 0}
                continue
            m = re.match(r"(\$\S+):", line)
            if m and kernel:
                fn = m.group(1).split("$")[-1]
                sizes[kernel][fn] = 0
                continue
            if kernel and re.match(r"\s+/\*[0-9a-f]{4}\*/\s+\S", line):
                sizes[kernel][fn] += 1
        for k, d in sorted(sizes.items(), key=lambda kv: -sum(kv[1].values())):
            name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
            if flt not in name:
                continue
            tot = sum(d.values())
            print("%6d instr %6.1f KB  %s" % (tot, tot * 16 / 1024, name[:110]))
            if flt:
                for f, n in d.items():   # layout order
                    fname = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
                    print("        %6d  %s" % (n, fname[:100]))
