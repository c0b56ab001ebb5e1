#!/usr/bin/env python3
"""Host<->device copy bandwidth of the box with pinned buffers, alone and both directions at once:
the floor under bench.py's `e2e` (which moves 16 B in and 8 B out per draw)."""
import json
import time

import torch

n = 1 << 27                                  # 1 GiB of doubles
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


gb = n * 8 / 1e9
t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"h2d_GBs": gb / t_h2d, "d2h_GBs": gb / t_d2h,
                  "both_GBs_each": gb / t_both, "note": "1 GiB pinned, torch copies on two streams"}))
