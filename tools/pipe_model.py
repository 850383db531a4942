#!/usr/bin/env python
"""A timeline MODEL (not a measurement) of hspmv_spmv_host's pipeline on config 5, to compare chunk plans before they
can be timed on hardware: x pieces go up on one copy engine, row-chunk kernels run on the compute stream as soon as the
last x piece they reference has landed, y chunks go down on the other copy engine.

Rates from the r2 bench line (profiles/r2_bench_n1_all_configs.json): PCIe 512 MB one way alone 9.2 ms, both ways at
once 10.66 ms (e2e.pcie_floor_ms); kernel 21.9 GB in 3.15 ms (16-warp row-range form) + ~8 us per launch.

    python tools/pipe_model.py          # equal x pieces (r2 measurements) vs pieces cut at the chunks' column ranges (now)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

G = 400
N = G ** 3
HALF_BW = G * G + G + 1
VS = 8
B_ALONE = 512e6 / 9.2e-3          # bytes / s, one direction busy
B_BOTH = 512e6 / 10.66e-3         # each direction, both busy
KERNEL_BPS = 21.9e9 / 3.15e-3
LAUNCH = 8e-6


def simulate(row_lo, x_lo, need):
    C = len(row_lo) - 1
    up = [(x_lo[c + 1] - x_lo[c]) * VS for c in range(C)]
    down = [(row_lo[c + 1] - row_lo[c]) * VS for c in range(C)]
    kern = [LAUNCH + (row_lo[c + 1] - row_lo[c]) / N * 21.9e9 / KERNEL_BPS for c in range(C)]
    t = 0.0
    up_i, up_left = 0, float(up[0])
    x_done = [None] * C
    k_done = [None] * C
    dn_i, dn_left, dn_active = 0, float(down[0]), False
    k_i, k_end = 0, None
    while dn_i < C:
        # start what can start
        if k_end is None and k_i < C and x_done[need[k_i]] is not None and x_done[need[k_i]] <= t:
            k_end = t + kern[k_i]
        dn_active = dn_i < C and k_done[dn_i] is not None and k_done[dn_i] <= t
        up_active = up_i < C
        rate_up = (B_BOTH if dn_active else B_ALONE) if up_active else 0.0
        rate_dn = (B_BOTH if up_active else B_ALONE) if dn_active else 0.0
        # next event
        cand = []
        if up_active:
            cand.append(t + up_left / rate_up if up_left > 0 else t)
        if dn_active:
            cand.append(t + dn_left / rate_dn if dn_left > 0 else t)
        if k_end is not None:
            cand.append(k_end)
        t2 = min(cand)
        dt = t2 - t
        if up_active:
            up_left -= rate_up * dt
        if dn_active:
            dn_left -= rate_dn * dt
        t = t2
        if up_active and up_left <= 1e-6:
            x_done[up_i] = t
            up_i += 1
            up_left = float(up[up_i]) if up_i < C else 0.0
        if k_end is not None and k_end <= t + 1e-15:
            k_done[k_i] = t
            k_i += 1
            k_end = None
        if dn_active and dn_left <= 1e-6:
            dn_i += 1
            dn_left = float(down[dn_i]) if dn_i < C else 0.0
    return t


def old_plan(C=32):
    row_lo = [((N * c // C) + 3) & ~3 for c in range(C + 1)]
    row_lo[C] = N
    x_lo = list(row_lo)
    need = []
    for c in range(C):
        cmax = min(N - 1, row_lo[c + 1] - 1 + HALF_BW)
        p = 0
        while p + 1 < C and x_lo[p + 1] <= cmax:
            p += 1
        need.append(p)
    return row_lo, x_lo, need


def new_plan():
    import heterogeneous_spmv_b200 as hs
    row_lo = hs.pipe_plan_host(N, N, VS)
    cmax = np.minimum(row_lo[1:] - 1 + HALF_BW, N - 1)
    row_lo, x_lo, need = hs.pipe_plan_host(N, N, VS, cmax)
    return [int(v) for v in row_lo], [int(v) for v in x_lo], [int(v) for v in need]


if __name__ == "__main__":
    print("model of one hspmv_spmv_host call on config 5 (512 MB up, 512 MB down); PCIe floor with both directions busy: 10.66 ms")
    a = simulate(*old_plan())
    print("  32 equal chunks, equal x pieces (a banded chunk waits for piece c+1): %.2f ms" % (a * 1e3))
    row_lo, x_lo, need = new_plan()
    b = simulate(row_lo, x_lo, need)
    print("  %d equal chunks, x pieces cut at the chunks' column ranges (waits for piece c): %.2f ms" % (len(row_lo) - 1, b * 1e3))
    for C in (16, 64, 128):
        rl = old_plan(C)[0]
        cm = [min(N - 1, rl[c + 1] - 1 + HALF_BW) for c in range(C)]
        xl = [0]
        for c in range(C):
            xl.append(max(xl[-1], min(N, (cm[c] + 1 + 3) & ~3)))
        xl[C] = N
        nd = []
        for c in range(C):
            p = 0
            while p + 1 < C and xl[p + 1] <= cm[c]:
                p += 1
            nd.append(p)
        print("  what if: %3d equal chunks with cut pieces: %.2f ms" % (C, simulate(rl, xl, nd) * 1e3))
    print("  (measured with the old plan on r2 boxes: 12.4 - 13.1 ms; the model leaves out event / launch latencies between the legs)")
