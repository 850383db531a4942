"""Tile / grid shape sweep on the two small BASELINE configurations (1: 5-pt 1000^2 RCM, 2: 27-pt 100^3),
L2 flushed between launches.  usage: python tools/sweep_small.py [--out gpurun_out/sweep_small.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
a = ap.parse_args()
out = open(a.out, "a") if a.out else None
KEYS = ("vec", "chunk", "tail", "rows_per_tile", "warps", "stages", "ctas_per_sm")
for name, (rp, ci, v), grid in (
        ("cfg1", G.laplace2d_5pt(1000, rcm=True), [dict(rows_per_tile=r, warps=w, stages=s) for r in (32, 64, 96, 160) for w in (8, 16) for s in (2, 4)]),
        ("cfg2", G.stencil27(100), [dict(rows_per_tile=r, warps=w, stages=s) for r in (16, 32) for w in (8, 12, 16) for s in (2, 3)])):
    n, nnz = len(rp) - 1, len(ci)
    bytes_ = nnz * 12 + (n + 1) * 4 + 2 * n * 8
    A = hs.Matrix.from_csr(rp, ci, v, keep_host=False)
    A.fill_x_synthetic(0)
    for opts in [dict()] + grid + [dict(ctas_per_sm=2, warps=8, rows_per_tile=grid[0]["rows_per_tile"]), dict(ctas_per_sm=2, warps=8, rows_per_tile=grid[-1]["rows_per_tile"])]:
        for k in KEYS:
            A.set_option(getattr(hs, "OPT_" + k.upper()), 0)
        for k, val in opts.items():
            A.set_option(getattr(hs, "OPT_" + k.upper()), val)
        try:
            ms = A.spmv_timed(hs.ALGO_CSR_STREAM, warmup=3, iters=25, flush_l2=True)
        except hs.HspmvError as e:
            print(name, opts, "failed", e, flush=True)
            continue
        med, mn = float(np.median(ms)), float(ms.min())
        rec = {"config": name, "opts": opts, "ms_median": med, "ms_min": mn, "frac": bytes_ / med / 1e6 / 6550.7, "kernel": A.last_kernel()}
        print("%s %-52s median %.4f min %.4f ms (%.3f) %s" % (name, opts, med, mn, rec["frac"], rec["kernel"]), flush=True)
        if out:
            out.write(json.dumps(rec) + "\n")
    # launch floor: the same kernel on one tile's worth of rows
    for k in KEYS:
        A.set_option(getattr(hs, "OPT_" + k.upper()), 0)
    A.close()
