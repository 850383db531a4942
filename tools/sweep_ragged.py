"""Kernel-shape sweep on a ragged but banded matrix (BASELINE config 3 stand-in: 5-dof shell blocks).
usage: python tools/sweep_ragged.py [--n 1500000] [--out gpurun_out/sweep_ragged.jsonl]"""
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
ap.add_argument("--n", type=int, default=1_500_000)
ap.add_argument("--out", default="")
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
dt = np.float64 if a.dtype == "f64" else np.float32
rp, ci, v = G.shell_blocks(a.n, dtype=dt)
n, nnz = len(rp) - 1, len(ci)
lens = np.diff(rp)
print("n=%d nnz=%d mean %.1f p50 %d p90 %d p99 %d max %d" % (n, nnz, lens.mean(), np.percentile(lens, 50),
                                                             np.percentile(lens, 90), np.percentile(lens, 99), lens.max()), flush=True)
vs = 8 if a.dtype == "f64" else 4
bytes_ = nnz * (vs + 4) + (n + 1) * 4 + 2 * n * vs
out = open(a.out, "a") if a.out else None


def run(A, algo, label, **opts):
    for k in (hs.OPT_VEC, hs.OPT_CHUNK, hs.OPT_TAIL, hs.OPT_ROWS_PER_TILE, hs.OPT_WARPS, hs.OPT_STAGES):
        A.set_option(k, 0)
    for k, val in opts.items():
        A.set_option(getattr(hs, "OPT_" + k.upper()), val)
    try:
        ms = float(np.median(A.spmv_timed(algo, warmup=3, iters=15)))
    except hs.HspmvError as e:
        print("  %-8s %-44s failed: %s" % (label, opts, e), flush=True)
        return
    rec = {"format": label, "opts": opts, "ms": ms, "gbs": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / 6550.7, "kernel": A.last_kernel()}
    print("  %-8s %-52s %.4f ms %7.1f GB/s (%.3f)  %s" % (label, opts, ms, rec["gbs"], rec["frac"], rec["kernel"]), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")


A = hs.Matrix.from_csr(rp, ci, v, keep_host=False)
A.fill_x_synthetic(0)
run(A, hs.ALGO_CSR_STREAM, "csr")
for vec, tail in ((2, 12), (4, 8), (4, 12), (4, 14), (8, 4), (8, 8), (8, 12), (16, 4), (16, 8)):
    for rpt in (32 // vec, 2 * 32 // vec, 4 * 32 // vec):
        run(A, hs.ALGO_CSR_STREAM, "csr", vec=vec, chunk=-1, tail=tail, rows_per_tile=rpt)
for vec, chunk in ((4, 4), (4, 6), (8, 3), (8, 4)):
    run(A, hs.ALGO_CSR_STREAM, "csr", vec=vec, chunk=chunk)
A.build_csrk(2, [16])
run(A, hs.ALGO_CSRK, "csr2-16")
for vec, tail in ((4, 8), (4, 12), (4, 14), (8, 8), (8, 12)):
    run(A, hs.ALGO_CSRK, "csr2-16", vec=vec, chunk=-1, tail=tail)
A.close()
for s in (4, 8):
    A = hs.Matrix.from_csr(rp, ci, v, keep_host=False)
    A.fill_x_synthetic(0)
    A.build_csrk(2, [s])
    run(A, hs.ALGO_CSRK, "csr2-%d" % s)
    for vec, tail in ((4, 12), (8, 8), (8, 12)):
        run(A, hs.ALGO_CSRK, "csr2-%d" % s, vec=vec, chunk=-1, tail=tail)
    A.close()
