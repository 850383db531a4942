"""Single-process multi-GPU (hspmv_mg_*, csrc/mg.cu) on the BASELINE config-5 stencil: ms per step on N devices of one host thread,
result checked against the closed form for x = 1 and against a numpy product on windows for the non-constant x.
usage: python tools/mg_bench.py [--gpus 8] [--g 400] [--k 1]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--g", type=int, default=400)
ap.add_argument("--k", type=int, default=1)
a = ap.parse_args()
nd = a.gpus or hs.device_count()
g = a.g
n = g ** 3
nnz = (3 * g - 2) ** 3
t0 = time.time()
sizes, _, _ = hs.auto_params(nnz, n, hs.ARCH_B200)
M = hs.MultiGPU.stencil(list(range(nd)), 27, g, np.float64, 1, k=a.k, sup_row_sizes=sizes)
build_s = time.time() - t0
part = M.partition()
x = G.vector_x(n)
M.poison_halos()
M.set_x(x)
M.spmv()
y = M.get_y()
bad = 0
rows = 0
W = 200_000
for lo in sorted({0, n - W} | {int(r) - W // 2 for r in part["row_lo"][1:-1]}):   # both ends + every partition boundary
    lo = max(0, min(n - W, lo))
    rp, ci, v = G.stencil27(g, mode="varied", row_lo=lo, row_hi=lo + W)
    prod = v * x[ci]
    yw = np.add.reduceat(prod, rp[:-1].astype(np.int64))
    sw = np.add.reduceat(np.abs(prod), rp[:-1].astype(np.int64))
    if a.k == 1:
        bad += int(np.count_nonzero(~(np.abs(y[lo:lo + W] - yw) <= 1e-12 * sw)))
        rows += W
if a.k > 1:   # rows are in the original numbering again after get_y: check everything against a second kernel instead
    A = hs.Matrix.stencil(27, g, np.float64, 1)
    A.fill_x_synthetic(0)
    A.spmv(hs.ALGO_CSR_VECTOR)
    y1 = A.get_y()
    A.close()
    bad = int(np.count_nonzero(~(np.abs(y - y1) <= 1e-12 * 39.0)))
    rows = n
ms = M.spmv_timed_batch(5, 20)
per_step = M.spmv_timed(2, 10)
rec = {"what": "hspmv_mg_* single process", "n_dev": nd, "g": g, "k": a.k, "build_s": build_s, "ms_per_step_batch": ms,
       "ms_per_step_synced_min": float(per_step.min()), "gflops": 2.0 * nnz / ms / 1e6, "rows_checked": rows, "rows_outside_bound": bad,
       "nan_rows": int(np.count_nonzero(~np.isfinite(y))), "halo_elems": [int(t) for t in part["halo_elems"]],
       "nnz_per_dev": [int(t) for t in part["nnz"]], "kernel": M.shard_kernel(0)}
print(json.dumps(rec), flush=True)
M.close()
sys.exit(1 if bad or rec["nan_rows"] else 0)
