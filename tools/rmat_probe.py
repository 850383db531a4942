"""BASELINE config 4 (R-MAT scale 24, edge factor 16) generated on the device with torch:
parity of every CSR kernel against the oracle + timings.  usage: python tools/rmat_probe.py [--scale 24] [--out f.jsonl]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import heterogeneous_spmv_b200 as hs  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402
import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--out", default="")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--only", default="", help="run a single algo a few times (for ncu): vector|merge|stream")
a = ap.parse_args()
os.environ.setdefault("OMP_SCHEDULE", "static")
t0 = time.time()
rp, ci, v = G.rmat_torch(a.scale, 16)
import torch  # noqa: E402
torch.cuda.empty_cache()
n, nnz = len(rp) - 1, len(ci)
lens = np.diff(rp)
print("rmat scale %d: n=%d nnz=%d mean %.1f p90 %d p99 %d max %d, generated in %.1f s" % (
    a.scale, n, nnz, lens.mean(), np.percentile(lens, 90), np.percentile(lens, 99), lens.max(), time.time() - t0), flush=True)
b = nnz * 12 + (n + 1) * 4 + 2 * n * 8
A = hs.Matrix.from_csr(rp, ci, v, keep_host=False)
x = G.vector_x(n, np.float64)
A.set_x(x)
out = open(a.out, "a") if a.out else None
if a.only:
    algo = {"vector": hs.ALGO_CSR_VECTOR, "merge": hs.ALGO_CSR_MERGE, "stream": hs.ALGO_CSR_STREAM}[a.only]
    ms = A.spmv_timed_batch(algo, warmup=2, iters=3)
    print(a.only, ms, A.last_kernel())
    sys.exit(0)
y_ref = O.csr_spmv(rp, ci, v, x, parallel=True)
scale = O.abs_rowsum(rp, ci, v, x)
for name, algo, opts in (("auto", hs.ALGO_AUTO, {}), ("csr_vector", hs.ALGO_CSR_VECTOR, {}),
                         ("csr_vector + L2 window on x", hs.ALGO_CSR_VECTOR, {hs.OPT_L2_PERSIST_X: 1}),
                         ("csr_merge", hs.ALGO_CSR_MERGE, {}), ("csr_merge tile 8192", hs.ALGO_CSR_MERGE, {hs.OPT_MERGE_TILE: 8192}),
                         ("csr_stream", hs.ALGO_CSR_STREAM, {})):
    for k in (hs.OPT_L2_PERSIST_X, hs.OPT_MERGE_TILE):
        A.set_option(k, 0)
    for k, val in opts.items():
        A.set_option(k, val)
    A.set_y(np.full(n, np.nan))
    A.spmv(algo)
    err = np.abs(A.get_y() - y_ref)
    bad = int(np.count_nonzero(err > 1e-12 * scale))
    ms = A.spmv_timed_batch(algo, warmup=3, iters=a.iters)
    rec = {"config": "cfg4 rmat scale %d ef 16 (torch generator)" % a.scale, "n": n, "nnz": nnz, "algo": name, "rows_outside_bound": bad,
           "ms_back_to_back": ms, "gbs": b / ms / 1e6, "frac": b / ms / 1e6 / 6550.7, "gflops": 2 * nnz / ms / 1e6, "kernel": A.last_kernel()}
    print("  %-30s bad=%d %8.3f ms %7.1f GB/s (%.3f) %6.1f GF/s | %s" % (name, bad, ms, rec["gbs"], rec["frac"], rec["gflops"], rec["kernel"]), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")
