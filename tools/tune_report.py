"""Heuristic vs autotuned kernel shape on the bench matrix (27-point stencil), fp64 and fp32.
usage: python tools/tune_report.py [--g 400] [--out gpurun_out/tune_report.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--g", type=int, default=400)
ap.add_argument("--out", default="")
a = ap.parse_args()
out = open(a.out, "a") if a.out else None
for dt_name, dt, vs in (("f64", np.float64, 8), ("f32", np.float32, 4)):
    A = hs.Matrix.stencil(27, a.g, dt, 0)
    i = A.info()
    A.fill_x_synthetic(0)
    b = i["nnz"] * (vs + 4) + (i["m"] + 1) * 4 + 2 * i["m"] * vs
    t0 = A.spmv_timed_batch(hs.ALGO_CSR_STREAM, warmup=3, iters=10)
    k0 = A.last_kernel()
    t1 = A.autotune(hs.ALGO_CSR_STREAM, effort=2)
    k1 = A.last_kernel()
    t1 = min(t1, A.spmv_timed_batch(hs.ALGO_CSR_STREAM, warmup=3, iters=10))
    rec = {"g": a.g, "dtype": dt_name, "heuristic_ms": t0, "heuristic_gbs": b / t0 / 1e6, "heuristic_gflops": 2 * i["nnz"] / t0 / 1e6,
           "heuristic_kernel": k0, "tuned_ms": t1, "tuned_gbs": b / t1 / 1e6, "tuned_gflops": 2 * i["nnz"] / t1 / 1e6, "tuned_kernel": k1}
    print(json.dumps(rec), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")
    A.close()
