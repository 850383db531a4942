"""CSR_BLOCK against the other CSR kernels on BASELINE configs 1-4 at full size (flushed + back to back).
usage: python tools/sweep_block.py [--configs cfg1,cfg2,cfg3,cfg4] [--out gpurun_out/sweep_block.jsonl]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402

PEAK = 6550.7
MAKE = {"cfg1": lambda: G.laplace2d_5pt(1000, rcm=True), "cfg2": lambda: G.stencil27(100, mode="varied"),
        "cfg3": lambda: G.shell_blocks(1_500_000), "cfg4": lambda: G.rmat_torch(24, 16)}

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="cfg1,cfg2,cfg3,cfg4")
ap.add_argument("--dtypes", default="f64,f32")
ap.add_argument("--out", default="")
a = ap.parse_args()
out = open(a.out, "a") if a.out else None
for cfg in a.configs.split(","):
    rp, ci, v = MAKE[cfg]()
    n, nnz = len(rp) - 1, len(ci)
    for dname in a.dtypes.split(","):
        dt = np.float64 if dname == "f64" else np.float32
        vs = 8 if dname == "f64" else 4
        A = hs.Matrix.from_csr(rp, ci, v.astype(dt), keep_host=False)
        A.set_x(G.vector_x(n, dt))
        bytes_ = nnz * (vs + 4) + (n + 1) * 4 + 2 * n * vs
        cases = [("auto", hs.ALGO_AUTO, {}), ("stream", hs.ALGO_CSR_STREAM, {}), ("vector", hs.ALGO_CSR_VECTOR, {})]
        for cap in (1024, 2048, 4096, 8192):
            cases.append(("block cap=%d" % cap, hs.ALGO_CSR_BLOCK, {hs.OPT_BLOCK_NNZ: cap}))
        for vec in (1, 2, 4, 8):
            cases.append(("block cap=2048 vec=%d" % vec, hs.ALGO_CSR_BLOCK, {hs.OPT_BLOCK_NNZ: 2048, hs.OPT_VEC: vec}))
        for name, algo, opts in cases:
            if cfg == "cfg4" and name == "stream":
                continue
            for k_, val in opts.items():
                A.set_option(k_, val)
            try:
                ms = float(np.median(A.spmv_timed(algo, warmup=3, iters=20, flush_l2=True)))
                b2b = A.spmv_timed_batch(algo, warmup=5, iters=20)
                rec = {"config": cfg, "dtype": dname, "case": name, "ms_flushed": ms, "ms_b2b": b2b, "frac_flushed": bytes_ / ms / 1e6 / PEAK,
                       "frac_b2b": bytes_ / b2b / 1e6 / PEAK, "kernel": A.last_kernel()}
            except hs.HspmvError as e:
                rec = {"config": cfg, "dtype": dname, "case": name, "error": str(e)}
            for k_ in opts:
                A.set_option(k_, 0)
            print(json.dumps(rec), flush=True)
            if out:
                out.write(json.dumps(rec) + "\n")
                out.flush()
        A.close()
