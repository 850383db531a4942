"""Product-staging flavour (csrk_prod_kernel) against the slot flavour on BASELINE configs 1-3 (+ the 400^3 stencil with --big).
usage: python tools/sweep_prod.py [--configs cfg3,cfg1,cfg2] [--out gpurun_out/sweep_prod.jsonl]"""
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
        "cfg3": lambda: G.shell_blocks(1_500_000), "cfg3b": lambda: G.banded_shell(1_500_000)}
ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="cfg3,cfg1,cfg2")
ap.add_argument("--dtypes", default="f64,f32")
ap.add_argument("--out", default="")
ap.add_argument("--formats", default="csr,csr2")
a = ap.parse_args()
out = open(a.out, "a") if a.out else None
KEYS = {"prod": hs.OPT_PROD, "vec": hs.OPT_VEC, "u": hs.OPT_TAIL, "rpt": hs.OPT_ROWS_PER_TILE, "warps": hs.OPT_WARPS, "stages": hs.OPT_STAGES,
        "cap": hs.OPT_STAGE_NNZ, "ctas": hs.OPT_CTAS_PER_SM}


def emit(rec):
    print(json.dumps(rec), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")
        out.flush()


for cfg in a.configs.split(","):
    rp, ci, v = MAKE[cfg]()
    n, nnz = len(rp) - 1, len(ci)
    mean = nnz / n
    for dname in a.dtypes.split(","):
        dt = np.float64 if dname == "f64" else np.float32
        vs = 8 if dname == "f64" else 4
        bytes_ = nnz * (vs + 4) + (n + 1) * 4 + 2 * n * vs
        for fmt in a.formats.split(","):
            A = hs.Matrix.from_csr(rp, ci, v.astype(dt), keep_host=fmt != "csr")
            if fmt != "csr":
                sizes, _, _ = hs.auto_params(nnz, n, hs.ARCH_B200)
                A.build_csrk(int(fmt[-1]), sizes)
            A.set_x(G.vector_x(n, dt))
            algo = hs.ALGO_CSR_STREAM if fmt == "csr" else hs.ALGO_CSRK
            cases = [{"prod": -1}, {}]
            vecs = sorted({max(1, min(32, 1 << int(np.ceil(np.log2(max(mean / t, 1)))))) for t in (5.0, 10.0, 20.0)})
            for vec in vecs:
                for u in (6, 8, 12):
                    if (vec, u) not in {(1, 8), (2, 8), (4, 8), (8, 8), (16, 8), (32, 8), (2, 12), (4, 12), (8, 12), (4, 6), (8, 6)}:
                        continue
                    for mult in ((1, 2) if fmt == "csr" else (1,)):
                        per_pass = 32 // vec
                        rpt = max(1, int(32.0 * u * mult / (mean * per_pass) + 0.5)) * per_pass
                        for warps, stages in ((16, 2), (12, 3), (16, 3)):
                            c = {"prod": 1, "vec": vec, "u": u, "warps": warps, "stages": stages}
                            if fmt == "csr":
                                c["rpt"] = rpt
                            cases.append(c)
            for c in cases:
                for k_, val in c.items():
                    A.set_option(KEYS[k_], val)
                try:
                    ms = float(np.median(A.spmv_timed(algo, warmup=3, iters=15, flush_l2=True)))
                    b2b = A.spmv_timed_batch(algo, warmup=5, iters=20)
                    rec = {"config": cfg, "dtype": dname, "format": fmt, "opts": c, "ms_flushed": ms, "ms_b2b": b2b,
                           "frac_flushed": bytes_ / ms / 1e6 / PEAK, "frac_b2b": bytes_ / b2b / 1e6 / PEAK, "kernel": A.last_kernel()}
                except hs.HspmvError as e:
                    rec = {"config": cfg, "dtype": dname, "format": fmt, "opts": c, "error": str(e)}
                for k_ in c:
                    A.set_option(KEYS[k_], 0)
                emit(rec)
            A.close()
