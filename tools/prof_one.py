"""Runs ONE kernel configuration a few times (for ncu).  python tools/prof_one.py --algo stream --g 200 --opts vec=2,warps=16,stages=2,rpt=16"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs  # noqa: E402

OPT = {"vec": hs.OPT_VEC, "warps": hs.OPT_WARPS, "stages": hs.OPT_STAGES, "cap": hs.OPT_STAGE_NNZ,
       "ctas": hs.OPT_CTAS_PER_SM, "rpt": hs.OPT_ROWS_PER_TILE, "l2": hs.OPT_L2_PERSIST_X, "tile": hs.OPT_MERGE_TILE}
ALGO = {"vector": hs.ALGO_CSR_VECTOR, "stream": hs.ALGO_CSR_STREAM, "merge": hs.ALGO_CSR_MERGE, "csrk": hs.ALGO_CSRK, "auto": hs.ALGO_AUTO}

ap = argparse.ArgumentParser()
ap.add_argument("--algo", default="stream")
ap.add_argument("--g", type=int, default=200)
ap.add_argument("--kind", type=int, default=27)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--opts", default="")
ap.add_argument("--k", type=int, default=0)
ap.add_argument("--sizes", default="32,8")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--hub", action="store_true", help="hspmv_relabel_hub_columns before timing (power-law matrices)")
ap.add_argument("--matrix", default="stencil", help="stencil (with --g/--kind) | cfg1 | cfg3 | cfg4 (BASELINE configs at full size)")
a = ap.parse_args()
dt = np.float64 if a.dtype == "f64" else np.float32
OPT.update({"chunk": hs.OPT_CHUNK, "tail": hs.OPT_TAIL, "block": hs.OPT_BLOCK_NNZ})
if a.matrix == "stencil":
    A = hs.Matrix.stencil(a.kind, a.g, dt, 0)
else:
    from heterogeneous_spmv_b200 import generators as G
    rp, ci, v = {"cfg1": lambda: G.laplace2d_5pt(1000, rcm=True), "cfg3": lambda: G.shell_blocks(1_500_000),
                 "cfg4": lambda: G.rmat_torch(24, 16)}[a.matrix]()
    A = hs.Matrix.from_csr(rp, ci, v.astype(dt), keep_host=a.k >= 2)
if a.k >= 2:
    A.build_csrk(a.k, [int(t) for t in a.sizes.split(",")])
if a.hub:
    A.relabel_hub_columns()
A.fill_x_synthetic(0)
for kv in filter(None, a.opts.split(",")):
    k, v = kv.split("=")
    A.set_option(OPT[k], int(v))
ms = A.spmv_timed(ALGO[a.algo], warmup=2, iters=a.iters)
print(A.last_kernel(), "median %.4f ms" % float(np.median(ms)))
