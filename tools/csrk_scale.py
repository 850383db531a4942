"""CSR-k at BASELINE config-5 scale on one GPU: generate the 27-point stencil in HBM, build
Band-k on the host (download -> bandk -> reorderA -> upload), time CSR-2 / CSR-3.
usage: python tools/csrk_scale.py [--g 400] [--k 2,3]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--g", type=int, default=400)
ap.add_argument("--k", default="2,3")
ap.add_argument("--out", default="")
ap.add_argument("--host-build", action="store_true", help="HSPMV_OPT_BANDK_HOST: whole Band-k build on the host")
a = ap.parse_args()
out = open(a.out, "a") if a.out else None
for k in [int(t) for t in a.k.split(",")]:
    A = hs.Matrix.stencil(27, a.g, np.float64, 0)
    i = A.info()
    sizes, _, _ = hs.auto_params(i["nnz"], i["m"], hs.ARCH_B200)
    A.set_option(hs.OPT_BANDK_HOST, 1 if a.host_build else 0)
    t0 = time.time()
    A.build_csrk(k, sizes)
    tb = time.time() - t0
    A.fill_x_synthetic(0)
    j = A.info()
    ms = A.spmv_timed(hs.ALGO_CSRK, warmup=3, iters=10)
    t = float(np.median(ms)) * 1e-3
    b = i["nnz"] * 12 + (i["m"] + 1) * 4 + 2 * i["m"] * 8 + (j["n_sr"] + j["n_ssr"] + 2) * 4
    rec = {"g": a.g, "k": k, "sizes": sizes[: k - 1], "n_sr": j["n_sr"], "n_ssr": j["n_ssr"], "bandk_build_s": tb, "bandk_build": "host" if a.host_build else "device (level-1 coarse graph + reorderA in HBM)", "ms": t * 1e3,
           "gbs": b / t / 1e9, "frac": b / t / 1e9 / 6550.7, "gflops": 2 * i["nnz"] / t / 1e9, "kernel": A.last_kernel()}
    print(json.dumps(rec), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")
    A.close()
