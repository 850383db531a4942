"""The reference's own OpenMP kernels on this host (SURVEY.md 8d "CPU baseline beside it"): omp_spmv
(spmv-csr/spmv.c) and CSRk_Graph::SpMV k = 2 / k = 3 (spmv-csrk/csrk.cpp) compiled unmodified into oracle/_ref,
fp64 build, all host threads, OMP_SCHEDULE=static, 5 warm-ups + 20 timed; the CSR-k super-row sizes are swept over a
subset of run_scripts/run_epyc.py:24's grid and the best is reported.  No GPU is used.
usage: python tools/cpu_reference_table.py [--configs 1,2,3] [--out f.jsonl]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402

cores = bench.cpu_setup_threads(force=True)
import numpy as np  # noqa: E402
import oracle as O  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="1,2,3")
ap.add_argument("--out", default="")
ap.add_argument("--sizes", default="8,16,32,64,128")
a = ap.parse_args()
out = open(a.out, "a") if a.out else None
if not O.have_ref():
    raise SystemExit("oracle/_ref is missing (built from /root/reference by oracle/Makefile)")
bench._force_omp_threads()
for cfg in [int(t) for t in a.configs.split(",")]:
    if cfg == 1:
        name, (rp, ci, v) = "cfg1 laplace5 1000^2 RCM", G.laplace2d_5pt(1000, rcm=True)
    elif cfg == 2:
        name, (rp, ci, v) = "cfg2 stencil27 100^3", G.stencil27(100)
    else:
        name, (rp, ci, v) = "cfg3 shell 5-dof blocks n=1.5M", G.shell_blocks(1_500_000)
    n, nnz = len(rp) - 1, len(ci)
    x = G.vector_x(n, np.float64)
    t, _ = O.ref_csr_spmv_timed(rp, ci, v, x, 5, 20)
    rec = {"config": name, "cores": cores, "format": "csr", "kernel": "omp_spmv (spmv-csr/spmv.c:92-114)", "ms": t * 1e3, "gflops": 2 * nnz / t / 1e9}
    print(json.dumps(rec), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")
    for k in (2, 3):
        best = None
        for s in [int(t) for t in a.sizes.split(",")]:
            sizes = [s] if k == 2 else [s, 8]
            t0 = time.time()
            R = O.RefCSRk(rp, ci, v, k, sizes)
            tb = time.time() - t0
            t, _ = R.spmv_timed(x, 5, 20)
            R.close()
            if best is None or t < best[0]:
                best = (t, sizes, tb)
        rec = {"config": name, "cores": cores, "format": "csr%d" % k, "kernel": "CSRk_Graph::SpMV (spmv-csrk/csrk.cpp:165-285)",
               "best_sizes": best[1], "sizes_tried": a.sizes, "ms": best[0] * 1e3, "gflops": 2 * nnz / best[0] / 1e9, "bandk_build_s": best[2]}
        print(json.dumps(rec), flush=True)
        if out:
            out.write(json.dumps(rec) + "\n")
