"""STS (L x = b on the level-set-ordered matrix): GPU solve against the compiled reference's lowerSTS on the host cores.
usage: python tools/sts_bench.py [--out gpurun_out/sts.jsonl]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ.setdefault("OMP_SCHEDULE", "static")
import heterogeneous_spmv_b200 as hs  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402
import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
a = ap.parse_args()
out = open(a.out, "a") if a.out else None
for name, make in (("cfg1 laplace5 1000^2 rcm", lambda: G.laplace2d_5pt(1000, rcm=True, mode="varied")),
                   ("cfg2 stencil27 100^3", lambda: G.stencil27(100, mode="varied")),
                   ("cfg3 shell blocks 1.5M", lambda: G.shell_blocks(1_500_000))):
    rp, ci, v = make()
    n = len(rp) - 1
    A = hs.Matrix.from_csr(rp, ci, v)
    t0 = time.time()
    A.build_sts()
    build_s = time.time() - t0
    s = A.sts()
    A.sts_compute_b()
    A.set_option(hs.OPT_STS_BARRIERS, 1)
    ms_bar = A.sts_solve_timed(3, 10)
    kern_bar = A.last_kernel()
    A.set_option(hs.OPT_STS_BARRIERS, 2)
    ms_flags = A.sts_solve_timed(3, 10)
    A.set_option(hs.OPT_STS_BARRIERS, 0)
    ms = A.sts_solve_timed(3, 10)
    x = A.sts_get_x()
    nnz_l = (len(ci) + n) // 2
    rec = {"matrix": name, "n": n, "nnz_L": nnz_l, "packs": len(s["pack_ptr"]) - 1, "build_s": build_s, "gpu_ms": float(np.median(ms)),
           "gpu_gflops": 2.0 * nnz_l / float(np.median(ms)) / 1e6, "max_abs_x_minus_1": float(np.max(np.abs(x - 1.0))), "kernel": A.last_kernel(),
           "level_scheduled_ms": float(np.median(ms_bar)), "barrier_free_ms": float(np.median(ms_flags)), "level_scheduled_kernel": kern_bar, "timed_out": A.sts_status()["timed_out"]}
    if O.have_ref():
        t0 = time.time()
        R = O.RefSTS(rp, ci, v)
        rec["ref_build_s"] = time.time() - t0
        rec["ref_ms"] = R.solve_timed(3, 10) * 1e3
        rec["ref_cores"] = len(os.sched_getaffinity(0))
        rec["perm_equal"] = bool(np.array_equal(R.arrays()["perm"], s["perm"]))
    print(json.dumps(rec), flush=True)
    if out:
        out.write(json.dumps(rec) + "\n")
    A.close()
