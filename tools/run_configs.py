"""BASELINE.json configs 1-4 at full size on one GPU: parity of every kernel against the oracle
(whole y, north_star bound) + device-timed throughput.  Writes JSON lines.
usage: python tools/run_configs.py [--configs 1,2,3,4] [--rmat-scale 24] [--out gpurun_out/configs.jsonl]"""
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

PEAK = 6550.7


def alg_bytes(m, n, nnz, vs, extra=0):
    return nnz * (vs + 4) + (m + 1) * 4 + n * vs + m * vs + extra


def check(y, y_ref, scale, tol):
    err = np.abs(y.astype(np.float64) - y_ref.astype(np.float64))
    bad = int(np.count_nonzero(err > tol * scale))
    return bad, float((err / np.maximum(scale, 1e-300)).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4")
    ap.add_argument("--rmat-scale", type=int, default=24)
    ap.add_argument("--rmat-on-device", action="store_true", help="generate config 4 with torch on the GPU (seconds instead of minutes)")
    ap.add_argument("--out", default="")
    ap.add_argument("--dtypes", default="f64,f32")
    a = ap.parse_args()
    out = open(a.out, "a") if a.out else None
    os.environ.setdefault("OMP_SCHEDULE", "static")
    for cfg in [int(t) for t in a.configs.split(",")]:
        t0 = time.time()
        if cfg == 1:
            name, (rp, ci, v) = "cfg1 laplace5 1000^2 RCM", G.laplace2d_5pt(1000, rcm=True)
            ksizes = [(2, None), (3, None), (2, [96]), (3, [8, 8])]
        elif cfg == 2:
            name, (rp, ci, v) = "cfg2 stencil27 100^3", G.stencil27(100)
            ksizes = [(2, None), (3, None), (2, [32]), (3, [20, 10])]
        elif cfg == 3:
            name, (rp, ci, v) = "cfg3 shell 5-dof blocks n=1.5M", G.shell_blocks(1_500_000)
            ksizes = [(2, None), (3, None), (3, [9, 12])]
        elif cfg == 30:
            name, (rp, ci, v) = "cfg3b hostile banded runs n=1.5M", G.banded_shell(1_500_000)
            ksizes = [(2, None), (3, None)]
        else:
            if a.rmat_on_device:
                name, (rp, ci, v) = "cfg4 rmat scale %d ef 16 (torch generator)" % a.rmat_scale, G.rmat_torch(a.rmat_scale, 16)
            else:
                name, (rp, ci, v) = "cfg4 rmat scale %d ef 16" % a.rmat_scale, G.rmat(a.rmat_scale, 16)
            ksizes = []           # Band-k needs a structurally symmetric matrix (reference quirk Q8)
        n = len(rp) - 1
        nnz = len(ci)
        print("== %s: n=%d nnz=%d (%.1f/row, max %d) generated in %.1fs" % (name, n, nnz, nnz / n, int(np.diff(rp).max()), time.time() - t0), flush=True)
        for dt_name in a.dtypes.split(","):
            dt = np.float64 if dt_name == "f64" else np.float32
            tol = 1e-12 if dt_name == "f64" else 1e-5
            vs = 8 if dt_name == "f64" else 4
            vv = v.astype(dt)
            x = G.vector_x(n, dt)
            y_ref = O.csr_spmv(rp, ci, vv, x, parallel=True)
            scale = O.abs_rowsum(rp, ci, vv, x)
            A = hs.Matrix.from_csr(rp, ci, vv)
            A.set_x(x)
            for algo in (hs.ALGO_AUTO, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_MERGE):
                A.set_y(np.full(n, np.nan, dt))
                A.spmv(algo)
                bad, worst = check(A.get_y(), y_ref, scale, tol)
                ms = A.spmv_timed(algo, warmup=3, iters=20, flush_l2=True)
                t = float(np.median(ms)) * 1e-3
                b2b = A.spmv_timed_batch(algo, warmup=5, iters=20)
                rec = {"config": name, "dtype": dt_name, "format": "csr", "algo": hs.ALGO_NAMES[algo], "kernel": A.last_kernel(),
                       "rows_outside_bound": bad, "max_err_over_scale": worst, "ms": t * 1e3, "ms_back_to_back": b2b,
                       "frac_back_to_back": alg_bytes(n, n, nnz, vs) / (b2b * 1e-3) / 1e9 / PEAK,
                       "gbs": alg_bytes(n, n, nnz, vs) / t / 1e9, "gflops": 2 * nnz / t / 1e9}
                rec["frac"] = rec["gbs"] / PEAK
                print("  %-4s %-11s bad=%d worst=%.1e  %8.3f ms %7.1f GB/s (%.3f) %7.1f GF/s  b2b %.3f ms (%.3f) | %s" % (
                    dt_name, rec["algo"], bad, worst, rec["ms"], rec["gbs"], rec["frac"], rec["gflops"], b2b, rec["frac_back_to_back"],
                    rec["kernel"]), flush=True)
                if out:
                    out.write(json.dumps(rec) + "\n")
                    out.flush()
            A.close()
            for k, sizes in ksizes:
                if sizes is None:
                    sizes, _, _ = hs.auto_params(nnz, n, hs.ARCH_B200)
                B = hs.Matrix.from_csr(rp, ci, vv)
                tb = time.time()
                B.build_csrk(k, sizes)
                tb = time.time() - tb
                c = B.csrk()
                o = O.bandk_build(rp, ci, k, sizes)
                exact = all(np.array_equal(c[key], o[key]) for key in o)
                B.set_x(x)
                B.set_y(np.full(n, np.nan, dt))
                B.spmv(hs.ALGO_CSRK)
                bad, worst = check(B.get_y(True), y_ref, scale, tol)
                ms = B.spmv_timed(hs.ALGO_CSRK, warmup=3, iters=20, flush_l2=True)
                t = float(np.median(ms)) * 1e-3
                b2b = B.spmv_timed_batch(hs.ALGO_CSRK, warmup=5, iters=20)
                i = B.info()
                rec = {"config": name, "dtype": dt_name, "format": "csr%d" % k, "sizes": sizes[: k - 1], "n_sr": i["n_sr"], "n_ssr": i["n_ssr"],
                       "index_work_bit_exact": bool(exact), "bandk_build_s": tb, "kernel": B.last_kernel(),
                       "rows_outside_bound": bad, "max_err_over_scale": worst, "ms": t * 1e3, "ms_back_to_back": b2b,
                       "gbs": alg_bytes(n, n, nnz, vs, (i["n_sr"] + i["n_ssr"] + 2) * 4) / t / 1e9, "gflops": 2 * nnz / t / 1e9}
                rec["frac"] = rec["gbs"] / PEAK
                rec["frac_back_to_back"] = rec["frac"] * rec["ms"] / b2b
                print("  %-4s csr-%d %-9s exact=%s bad=%d worst=%.1e  %8.3f ms %7.1f GB/s (%.3f) %7.1f GF/s  b2b %.3f ms (%.3f) build %.2fs | %s" % (
                    dt_name, k, sizes[: k - 1], exact, bad, worst, rec["ms"], rec["gbs"], rec["frac"], rec["gflops"], b2b,
                    rec["frac_back_to_back"], tb, rec["kernel"]), flush=True)
                if out:
                    out.write(json.dumps(rec) + "\n")
                    out.flush()
                B.close()


if __name__ == "__main__":
    main()
