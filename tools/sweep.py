"""Kernel parameter sweep on one GPU (development tool; prints a table, writes JSON lines).
usage: python tools/sweep.py [--g 200] [--kind 27] [--dtype f64] [--out gpurun_out/sweep.jsonl] [--quick]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heterogeneous_spmv_b200 as hs  # noqa: E402


def alg_bytes(m, n, nnz, vs, n_sr=0, n_ssr=0):
    return nnz * (vs + 4) + (m + 1) * 4 + n * vs + m * vs + (n_sr + n_ssr + 2) * 4 * (1 if n_sr else 0)


def run(A, algo, opts, iters, flush, peak, label, out):
    for k in (hs.OPT_VEC, hs.OPT_WARPS, hs.OPT_STAGES, hs.OPT_STAGE_NNZ, hs.OPT_CTAS_PER_SM, hs.OPT_ROWS_PER_TILE,
              hs.OPT_L2_PERSIST_X, hs.OPT_MERGE_TILE, hs.OPT_CHUNK):
        A.set_option(k, 0)
    for k, v in opts.items():
        A.set_option(k, v)
    try:
        ms = A.spmv_timed(algo, warmup=3, iters=iters, flush_l2=flush)
    except hs.HspmvError as e:
        print("%-60s FAILED %s" % (label, e))
        return None
    i = A.info()
    vs = 8 if A.dtype == np.float64 else 4
    b = alg_bytes(i["m"], i["n"], i["nnz"], vs, i["n_sr"] if algo == hs.ALGO_CSRK else 0, i["n_ssr"] if algo == hs.ALGO_CSRK else 0)
    t = float(np.median(ms)) * 1e-3
    tmin = float(ms.min()) * 1e-3
    rec = {"label": label, "kernel": A.last_kernel(), "ms_med": t * 1e3, "ms_min": tmin * 1e3, "gbs": b / t / 1e9,
           "gflops": 2 * i["nnz"] / t / 1e9, "frac": b / t / 1e9 / peak, "flush": bool(flush)}
    print("%-44s %8.3f ms  %7.1f GB/s  %6.1f GF/s  frac %.3f  | %s" % (label, rec["ms_med"], rec["gbs"], rec["gflops"], rec["frac"], rec["kernel"]))
    if out:
        out.write(json.dumps(rec) + "\n")
        out.flush()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--g", type=int, default=200)
    ap.add_argument("--kind", type=int, default=27)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--flush", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--csrk", type=int, default=1)
    ap.add_argument("--peak", type=float, default=6550.7)
    a = ap.parse_args()
    dt = np.float64 if a.dtype == "f64" else np.float32
    out = open(a.out, "a") if a.out else None
    print(hs.device_info())
    t0 = time.time()
    A = hs.Matrix.stencil(a.kind, a.g, dt, 0)
    A.fill_x_synthetic(0)
    i = A.info()
    print("matrix kind=%d g=%d m=%d nnz=%d (%.2f GB alg) gen %.2fs" % (a.kind, a.g, i["m"], i["nnz"], alg_bytes(i["m"], i["n"], i["nnz"], dt().itemsize) / 1e9, time.time() - t0))
    tag = "k%d g%d %s " % (a.kind, a.g, a.dtype)
    F = a.flush
    vecs = (2, 4, 8, 16) if a.kind == 27 else (1, 2, 4)
    for vec in vecs:
        run(A, hs.ALGO_CSR_VECTOR, {hs.OPT_VEC: vec}, a.iters, F, a.peak, tag + "vector vec=%d" % vec, out)
    run(A, hs.ALGO_CSR_MERGE, {}, a.iters, F, a.peak, tag + "merge tile=4096", out)
    if not a.quick:
        for tile, c in ((2048, 4), (4096, 6), (8192, 3)):
            run(A, hs.ALGO_CSR_MERGE, {hs.OPT_MERGE_TILE: tile, hs.OPT_CTAS_PER_SM: c}, a.iters, F, a.peak, tag + "merge tile=%d ctas=%d" % (tile, c), out)
    svecs = (1, 2, 4) if a.kind == 27 else (1, 2)
    for vec in svecs:
        run(A, hs.ALGO_CSR_STREAM, {hs.OPT_VEC: vec}, a.iters, F, a.peak, tag + "stream vec=%d default" % vec, out)
    if not a.quick:
        best = None
        if a.kind == 27:
            grid = [(1, 9, 32), (1, 13, 32), (1, 8, 32), (1, -1, 32), (1, 9, 64), (2, 13, 16), (2, 13, 32), (2, 6, 16), (2, 8, 32), (2, -1, 16),
                    (4, 6, 16), (4, 6, 24), (4, 6, 32), (4, 3, 24), (4, -1, 24), (4, -1, 16), (8, 3, 16), (8, 3, 24), (8, -1, 16)]
        else:
            grid = [(1, 4, 96), (1, 4, 128), (1, 4, 64), (1, -1, 96), (2, -1, 64), (1, 4, 160)]
        for vec, ch, rpt in grid:
            for warps, stages in ((16, 2), (12, 3), (8, 4)):
                r = run(A, hs.ALGO_CSR_STREAM, {hs.OPT_VEC: vec, hs.OPT_WARPS: warps, hs.OPT_STAGES: stages, hs.OPT_ROWS_PER_TILE: rpt, hs.OPT_CHUNK: ch},
                        a.iters, F, a.peak, tag + "stream vec=%d ch=%d w=%d s=%d rpt=%d" % (vec, ch, warps, stages, rpt), out)
                if r and (best is None or r["gbs"] > best["gbs"]):
                    best = r
        print("BEST stream:", best["label"], "%.1f GB/s" % best["gbs"], best["kernel"])
    if a.csrk and i["nnz"] < 300e6:
        for k, sizes in ((2, [16]), (2, [32]), (3, [16, 16]), (3, [32, 8])):
            B = hs.Matrix.stencil(a.kind, a.g, dt, 0)
            t0 = time.time()
            B.build_csrk(k, sizes)
            tb = time.time() - t0
            B.fill_x_synthetic(0)
            j = B.info()
            print("  band-k k=%d sizes=%s: n_sr=%d n_ssr=%d build %.2fs" % (k, sizes, j["n_sr"], j["n_ssr"], tb))
            for vec in svecs[:3]:
                run(B, hs.ALGO_CSRK, {hs.OPT_VEC: vec}, a.iters, F, a.peak, tag + "csrk k=%d %s vec=%d" % (k, sizes, vec), out)
            if not a.quick:
                for vec, ch in ((1, 9), (2, 13), (4, 6), (4, -1)):
                    for warps, stages in ((16, 2), (12, 3)):
                        run(B, hs.ALGO_CSRK, {hs.OPT_VEC: vec, hs.OPT_WARPS: warps, hs.OPT_STAGES: stages, hs.OPT_CHUNK: ch}, a.iters, F, a.peak,
                            tag + "csrk k=%d %s vec=%d w=%d s=%d ch=%d" % (k, sizes, vec, warps, stages, ch), out)
            B.close()
    A.close()


if __name__ == "__main__":
    main()
