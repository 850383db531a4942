#!/usr/bin/env python3
"""run_scripts-compatible runner for the B200 drivers.

Does what the reference's per-machine scripts do -- run_scripts/run_cuda.py (spmv.exe / spmv-auto over a
directory of matrices), run_cuda_new.py:85-94 (spmv-auto-volta) and run_tuning.py:26,89 (spmv over a grid of
super-row sizes) -- with relative paths and a command line instead of edited constants:

    python run_b200.py --matrices DIR [--kernel cuda-spmv-csr|cuda-spmv-csrk|tuning] [--runs 20]
                       [--record FILE.csv] [--bin DIR] [--sizes 16,24,32,...] [--extra "--fp32 --cache"]

Same per-run stdout files (runs/<kernel>/<matrix>_cuda_1_<threads>.txt), same scrape of the drivers' output
(`find("TimeMin:") + 8 ... "\\n"`, run_cuda.py:110-122) and the same record lines:
    kernel, matrix, sch, chunk, threads, min, max, avg,                 (run_cuda.py:125)
    kernel, matrix, sch, chunk, threads, (srs), min, max, avg,          (run_tuning.py:125)
CSR files are read 0-based by cuda-spmv-csr and 1-based by cuda-spmv-csrk, like the reference's (Q1).
"""
import argparse
import os
import pathlib
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRSS = [16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768]        # run_tuning.py:26


def scrape(out_utf, key):
    """the reference's slicing, verbatim in behaviour: text between `key` and the next newline"""
    loc = out_utf.find(key)
    if loc < 0:
        return ""
    return out_utf[loc + 8: out_utf.find("\n", loc)]


def run_one(cmd, io_file, timeout):
    try:
        trun = subprocess.run(cmd, stdout=subprocess.PIPE, timeout=timeout)
    except subprocess.TimeoutExpired:
        print("Timeout: ", " ".join(cmd))
        return None
    out_utf = trun.stdout.decode("utf-8", errors="replace")
    with open(io_file, "w") as fid:
        fid.write(out_utf)
    return out_utf


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--matrices", required=True, help="directory of .mtx.rcm.csr files (mat_loc in the reference scripts)")
    ap.add_argument("--kernel", default="cuda-spmv-csrk", choices=["cuda-spmv-csr", "cuda-spmv-csrk", "tuning"])
    ap.add_argument("--runs", type=int, default=20)
    ap.add_argument("--record", default="b200_spmv.csv")
    ap.add_argument("--bin", default=os.path.join(HERE, "..", "bin"), help="where the drivers were built (make -C csrc)")
    ap.add_argument("--out", default="runs", help="directory of the per-run stdout files")
    ap.add_argument("--sizes", default=",".join(map(str, SRSS)), help="tuning: super-row sizes to sweep")
    ap.add_argument("--extra", default="", help="extra driver flags, e.g. '--fp32 --cache'")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1, help="recorded like the reference does; host side only")
    ap.add_argument("--timeout", type=int, default=600)
    a = ap.parse_args(argv)

    label = "cuda-spmv-csrk/cuda" if a.kernel == "tuning" else a.kernel
    exe = {"cuda-spmv-csr": os.path.join(a.bin, "cuda-spmv-csr", "spmv.exe"),
           "cuda-spmv-csrk": os.path.join(a.bin, "cuda-spmv-csrk", "cuda", "spmv-auto-b200"),
           "tuning": os.path.join(a.bin, "cuda-spmv-csrk", "cuda", "spmv")}[a.kernel]
    if not os.path.exists(exe):
        sys.exit("driver not found: %s (build with `make -C heterogeneous-spmv_b200/csrc`)" % exe)
    temp_dir = pathlib.Path(a.out) / label
    temp_dir.mkdir(parents=True, exist_ok=True)
    os.environ.setdefault("OMP_PLACES", "cores")
    os.environ.setdefault("OMP_PROC_BIND", "true")
    os.environ["OMP_NUM_THREADS"] = str(a.threads)
    sch, chunk = "cuda", 1
    extra = a.extra.split()
    print("-------" + label + "------------")
    for mat in sorted(os.listdir(a.matrices)):
        mat_long = os.path.join(a.matrices, mat)
        if not os.path.isfile(mat_long) or mat.endswith(".hspmvbin"):
            continue
        sweeps = [int(t) for t in a.sizes.split(",")] if a.kernel == "tuning" else [None]
        for srs in sweeps:
            print(label + "_" + mat + "_" + sch + ("" if srs is None else "_%d" % srs))
            io_file = str(temp_dir / (mat + "_" + sch + "_" + str(chunk) + "_" + str(a.threads) + ("" if srs is None else "_%d" % srs) + ".txt"))
            cmd = [exe, mat_long, str(a.runs)] + ([] if srs is None else [str(srs)]) + extra
            out_utf = run_one(cmd, io_file, a.timeout)
            if out_utf is None:
                continue
            min_val, max_val, avg_val = (scrape(out_utf, k) for k in ("TimeMin:", "TimeMax:", "TimeAvg:"))
            mid = "" if srs is None else "(" + str(srs) + "), "
            line = label + ", " + mat + ", " + sch + ", " + str(chunk) + ", " + str(a.threads) + ", " + mid + min_val + ", " + max_val + ", " + avg_val + ", \n"
            with open(a.record, "a+") as fid_rec:
                fid_rec.write(line)
            print(line)


if __name__ == "__main__":
    main()
