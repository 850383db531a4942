#!/usr/bin/env python
"""bench.py -- the reference's headline measurement on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload stencil400|rmat24]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): fp64 SpMV GFLOP/s (2*nnz/t) with the HBM roofline beside it.

Workload `stencil400` (default) = BASELINE.json configs[4]: the 3-D 27-point stencil on a 400^3 grid
(n = 64,000,000, nnz = 1,719,374,392, 21.9 GB of algorithmic traffic per SpMV), the configuration the
"1/2/4/8 B200" part of the metric is quoted on; it fits one GPU, so N = 1 runs the same matrix and N > 1 is
strong scaling of it (nnz-balanced row blocks + x halo exchange over NVLink).  A step is one y = A*x over the
whole matrix.  The matrix is far larger than L2 (126 MB), so no flush is needed between timed iterations.

At N > 1 the result is CHECKED, not just timed: the halo regions of both x buffers start as NaN, the boundary
slices of x that are exchanged change every step (so every step needs a fresh halo), and after the timed loop
every rank compares all of its rows of y with a second kernel run on a locally built copy of the expected x
and with a numpy product on windows at both block ends and in the middle; the line carries the `parity`
record and the process exits non-zero if a row is outside the north_star bound.

Workload `rmat24` = BASELINE.json configs[3]: R-MAT scale 24 (n = 16.8 M, nnz ~ 256 M), CSR with row-length
binning; N > 1 = replicated x, row-sharded y, no per-step collective (all_gather mode timed beside it).  The
default line also carries it as `rmat24` at every N, and configs[0..2] as `configs` at N = 1.

One JSON line on stdout (rank 0).  `value` = device-timed (CUDA events, max over ranks) with inputs resident
in HBM; `e2e` = the same metric through the C-ABI call with HOST buffers (hspmv_spmv_host: H2D of x, kernel,
D2H of y inside the timed region).  `--impl reference` times the reference's own OpenMP spmv-csr (oracle/_ref,
compiled from the unmodified sources; the oracle port if that is missing) on the host cores -- on the WHOLE
matrix when the host has the memory for it (22 GB), else on the largest leading row block that fits.
"""
import argparse
import bisect
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_DEFAULT = 400
CPU_SAMPLE_ROWS = 4_000_000   # bounded sample for the cpu_baseline leg of OUR line: first 4 M rows (25 z-planes of 400^2)
NOMINAL_GBS = 8000.0


# ----------------------------------------------------------------------------- closed forms
def line_len(t, g):
    return 1 if g == 1 else 3 - (t == 0) - (t == g - 1)


def line_prefix(t, g):
    return t if g == 1 else (0 if t == 0 else 3 * t - 1)


def stencil27_before(i, g):
    """nonzeros in rows < i (same closed form as csrc/gen.cuh and generators.py)."""
    S = 1 if g == 1 else 3 * g - 2
    if i >= g ** 3:
        return S ** 3
    x, y, z = i % g, (i // g) % g, i // (g * g)
    return line_prefix(z, g) * S * S + line_len(z, g) * (line_prefix(y, g) * S + line_len(y, g) * line_prefix(x, g))


class _RowPtr:
    def __init__(self, g):
        self.g = g

    def __len__(self):
        return self.g ** 3 + 1

    def __getitem__(self, i):
        return stencil27_before(i, self.g)


def partition_stencil(g, parts):
    """hspmv_partition_rows' rule on the closed-form row pointer: row_lo[p] = first row r with
    row_ptr[r] >= floor(p*nnz/P)."""
    rp = _RowPtr(g)
    nnz = stencil27_before(g ** 3, g)
    out = [0]
    for p in range(1, parts):
        out.append(max(out[-1], bisect.bisect_left(rp, (nnz * p) // parts)))
    out.append(g ** 3)
    return out


def alg_bytes(m, n, nnz, vs=8, extra=0):
    """nnz*(V+4) + (m+1)*4 + n*V (x once) + m*V (y) [+ CSR-k maps]."""
    return nnz * (vs + 4) + (m + 1) * 4 + n * vs + m * vs + extra


def step_scale(t):
    """factor applied to the exchanged slices of x at step t (exact in fp64 and fp32)"""
    return 1.0 + (t % 8) / 8.0


def merge_segments(segs):
    out = []
    for a, b in sorted(segs):
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return [(a, b) for a, b in out]


# ----------------------------------------------------------------------------- clocks (NVML, in process)
class ClockSampler:
    """SM clock + throttle reasons sampled every 20 ms by a thread of this process through NVML
    (nvidia_ml_py): no nvidia-smi fork next to the timed region."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake"}

    def __init__(self, index, uuid=None):
        self.samples = []
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            h = None
            if uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)) if not str(uuid).startswith("GPU-") else str(uuid))
                except Exception:
                    h = None
            self.h = h if h is not None else nv.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.ok = False

    def _reasons(self):
        nv = self.nv
        for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            fn = getattr(nv, name, None)
            if fn is not None:
                try:
                    return int(fn(self.h))
                except Exception:
                    continue
        return 0

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append((time.time(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), self._reasons()))
            except Exception:
                pass
            self._stop.wait(0.02)

    def stop(self, t_lo, t_hi):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        self._stop.set()
        self.t.join(timeout=1.0)
        sm, reasons = [], set()
        for ts, clk, bits in self.samples:
            if not (t_lo - 0.05 <= ts <= t_hi + 0.05):
                continue
            sm.append(clk)
            for bit, name in self.REASONS.items():
                if bits & bit:
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml in-process, 20 ms"}


# ----------------------------------------------------------------------------- CPU arms
def cpu_setup_threads(force=False):
    """OpenMP environment of the CPU arms (run_scripts/run_epyc.py:78-80 style): all host cores,
    static schedule.  Must run before libgomp is loaded; torchrun exports OMP_NUM_THREADS=1, which
    the CPU arm overrides (force)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if force or "OMP_NUM_THREADS" not in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_SCHEDULE", "static")      # spmv-csr uses schedule(runtime); libgomp's default is dynamic,1
    os.environ.setdefault("OMP_PROC_BIND", "true")
    os.environ.setdefault("OMP_PLACES", "cores")
    return int(os.environ["OMP_NUM_THREADS"])


def _force_omp_threads():
    """torchrun exports OMP_NUM_THREADS=1 and libgomp may have read it before cpu_setup_threads()
    ran (it reads the environment once, when it is loaded): set the team size of the libgomp
    instance the CPU arm links to, and report what it will really use."""
    import ctypes as C
    want = int(os.environ.get("OMP_NUM_THREADS", "1"))
    try:
        gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
        gomp.omp_set_num_threads(want)
        return int(gomp.omp_get_max_threads())
    except OSError:
        return want


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    return O


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


def cpu_csr_timed(rp, ci, v, x, warmup, iters):
    """(mean seconds per omp_spmv, kind): the compiled reference if oracle/_ref is there, else the oracle port."""
    import ctypes as C
    import numpy as np
    O = _oracle()
    if O.have_ref():
        L = O.ref("csr", np.float64)
        _force_omp_threads()
        y = np.zeros(len(rp) - 1)
        fp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
        L.ref_csr_spmv_timed.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, fp, fp, fp, C.c_int, C.c_int, C.POINTER(C.c_double)]
        L.ref_csr_spmv_timed.restype = C.c_double
        avg = C.c_double()
        # the reference's omp_spmv only ever indexes x by column id, so a rectangular sample (rows [0,R) with
        # columns up to R + g^2+g+1) is a valid input for it
        L.ref_csr_spmv_timed(len(rp) - 1, len(x), len(ci), np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32),
                             v, x, y, warmup, iters, C.byref(avg))
        return float(avg.value), "reference"      # mean over the timed calls, like the GPU arm's K-step mean
    _force_omp_threads()
    t, _ = O.csr_spmv_timed(rp, ci, v, x, warmup, iters)   # (min; the port has no mean output)
    return t, "port"


def cpu_stencil_sample(g, rows, warmup, iters):
    """(seconds per SpMV, nnz of the sample, kind) on rows [0,rows) of the stencil; arrays built by the oracle's C
    closed-form generator, so this leg touches nothing of the GPU library."""
    from heterogeneous_spmv_b200 import generators as G
    O = _oracle()
    rows = min(rows, g ** 3)
    n_cols = min(g ** 3, rows + g * g + g + 1)
    rp, ci, v = O.stencil27(g, 0, rows)
    t, kind = cpu_csr_timed(rp, ci, v, G.vector_x(n_cols), warmup, iters)
    return t, int(len(ci)), kind


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation (spmv-csr omp_spmv) on the host cores, on our arm's
    config.  Rank 0 only; maps oracle/ libraries only."""
    if rank != 0:
        return
    cores = cpu_setup_threads(force=True)
    if args.workload == "rmat24":
        return run_reference_rmat(args, cores)
    g = args.g
    n, nnz_all = g ** 3, stencil27_before(g ** 3, g)
    need = nnz_all * 12 + n * 20 + (64 << 20)
    avail = mem_available_bytes()
    rows, why = n, "whole matrix"
    if avail and need * 1.10 > avail:
        planes = max(1, int(g * (avail * 0.8) / need))
        rows = planes * g * g
        why = "host has %.1f GB available, the matrix needs %.1f GB: leading %d of %d z-planes" % (avail / 1e9, need / 1e9, planes, g)
    t, nnz, kind = cpu_stencil_sample(g, rows, args.warmup, args.steps)
    gf = 2.0 * nnz / t / 1e9
    sample = "rows [0,%d) of stencil27 g=%d (%s): %d nnz; mean of %d timed omp_spmv calls after %d warm-ups, OMP_SCHEDULE=static" % (
        rows, g, why, nnz, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": "fp64 SpMV GFLOP/s", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "stencil27_g%d_csr" % g, "n": n, "nnz": nnz_all, "format": "csr", "rows_timed": rows},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_reference_rmat(args, cores):
    import numpy as np
    from heterogeneous_spmv_b200 import generators as G
    rp, ci, v, how = rmat_arrays(G, args.rmat_scale)
    n = len(rp) - 1
    t, kind = cpu_csr_timed(rp, ci, v, G.vector_x(n), args.warmup, args.steps)
    gf = 2.0 * len(ci) / t / 1e9
    print(json.dumps({
        "impl": "reference", "metric": "fp64 SpMV GFLOP/s", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "rmat%d_csr" % args.rmat_scale, "n": n, "nnz": int(len(ci)), "format": "csr", "generator": how},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": kind,
                         "sample": "whole matrix, mean of %d omp_spmv calls after %d warm-ups" % (args.steps, args.warmup)},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)
    del np


def rmat_arrays(G, scale):
    """R-MAT (a,b,c,d) = (.57,.19,.19,.05), edge factor 16, host CSR arrays: torch generator on the GPU when there is
    one (seconds), numpy otherwise (minutes at scale 24)."""
    try:
        import torch
        if torch.cuda.is_available():
            return G.rmat_torch(scale, 16) + ("generators.rmat_torch (cuda)",)
    except Exception:
        pass
    return G.rmat(scale, 16) + ("generators.rmat (numpy)",)


# ----------------------------------------------------------------------------- helpers of our arm
def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    return peak_gbs, src


def bind_near_gpu(local_rank):
    """Pin this process to the CPUs NVML names as local to the GPU before the pinned host buffers are allocated
    (first touch then puts them on that NUMA node).  Returns a short description."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to gpu %d" % (len(cpus), local_rank)
        return "no narrower affinity reported (%d cpus)" % len(allowed)
    except Exception as e:      # noqa: BLE001
        return "unavailable (%s)" % type(e).__name__


def build_phases(B):
    """wall seconds of the Band-k build by phase (hspmv_build_phases); never the reason for a missing bench line"""
    try:
        return B.build_phases()
    except Exception:      # noqa: BLE001
        return None


def timed_formats(A, hs, algo, n, nnz, vs, extra=0, peak=6550.7, iters=20):
    """flushed (L2 cleared before every launch, median) and back-to-back times of one handle + roofline fractions"""
    import numpy as np
    ms = A.spmv_timed(algo, warmup=3, iters=iters, flush_l2=True)
    t = float(np.median(ms))
    # the same with L2 left CLEAN (a read pass over a second buffer after the written one): a launch timed right after
    # the memset also pays for the write-back of the memset's dirty lines, which is not part of the product
    t_clean = float(np.median(A.spmv_timed(algo, warmup=1, iters=iters, flush_l2=2)))
    b2b = A.spmv_timed_batch(algo, warmup=5, iters=iters)
    b = alg_bytes(n, n, nnz, vs, extra)
    try:
        graph = A.spmv_graph_timed(algo, iters=iters, reps=5)     # the same products as one CUDA-graph replay per `iters`
    except Exception:      # noqa: BLE001
        graph = None
    return {"ms_flushed": t, "ms_flushed_clean_l2": t_clean, "ms_back_to_back": b2b, "ms_graph_replay": graph,
            "gflops_flushed": 2.0 * nnz / (t * 1e-3) / 1e9,
            "gflops_back_to_back": 2.0 * nnz / (b2b * 1e-3) / 1e9,
            "roofline": {"algorithmic_bytes": b, "frac_measured_flushed": b / (t * 1e-3) / 1e9 / peak,
                         "frac_measured_flushed_clean_l2": b / (t_clean * 1e-3) / 1e9 / peak,
                         "frac_measured_back_to_back": b / (b2b * 1e-3) / 1e9 / peak,
                         "frac_nominal_flushed": b / (t * 1e-3) / 1e9 / NOMINAL_GBS,
                         "frac_nominal_back_to_back": b / (b2b * 1e-3) / 1e9 / NOMINAL_GBS},
            "kernel": A.last_kernel()}


def numpy_spmv(rp, ci, v, x):
    """closed-form check product (numpy, independent of the library and of oracle/): y and sum_j |a_ij x_j|"""
    import numpy as np
    prod = v * x[ci]
    nz = np.diff(rp) > 0
    y = np.zeros(len(rp) - 1, dtype=np.float64)
    s = np.zeros(len(rp) - 1, dtype=np.float64)
    if len(ci):
        starts = rp[:-1][nz].astype(np.int64)
        y[nz] = np.add.reduceat(prod.astype(np.float64), starts)
        s[nz] = np.add.reduceat(np.abs(prod).astype(np.float64), starts)
    return y, s


def parity_vs_numpy(A, hs, rp, ci, v, x, algo, tol=1e-12, perm_order=True):
    """whole y of handle A (x uploaded here) against the numpy product; rows outside tol * sum|a x|"""
    import numpy as np
    A.set_x(x)
    A.set_y(np.full(len(rp) - 1, np.nan, v.dtype))
    A.spmv(algo)
    y = A.get_y(True)
    y_ref, scale = numpy_spmv(rp, ci, v, x)
    err = np.abs(y.astype(np.float64) - y_ref)
    return {"rows_checked": int(len(y)), "rows_outside_bound": int(np.count_nonzero(~(err <= tol * scale))),
            "max_err_over_scale": float(np.nanmax(err / np.maximum(scale, 1e-300))) if len(y) else 0.0}


def copy_same_bytes(torch, nbytes, iters=20):
    """What a plain device copy moving the SAME number of bytes (half read, half written) achieves at this size, with L2
    flushed by a write + read pass before every copy and back to back: the size-matched version of the peak in
    MEASURED_PEAKS.json (which is a 2 x 2 GiB copy; an 80 MB transfer never reaches it -- ramp-up and launch are
    a fixed ~10 us)."""
    n = max(1, nbytes // 16)
    a = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
    b = torch.empty_like(a)
    fl_w = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fl_r = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(iters):
        fl_w.fill_(i & 0xff)
        fl_r.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        b.copy_(a)
    e0.record()
    for _ in range(iters):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    return {"bytes": int(n * 16), "ms_flushed_clean_l2": ts[len(ts) // 2], "ms_back_to_back": e0.elapsed_time(e1) / iters}


def config_section(hs, np, peak, no_cpu):
    """BASELINE configs[0..2] (cfg1-cfg3) on one GPU: CSR (AUTO kernel) and CSR-k k = 2 / 3 (B200 heuristic sizes,
    Band-k built in HBM), each with whole-y parity against a numpy product, flushed and back-to-back times, roofline
    fractions of the measured and of the nominal 8 TB/s peak, and the compiled reference's omp_spmv / CSRk_Graph::SpMV
    on the host cores beside it."""
    from heterogeneous_spmv_b200 import generators as G
    out = {}
    O = None if no_cpu else _oracle()
    specs = [("cfg1", "laplace5_g1000_rcm", lambda: G.laplace2d_5pt(1000, rcm=True)),
             ("cfg2", "stencil27_g100", lambda: G.stencil27(100, mode="varied")),
             ("cfg3", "shell_blocks_n1500000", lambda: G.shell_blocks(1_500_000))]
    for key, name, make in specs:
        t0 = time.time()
        try:
            rp, ci, v = make()
            n, nnz = len(rp) - 1, len(ci)
            x = G.vector_x(n)
            rec = {"workload": name, "n": n, "nnz": nnz, "formats": {}}
            try:
                import torch
                rec["copy_same_bytes"] = copy_same_bytes(torch, alg_bytes(n, n, nnz))
            except Exception as e:      # noqa: BLE001
                rec["copy_same_bytes"] = {"error": "%s: %s" % (type(e).__name__, e)}
            A = hs.Matrix.from_csr(rp, ci, v)
            f = timed_formats(A, hs, hs.ALGO_AUTO, n, nnz, 8, 0, peak)
            f["parity"] = parity_vs_numpy(A, hs, rp, ci, v, x, hs.ALGO_AUTO)
            rec["formats"]["csr"] = f
            A.close()
            sizes, _, _ = hs.auto_params(nnz, n, hs.ARCH_B200)
            for k in (2, 3):
                B = hs.Matrix.from_csr(rp, ci, v)
                tb = time.time()
                B.build_csrk(k, sizes)
                tb = time.time() - tb
                i = B.info()
                f = timed_formats(B, hs, hs.ALGO_CSRK, n, nnz, 8, (i["n_sr"] + i["n_ssr"] + 2) * 4, peak)
                f.update({"sizes": sizes[: k - 1], "n_sr": i["n_sr"], "n_ssr": i["n_ssr"], "bandk_build_s": tb,
                          "bandk_phases_s": build_phases(B)})
                f["parity"] = parity_vs_numpy(B, hs, rp, ci, v, x, hs.ALGO_CSRK)
                rec["formats"]["csr%d" % k] = f
                B.close()
            if O is not None and O.have_ref():
                cores = cpu_setup_threads()
                t, kind = cpu_csr_timed(rp, ci, v, x, 5, 20)
                cpu = {"cores": cores, "kind": kind, "csr": {"gflops": 2.0 * nnz / t / 1e9, "ms": t * 1e3}}
                O.ref("csrk", np.float64)
                _force_omp_threads()
                for k in (2, 3):
                    best = None
                    for s in ((16, 64) if k == 2 else (16,)):
                        R = O.RefCSRk(rp, ci, v, k, [s, 8])
                        tk, _ = R.spmv_timed(x, 3, 10)
                        R.close()
                        if best is None or tk < best[0]:
                            best = (tk, s)
                    cpu["csr%d" % k] = {"gflops": 2.0 * nnz / best[0] / 1e9, "ms": best[0] * 1e3,
                                        "sizes": [best[1]] if k == 2 else [best[1], 8]}
                rec["cpu_reference"] = cpu
            if key in ("cfg1", "cfg2", "cfg3"):
                # triangular solve L x = b on the level-set-ordered matrix (k = 2, LS ordering), b = L * 1,
                # beside the compiled reference's lowerSTS on the host cores
                S = hs.Matrix.from_csr(rp, ci, v)
                tb = time.time()
                S.build_sts()
                tb = time.time() - tb
                S.sts_compute_b()
                ms = S.sts_solve_timed(3, 10)
                xs = S.sts_get_x()
                nnz_l = (nnz + n) // 2
                sts = {"packs": int(len(S.sts()["pack_ptr"]) - 1), "nnz_L": nnz_l, "build_s": tb, "ms": float(np.median(ms)),
                       "gflops": 2.0 * nnz_l / float(np.median(ms)) / 1e6, "max_abs_x_minus_1": float(np.max(np.abs(xs - 1.0))),
                       "algorithmic_gbs": (nnz_l * 12 + n * 28) / float(np.median(ms)) / 1e6, "kernel": S.last_kernel()}
                S.close()
                if O is not None and O.have_ref():
                    R = O.RefSTS(rp, ci, v)
                    sts["cpu_reference_ms"] = R.solve_timed(3, 10) * 1e3
                    sts["cpu_reference_x_is_one"] = bool(np.all(R.solve() == 1.0))
                rec["sts"] = sts
            rec["seconds"] = time.time() - t0
            out[key] = rec
        except Exception as e:      # noqa: BLE001  (an extra, never the reason for a missing bench line)
            out[key] = {"workload": name, "error": "%s: %s" % (type(e).__name__, e)}
    return out


# ----------------------------------------------------------------------------- R-MAT (BASELINE configs[3])
def rmat_section(args, hs, hd, np, torch, dist, rank, world, peak, steps, warmup, full_line=False):
    """R-MAT scale S on `world` GPUs: CSR, AUTO kernel; N > 1 = replicated x, nnz-balanced row blocks of y, a step is
    the local SpMV on every rank (no collective while x is static); the all_gather that rebuilds x from y (power
    iteration mode) is timed beside it.  Every rank builds the same matrix with the same seeded torch generator on its
    own GPU and keeps its block.  Parity: the block's y against a numpy product (all rows of the block)."""
    from heterogeneous_spmv_b200 import generators as G
    t0 = time.time()
    rp, ci, v = G.rmat_torch(args.rmat_scale, 16)
    torch.cuda.empty_cache()
    n, nnz = len(rp) - 1, len(ci)
    row_lo = [int(t) for t in hs.partition_rows(rp, world)] if world > 1 else [0, n]
    lo, hi = row_lo[rank], row_lo[rank + 1]
    e0, e1 = int(rp[lo]), int(rp[hi])
    rp_l = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
    ci_l, v_l = ci[e0:e1], v[e0:e1]
    gen_s = time.time() - t0
    A = hs.Matrix.from_csr(rp_l, ci_l, v_l, n_cols=n, keep_host=False)
    x = G.vector_x(n)
    if not args.no_hub:
        # columns renumbered by popularity in HBM (one-time set-up like Band-k): hot x entries become neighbours.  With
        # N > 1 every rank relabels by the counts of ITS row block, so the ranks' x numberings differ; x is set per rank.
        t_hub = time.time()
        perm_cols = A.relabel_hub_columns()
        t_hub = time.time() - t_hub
    par = parity_vs_numpy(A, hs, rp_l, ci_l, v_l, x, hs.ALGO_AUTO)
    m_loc, nnz_loc = hi - lo, e1 - e0
    stream = torch.cuda.current_stream().cuda_stream
    y_local = torch.zeros(m_loc, dtype=torch.float64, device="cuda")
    x_full = hd.wrap_device_ptr(A.x_ptr(), n, torch.float64)

    def compute():
        A.spmv(hs.ALGO_AUTO, d_x=x_full.data_ptr(), d_y=y_local.data_ptr(), stream=stream)

    rep = hd.ReplicatedSpMV(row_lo, rank, compute)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    tick = torch.zeros(1, device="cuda")

    def timed(fn, k):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.all_reduce(tick)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / k], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    for _ in range(warmup):
        rep.step()
    l0 = hs.launch_count()
    ms = timed(rep.step, steps)
    launches = hs.launch_count() - l0
    kernel = A.last_kernel()
    out = {"workload": "rmat%d_csr" % args.rmat_scale, "n": n, "nnz": nnz, "n_gpus": world,
           "x": "replicated" if world > 1 else "single GPU", "ms_per_step": ms,
           "gflops": 2.0 * nnz / (ms * 1e-3) / 1e9, "kernel": kernel, "parity": par,
           "generate_s": gen_s, "launches_per_step_per_rank": launches / max(steps, 1),
           "hub_column_relabel": (not args.no_hub), "hub_relabel_s": (t_hub if not args.no_hub else None)}
    b = alg_bytes(m_loc, n, nnz_loc)
    out["roofline"] = {"algorithmic_bytes_per_rank": b, "achieved_gbs_per_rank": b / (ms * 1e-3) / 1e9,
                       "frac_measured": b / (ms * 1e-3) / 1e9 / peak, "frac_nominal": b / (ms * 1e-3) / 1e9 / NOMINAL_GBS}
    if world > 1:
        perm_t = None if args.no_hub else torch.from_numpy(perm_cols.astype(np.int64)).cuda()
        y_all = torch.empty(n, dtype=torch.float64, device="cuda") if perm_t is not None else None

        def step_gather():
            rep.step()
            if perm_t is None:
                rep.all_gather_y(y_local, x_full)
            else:   # y blocks arrive in the original numbering; this rank's kernel wants x_kernel[i] = x[perm[i]]
                rep.all_gather_y(y_local, y_all)
                torch.index_select(y_all, 0, perm_t, out=x_full)
        x_keep = x_full.clone()
        for _ in range(2):
            step_gather()
            x_full.copy_(x_keep)
        ms_g = timed(step_gather, max(3, steps // 2))
        x_full.copy_(x_keep)
        out["all_gather_mode"] = {"ms_per_step": ms_g, "gflops": 2.0 * nnz / (ms_g * 1e-3) / 1e9,
                                  "what": "local SpMV + all_gather of the y blocks into every rank's x (y feeds the next x)"}
    bad = torch.tensor([par["rows_outside_bound"], par["rows_checked"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(bad)
    out["parity"] = {"rows_checked": int(bad[1]), "rows_outside_bound": int(bad[0]),
                     "max_err_over_scale_rank0": par["max_err_over_scale"], "against": "numpy product of every rank's row block"}
    if full_line:
        # e2e: x from pinned host memory up, kernel, the block's y down, every step
        x_host = torch.empty(n, dtype=torch.float64).pin_memory()
        x_host.copy_(torch.from_numpy(x))
        y_host = torch.empty(m_loc, dtype=torch.float64).pin_memory()
        A.spmv_host_ptr(x_host.data_ptr(), y_host.data_ptr(), hs.ALGO_AUTO)
        barrier()
        k = max(3, min(steps, 10))
        e2e_ms = A.spmv_host_timed(x_host.data_ptr(), y_host.data_ptr(), k, hs.ALGO_AUTO) / k
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": 2.0 * nnz / (float(t[0]) * 1e-3) / 1e9, "unit": "GFLOP/s",
                      "h2d_bytes_per_step": n * 8 * world, "d2h_bytes_per_step": n * 8, "steps": k}
    A.close()
    return out


# ----------------------------------------------------------------------------- our arm: the stencil
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import heterogeneous_spmv_b200 as hs
    from heterogeneous_spmv_b200 import dist as hd
    from heterogeneous_spmv_b200 import generators as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the SpMV path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    hs.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peak_gbs, peak_src = load_peaks()
    n_warm = max(args.warmup, 3)

    if args.workload == "rmat24":
        sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None)) if rank == 0 else None
        t_w0 = time.time()
        r = rmat_section(args, hs, hd, np, torch, dist, rank, world, peak_gbs, args.steps, n_warm, full_line=True)
        clocks = sampler.stop(t_w0, time.time()) if sampler else None
        if rank == 0:
            line = {"metric": "fp64 SpMV GFLOP/s", "value": r["gflops"], "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                    "warmup": n_warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": {"workload": r["workload"], "n": r["n"], "nnz": r["nnz"], "format": "csr", "kernel": r["kernel"], "x": r["x"],
                               "l2": "matrix stream 3.1 GB >> L2; x (134 MB) > L2 (126 MB): no flush needed"},
                    "roofline": {"bound": "hbm", "achieved": r["roofline"]["achieved_gbs_per_rank"], "peak": peak_gbs, "unit": "GB/s",
                                 "frac": r["roofline"]["frac_measured"], "traffic": None, "peak_source": peak_src,
                                 "frac_of_nominal_8TBs": r["roofline"]["frac_nominal"]},
                    "e2e": r["e2e"], "parity": r["parity"], "gpu_launches": int(r["launches_per_step_per_rank"] * args.steps * world),
                    "clocks": clocks}
            if "all_gather_mode" in r:
                line["all_gather_mode"] = r["all_gather_mode"]
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        if r["parity"]["rows_outside_bound"]:
            raise SystemExit("bench.py: R-MAT parity failed")
        return

    g = args.g
    n = g ** 3
    nnz_total = stencil27_before(n, g)
    algo = {"stream": hs.ALGO_CSR_STREAM, "vector": hs.ALGO_CSR_VECTOR, "merge": hs.ALGO_CSR_MERGE,
            "auto": hs.ALGO_AUTO}[args.algo]

    # ---- this rank's block
    fmt_k = {"csr": 1, "csr2": 2, "csr3": 3}[args.format]
    if fmt_k == 1:
        row_lo = partition_stencil(g, world)
        lo, hi = row_lo[rank], row_lo[rank + 1]
        if world > 1:   # column window of this block = the part of x it needs (owned rows + halo)
            probe = hs.Matrix.stencil(27, g, np.float64, 0, row_lo=lo, row_hi=hi)
            cmin, cmax = probe.col_range()
            probe.close()
        else:
            cmin, cmax = 0, n - 1
        A = hs.Matrix.stencil(27, g, np.float64, 0, row_lo=lo, row_hi=hi, col_shift=cmin, n_cols=cmax - cmin + 1)
    else:
        # CSR-k: every rank builds the same Band-k-reordered matrix in its own HBM (22 GB, deterministic), the
        # rows are cut nnz-balanced at super-(super-)row boundaries, and the rank keeps its shard
        full = hs.Matrix.stencil(27, g, np.float64, 0)
        sizes, _, _ = hs.auto_params(nnz_total, n, hs.ARCH_B200)
        full.build_csrk(fmt_k, sizes)
        maps = full.csrk()
        align = maps["sr_ptr"] if fmt_k == 2 else maps["sr_ptr"][maps["ssr_ptr"]]
        row_lo = [int(t) for t in hs.partition_rows(full.row_ptr(), world, align_ptr=align)]
        del maps, align
        lo, hi = row_lo[rank], row_lo[rank + 1]
        if world > 1:
            wide = full.slice_rows(lo, hi)
            cmin, cmax = wide.col_range()
            wide.close()
            A = full.slice_rows(lo, hi, cmin, cmax - cmin + 1)
            full.close()
        else:
            cmin, cmax = 0, n - 1
            A = full
        algo = hs.ALGO_CSRK
    if args.warps:
        A.set_option(hs.OPT_WARPS, args.warps)
    info = A.info()
    m_loc, nnz_loc = info["m"], info["nnz"]
    win = cmax - cmin + 1
    # x lives in the handle's own cudaMalloc'ed vector so that neighbours can map it (CUDA IPC)
    x_local = hd.wrap_device_ptr(A.x_ptr(), win, torch.float64)
    A.fill_x_synthetic(cmin)                                       # generators.vector_x, global index
    y_local = torch.full((m_loc,), float("nan"), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def compute(a, b, x_ptr=None):
        if fmt_k == 1:
            A.spmv_range(a, b, algo, d_x=x_ptr or x_local.data_ptr(), d_y=y_local.data_ptr(), stream=stream)
        else:   # the CSR-k kernel walks whole super-rows: the shard runs as one launch once its halo is in
            assert (a, b) == (0, m_loc)
            A.spmv(algo, d_x=x_ptr or x_local.data_ptr(), d_y=y_local.data_ptr(), stream=stream)

    exchange = args.exchange
    sh = None
    x_bufs = [x_local]
    my_segs = []            # local index ranges of x that I push to neighbours (they change every step)
    all_segs = []           # global index ranges that ANY rank pushes
    if world > 1:
        ranges = hd.gather_col_ranges((cmin, cmax))
        interior = A.interior_rows(lo - cmin, hi - cmin) if fmt_k == 1 else (0, 0)
        if exchange in ("peer", "peer-unfused") and not hd.halo_is_symmetric(row_lo, ranges):
            exchange = "nccl"   # the ack-less ping-pong needs senders == receivers (dist.PeerHalo)
        peer = fused = None
        if exchange in ("peer", "peer-unfused"):
            # second x buffer (ping-pong by step parity, see PeerHalo), also cudaMalloc'ed so that it can be mapped
            x2_ptr = hs.peer_alloc((win + 16) * 8)
            x2 = hd.wrap_device_ptr(x2_ptr, win, torch.float64)
            x2.copy_(x_local)
            x_bufs = [x_local, x2]
            peer = hd.PeerHalo(hs, [A.x_ptr(), x2_ptr], 8, row_lo, ranges, rank)
            if exchange == "peer" and fmt_k == 1:
                def fused(a, b, flags, nw, step, timeout, x_ptr):
                    A.spmv_halo(a, b, d_x=x_ptr, d_y=y_local.data_ptr(), wait_flags=flags, n_wait=nw,
                                wait_step=step, timed_out=timeout, stream=stream)
        sh = hd.ShardedSpMV(row_lo, ranges, rank, x_local, compute, interior, peer=peer, fused=fused)
        halo_bytes = sh.halo_elems * 8
        my_segs = merge_segments([(a - cmin, b - cmin) for _, a, b in sh.send])
        for r in range(world):
            _, send_r = hd.halo_plan(row_lo, ranges, r)
            all_segs += [(a, b) for _, a, b in send_r]
        all_segs = merge_segments(all_segs)
        # poison: nothing but a correct push may ever make a halo entry finite
        for xb in x_bufs:
            for _, a, b in sh.recv:
                xb[a - cmin: b - cmin] = float("nan")
        torch.cuda.synchronize()
        dist.barrier()
    else:
        halo_bytes = 0

    t_step = [0]     # steps run so far (= the scale index of the x the last step used)

    upd = [None]     # comm-stream event: the exchanged slices of the NEXT step's x buffer are written

    def write_next_x(t_next, xb, torch_stream):
        """x of step t_next: the slices that go to the neighbours change every step (one small launch)"""
        hs.fill_x_segments(xb.data_ptr(), np.float64, my_segs, cmin, step_scale(t_next), torch_stream.cuda_stream)

    def step():
        t_step[0] += 1
        if world == 1:
            compute(0, m_loc)
            return
        t = t_step[0]
        if args.no_exchange:   # diagnosis only: local kernels without the halo traffic (halo stays NaN; invalid as a result)
            compute(0, m_loc)
            return
        if sh.peer is not None and args.x_update == "main":
            sh.peer.reuse_buffer()
            write_next_x(t, x_bufs[(sh.peer.step + 1) % 2], torch.cuda.current_stream())
            sh.step(overlap=not args.no_overlap)
        elif sh.peer is not None:
            # The slices of step t's buffer were written on the COMM stream behind the pushes of step t-1 (below), i.e.
            # while the previous product was still running: the update of x is the test's scaffolding, not part of the
            # product, and stays off the main stream's critical path.  The main stream only waits for its event.
            if upd[0] is None:      # very first step: nothing was prepared yet
                sh.peer.reuse_buffer()
                write_next_x(t, x_bufs[(sh.peer.step + 1) % 2], torch.cuda.current_stream())
            else:
                torch.cuda.current_stream().wait_event(upd[0])
            sh.step(overlap=not args.no_overlap)
            # buffer of step t+1: last read by my pushes of step t-1 (earlier on this same comm stream) and by my kernel
            # of step t-1 (finished before this step's `ready` event, which the comm stream has already waited for)
            write_next_x(t + 1, x_bufs[(sh.peer.step + 1) % 2], sh.peer.comm)
            if upd[0] is None:
                upd[0] = torch.cuda.Event()
            upd[0].record(sh.peer.comm)
        else:
            write_next_x(t, x_local, torch.cuda.current_stream())
            sh.step(overlap=not args.no_overlap)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    tick = torch.zeros(1, device="cuda")

    def rendezvous():
        # The hosts leave dist.barrier() milliseconds apart; with a halo exchange the early ranks'
        # first kernels would spend that skew spinning on their neighbours' flags inside the timed
        # region.  A stream-ordered all-reduce completes on every GPU at the same moment (when the
        # last host has enqueued it), so every rank's start event is recorded together.
        if world > 1:
            dist.all_reduce(tick)

    # ---- expected result of the LAST step run so far, for the parity record
    x_ref = torch.empty(win + 16, dtype=torch.float64, device="cuda") if not args.no_parity else None
    y_ref = torch.empty(m_loc, dtype=torch.float64, device="cuda") if not args.no_parity else None

    def expected_x_into(buf):
        t = t_step[0]
        hs.fill_x_segments(buf.data_ptr(), np.float64, [(0, win)], cmin, 1.0, stream)
        if world > 1:
            mine = [(max(a, cmin) - cmin, min(b, cmax + 1) - cmin) for a, b in all_segs if a <= cmax and b > cmin]
            for i in range(0, len(mine), 8):
                hs.fill_x_segments(buf.data_ptr(), np.float64, mine[i:i + 8], cmin, step_scale(t), stream)

    def check_parity():
        """all rows of this rank's y (result of the last step) against (1) another kernel on a locally built copy
        of the expected x, (2) a numpy product on windows at both ends of the block and in its middle"""
        torch.cuda.synchronize()
        expected_x_into(x_ref)
        alt = hs.ALGO_CSR_VECTOR if fmt_k == 1 else hs.ALGO_CSRK
        if fmt_k == 1:
            A.spmv_range(0, m_loc, alt, d_x=x_ref.data_ptr(), d_y=y_ref.data_ptr(), stream=stream)
        else:
            A.spmv(alt, d_x=x_ref.data_ptr(), d_y=y_ref.data_ptr(), stream=stream)
        torch.cuda.synchronize()
        # classic values (26 / -1), x >= 1: sum_j |a_ij x_j| >= 26 + 7 on every row, so this uniform bound is
        # TIGHTER than north_star's 1e-12 * sum_j |a_ij x_j|
        bound = 1e-12 * 33.0
        err = (y_local - y_ref).abs()
        bad_dev = int((~(err <= bound)).sum().item())       # NaN counts as outside
        nan_y = int(torch.isnan(y_local).sum().item())
        rows_np, bad_np = 0, 0
        if fmt_k == 1:
            W = min(4096, m_loc)
            wins = sorted({(0, W), (max(0, m_loc // 2 - W // 2), min(m_loc, m_loc // 2 - W // 2 + W)), (m_loc - W, m_loc)})
            t = t_step[0]
            for a, b in wins:
                rp_w, ci_w, v_w = G.stencil27(g, row_lo=lo + a, row_hi=lo + b)
                c0, c1 = int(ci_w.min()), int(ci_w.max()) + 1
                xg = 1.0 + (np.arange(c0, c1, dtype=np.int64) % 97) / 128.0      # generators.vector_x on [c0, c1)
                if world > 1:
                    for sa, sb in all_segs:
                        u, w = max(sa, c0), min(sb, c1)
                        if u < w:
                            xg[u - c0: w - c0] *= step_scale(t)
                yw, sw = numpy_spmv(rp_w, ci_w - c0, v_w, xg)
                got = y_local[a:b].cpu().numpy()
                bad_np += int(np.count_nonzero(~(np.abs(got - yw) <= 1e-12 * sw)))
                rows_np += b - a
        return {"rows_checked": m_loc, "rows_outside_bound": bad_dev, "nan_rows": nan_y,
                "numpy_rows_checked": rows_np, "numpy_rows_outside_bound": bad_np}

    def reduce_parity(p):
        keys = ["rows_checked", "rows_outside_bound", "nan_rows", "numpy_rows_checked", "numpy_rows_outside_bound"]
        t = torch.tensor([float(p[k]) for k in keys], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return {k: int(t[i]) for i, k in enumerate(keys)}

    # ---- settle + verify BEFORE timing: a few steps, each checked (untimed, not part of the warm-up count)
    pre = None
    if not args.no_parity and not args.no_exchange:
        for _ in range(3 if world > 1 else 1):
            step()
            p = reduce_parity(check_parity())
            if pre is None:
                pre = p
            else:
                for k in p:
                    pre[k] += p[k]

    if world > 1 and os.environ.get("HSPMV_SPIN_STATS") and sh.peer is not None and not args.no_exchange:
        sh.peer.trace = []
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None)) if rank == 0 else None
    for _ in range(n_warm):
        step()
    # ---- timed: K steps, device events, max over ranks
    spin_diag = world > 1 and bool(os.environ.get("HSPMV_SPIN_STATS")) and sh.peer is not None
    if spin_diag:
        sh.peer.spin_reset()
    barrier()
    launches0 = hs.launch_count()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rendezvous()
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
    e1.record()
    torch.cuda.synchronize()
    t_wall_timed = time.time()
    ms = e0.elapsed_time(e1)
    launches = hs.launch_count() - launches0
    if world > 1 and os.environ.get("HSPMV_SPIN_STATS"):
        print("rank %d own ms/step %.4f (rows %d, nnz %d)" % (rank, ms / args.steps, m_loc, nnz_loc), flush=True)
    if world > 1 and sh.peer is not None and sh.peer.trace is not None:
        d = sh.peer.push_delays_ms()[-args.steps:] or [0.0]
        print("rank %d push completes %.3f ms (median) / %.3f (max) after the kernel's launch point" % (
            rank, sorted(d)[len(d) // 2], max(d)), flush=True)
        sh.peer.trace = None
    if spin_diag:
        mx_us, mean_us, hist = sh.peer.spin_stats()
        print("rank %d halo spin (timed steps): max %.1f us, mean %.1f us per warp-wait, histogram "
              "[<4,<17,<67,<267,<1067,more us] %s" % (rank, mx_us, mean_us, hist), flush=True)
    barrier()
    kernel_name = A.last_kernel()
    # ---- parity of the LAST TIMED step
    parity = None
    if not args.no_parity and not args.no_exchange:
        parity = reduce_parity(check_parity())
        parity.update({"halo_poisoned": world > 1, "x_changes_per_step": world > 1, "steps_run": t_step[0],
                       "pre_timing_steps_checked": 3 if world > 1 else 1,
                       "pre_timing_rows_outside_bound": pre["rows_outside_bound"] + pre["numpy_rows_outside_bound"] + pre["nan_rows"],
                       "bound": "device: |y - y_alt| <= 1e-12*33 (<= 1e-12*sum|a x| on every row); numpy windows: 1e-12*sum_j|a_ij x_j|",
                       "against": "%s kernel on a locally built expected x (all rows) + numpy product on 3 windows of 4096 rows per rank" % (
                           "csr_vector" if fmt_k == 1 else "the same CSR-k")})
    del x_ref, y_ref
    torch.cuda.empty_cache()

    # ---- e2e: the C-ABI call with HOST buffers, every step: x up from pinned memory, kernel, y down.
    # At N > 1 every rank uploads the window of x its rows reference ([cmin, cmax]: own slice + halo)
    # straight from host memory and downloads its rows of y -- the same pipelined call as at N = 1;
    # with x on the host there is nothing to exchange between the GPUs in this leg.
    affinity = bind_near_gpu(local_rank) if world > 1 else "n/a"
    A.fill_x_synthetic(cmin)
    torch.cuda.synchronize()
    x_host = torch.empty(win, dtype=torch.float64).pin_memory()
    x_host.copy_(x_local.cpu())
    y_host = torch.empty(m_loc, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    A.spmv_host_ptr(x_host.data_ptr(), y_host.data_ptr(), algo)          # warm-up (first touch of the pinned pages)
    e2e_mode = "pipelined"
    if world > 1:
        # All ranks copy through one host: whether overlapping the two PCIe directions (pipelined call) beats
        # strictly sequential copies depends on how much bandwidth the host side has left at this N (measured on
        # this pool: pipelined wins at N = 2, loses at N = 4).  Two untimed-for-the-record calls of each mode,
        # max over ranks, pick one mode for everybody, then measure.
        trial = []
        for mode in (0, 1):
            A.set_option(hs.OPT_NO_PIPELINE, mode)
            A.spmv_host_ptr(x_host.data_ptr(), y_host.data_ptr(), algo)
            barrier()
            trial.append(A.spmv_host_timed(x_host.data_ptr(), y_host.data_ptr(), 2, algo))
        t = torch.tensor(trial, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sequential = bool(t[1] < t[0])
        A.set_option(hs.OPT_NO_PIPELINE, 1 if sequential else 0)
        e2e_mode = "sequential" if sequential else "pipelined"
    barrier()
    rendezvous()
    torch.cuda.synchronize()
    e2e_ms = A.spmv_host_timed(x_host.data_ptr(), y_host.data_ptr(), e2e_steps, algo)
    chk = float(y_host[:1000].abs().sum())
    h2d_bytes = win * 8
    # PCIe floor at this N: every rank moves its window up and its rows down at the same time, nothing else
    barrier()
    rendezvous()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    s_up.wait_stream(torch.cuda.current_stream())
    s_dn.wait_stream(torch.cuda.current_stream())
    c0.record()
    s_up.wait_event(c0)
    s_dn.wait_event(c0)
    with torch.cuda.stream(s_up):
        x_local.copy_(x_host, non_blocking=True)
    with torch.cuda.stream(s_dn):
        y_host.copy_(y_local, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_up)
    torch.cuda.current_stream().wait_stream(s_dn)
    c1.record()
    torch.cuda.synchronize()
    pcie_floor_ms = c0.elapsed_time(c1)
    barrier()

    # ---- reduce over ranks
    stats = torch.tensor([ms, e2e_ms, float(launches), float(h2d_bytes), pcie_floor_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, launches, h2d_bytes, pcie_floor_ms = float(mx[0]), float(mx[1]), int(sm[2]), int(sm[3]), float(mx[4])
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None     # timed loop + e2e leg

    # ---- BASELINE configs[3] inside the same line (every N): R-MAT with replicated x
    rm = None
    if not args.no_rmat:
        A_keep = None
        if world == 1:      # the 22 GB stencil stays resident next to the 3.4 GB R-MAT matrix (180 GB of HBM)
            A_keep = A
        try:
            rm = rmat_section(args, hs, hd, np, torch, dist, rank, world, peak_gbs, max(5, min(args.steps, 20)), 3)
        except Exception as e:      # noqa: BLE001
            rm = {"error": "%s: %s" % (type(e).__name__, e)}
        del A_keep

    failed = False
    if rank == 0:
        ms_step = ms / args.steps
        value = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9
        e2e_value = 2.0 * nnz_total / (e2e_ms / e2e_steps * 1e-3) / 1e9
        # roofline of the dominant kernel: at N = 1 the step IS one launch of it; at N > 1 it is one fused launch
        # per rank (interior + boundary tiles), so per-rank bytes / step time is reported
        b_launch = alg_bytes(m_loc, win, nnz_loc)
        achieved = b_launch / (ms_step * 1e-3) / 1e9
        traffic = None
        if world == 1:      # ncu dram bytes of ONE launch at N = 1 (profiles/traffic.json); no capture at N > 1
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("stencil27_g%d" % g)
            except Exception:
                pass
        line = {
            "metric": "fp64 SpMV GFLOP/s", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "stencil27_g%d_%s" % (g, args.format), "n": n, "nnz": nnz_total, "format": args.format,
                       "kernel": kernel_name, "partition": "nnz-balanced row blocks" if world > 1 else "single GPU",
                       "halo_bytes_per_rank_per_step": halo_bytes, "overlap": world > 1 and not args.no_overlap and fmt_k == 1,
                       "exchange": ("none" if args.no_exchange else exchange) if world > 1 else "n/a",
                       "x_update_per_step": "exchanged slices of x rewritten every step (one small launch)" if world > 1 else "x static",
                       "l2": "input (21.9 GB) >> L2 (126 MB): no flush needed",
                       "start": "host barrier + synchronize, then a stream-ordered all-reduce so all ranks' start "
                                "events coincide" if world > 1 else "synchronize"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / NOMINAL_GBS,
                         "algorithmic_bytes_per_launch": b_launch},
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": n * 8,
                    "call": "hspmv_spmv_host on every rank's window of x (%s: %s)" % (
                        e2e_mode, "H2D pieces / row-chunk kernels / D2H chunks overlapped" if e2e_mode == "pipelined"
                        else "H2D, kernel, D2H in sequence -- faster than the overlapped form at this N on this host"),
                    "steps": e2e_steps, "checksum": chk, "ms_per_step": e2e_ms / e2e_steps,
                    "pcie_floor_ms": pcie_floor_ms, "pcie_floor_what": "max over ranks of: this rank's x window up and its y rows down, "
                                                                       "both directions at once, all ranks together, no kernel",
                    "host_affinity": affinity},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
        }
        if parity is not None:
            line["parity"] = parity
            failed = bool(parity["rows_outside_bound"] or parity["nan_rows"] or parity["numpy_rows_outside_bound"] or
                          parity["pre_timing_rows_outside_bound"])
        if rm is not None:
            line["rmat24"] = rm
            failed = failed or bool(rm.get("parity", {}).get("rows_outside_bound", 0))
        if world == 1 and not args.no_csrk and fmt_k == 1:
            # the same matrix in the reference's own formats: Band-k (coarse graph + reorderA in HBM, RCM of the
            # coarse graph on the host), then the CSR-k kernel; 10 launches back to back after 3 warm-ups
            formats = {"csr": {"gflops": value, "ms": ms_step}}
            for k in (2, 3):
                B = None
                try:
                    B = hs.Matrix.stencil(27, g, np.float64, 0)
                    sizes, _, _ = hs.auto_params(nnz_total, n, hs.ARCH_B200)
                    t0 = time.time()
                    B.build_csrk(k, sizes)
                    tb = time.time() - t0
                    B.fill_x_synthetic(0)
                    t = B.spmv_timed_batch(hs.ALGO_CSRK, warmup=3, iters=10)
                    i = B.info()
                    gbs = alg_bytes(n, n, nnz_total, 8, (i["n_sr"] + i["n_ssr"] + 2) * 4) / (t * 1e-3) / 1e9
                    formats["csr%d" % k] = {"gflops": 2.0 * nnz_total / (t * 1e-3) / 1e9, "ms": t, "sizes": sizes[: k - 1],
                                            "n_sr": i["n_sr"], "n_ssr": i["n_ssr"], "bandk_build_s": tb,
                                            "bandk_phases_s": build_phases(B), "gbs": gbs, "frac_measured": gbs / peak_gbs,
                                            "frac_nominal": gbs / NOMINAL_GBS, "kernel": B.last_kernel()}
                except Exception as e:      # noqa: BLE001  (an extra section, never the reason for a missing bench line)
                    formats["csr%d" % k] = {"error": "%s: %s" % (type(e).__name__, e)}
                finally:
                    if B is not None:
                        B.close()
            line["formats"] = formats
        A.close()
        if world == 1 and not args.no_configs:
            line["configs"] = config_section(hs, np, peak_gbs, args.no_cpu)
        if world == 1 and not args.no_cpu:
            cores = cpu_setup_threads()
            try:
                t, nnz_s, kind = cpu_stencil_sample(g, args.cpu_rows, 5, 20)
                line["cpu_baseline"] = {
                    "value": 2.0 * nnz_s / t / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
                    "sample": "rows [0,%d) of the same matrix (%d nnz), omp_spmv, mean of 20 after 5 warm-ups, "
                              "OMP_SCHEDULE=static" % (min(args.cpu_rows, n), nnz_s)}
            except Exception as e:      # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "GFLOP/s", "cores": cores, "kind": "unavailable",
                                        "sample": "%s: %s" % (type(e).__name__, e)}
            # the spmv.h compat entry point on pageable (malloc'ed) arrays: what an unmodified reference driver calls
            try:
                line["e2e_pageable_compat"] = pageable_compat(hs, np, G)
            except Exception as e:      # noqa: BLE001
                line["e2e_pageable_compat"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        flag = torch.tensor([1.0 if failed else 0.0], device="cuda")
        dist.all_reduce(flag)
        failed = bool(flag[0] > 0)
        if sh.peer is not None:
            if sh.peer.timed_out():
                raise SystemExit("bench.py: a halo flag wait timed out (a peer never arrived)")
            sh.peer.close()
        dist.destroy_process_group()
    if failed:
        raise SystemExit("bench.py: parity FAILED (see the `parity` record)")


def pageable_compat(hs, np, G, g2=200):
    """omp_spmv() of include/spmv.h (hspmv_omp_spmv_f64) on ordinary numpy (pageable) arrays, wall clock, on the
    27-point 200^3 stencil (8 M rows, 213 M nnz): matrix uploaded on the first call, x up / y down on every call."""
    rp, ci, v = _oracle().stencil27(g2)
    n = len(rp) - 1
    x = G.vector_x(n)
    y = np.zeros(n)
    hs.omp_spmv(n, n, len(ci), rp, ci, v, x, y)         # first call: upload + cache
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        hs.omp_spmv(n, n, len(ci), rp, ci, v, x, y)
        ts.append(time.perf_counter() - t0)
    hs.lib().hspmv_omp_spmv_cache_clear()
    t = sorted(ts)[len(ts) // 2]
    return {"workload": "stencil27_g%d_csr" % g2, "n": n, "nnz": int(len(ci)), "ms_per_call": t * 1e3,
            "gflops": 2.0 * len(ci) / t / 1e9, "what": "wall clock of omp_spmv() on pageable host arrays, median of 5 (x up, kernel, y down)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="stencil400", choices=["stencil400", "rmat24"],
                    help="stencil400 = BASELINE configs[4] (the headline); rmat24 = configs[3] (replicated x when N > 1)")
    ap.add_argument("--g", type=int, default=G_DEFAULT, help="grid edge of the 27-point stencil (400 = BASELINE config 5)")
    ap.add_argument("--rmat-scale", type=int, default=24)
    ap.add_argument("--algo", default="stream", choices=["stream", "vector", "merge", "auto"])
    ap.add_argument("--cpu-rows", type=int, default=CPU_SAMPLE_ROWS)
    ap.add_argument("--format", default="csr", choices=["csr", "csr2", "csr3"],
                    help="csr (default, the headline), or the Band-k-reordered matrix as CSR-2 / CSR-3 (sharded at "
                         "super-row / super-super-row boundaries when N > 1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs")
    ap.add_argument("--no-csrk", action="store_true", help="skip the CSR-2 / CSR-3 timings of the same matrix (N = 1 only)")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1-cfg3 section (N = 1 only)")
    ap.add_argument("--no-rmat", action="store_true", help="skip the R-MAT (cfg4) section")
    ap.add_argument("--no-hub", action="store_true", help="R-MAT: skip the hub-column relabel (hspmv_relabel_hub_columns)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity record (diagnosis only)")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--x-update", default="main", choices=["main", "comm"],
                    help="N > 1: where the per-step rewrite of the exchanged x slices is enqueued: main = on the product's stream, in "
                         "front of the kernel (default); comm = on the comm stream behind the previous step's pushes")
    ap.add_argument("--warps", type=int, default=0, help="diagnosis: warps per CTA of the tile engine (0 = the library's choice)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer-unfused", "nccl"],
                    help="halo exchange: peer = copy-engine NVLink copies into IPC-mapped neighbour buffers + one fused "
                         "kernel per step that waits on the halo flags itself (default); peer-unfused = same transport, "
                         "separate interior/boundary launches; nccl = batch_isend_irecv")
    ap.add_argument("--no-exchange", action="store_true", help="diagnosis: skip the halo exchange (invalid as a result)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # launched without torchrun: re-launch ourselves under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    # before anything loads libgomp: the CPU legs use every host core
    cpu_setup_threads(force=(args.impl == "reference" or world == 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
