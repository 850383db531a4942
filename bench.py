#!/usr/bin/env python
"""bench.py -- the reference's headline measurement on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): fp64 SpMV GFLOP/s (2*nnz/t) with the HBM roofline beside it.
Workload: BASELINE.json configs[4] -- the 3-D 27-point stencil on a 400^3 grid (n = 64,000,000,
nnz = 1,719,374,392, 21.9 GB of algorithmic traffic per SpMV), the configuration the
"1/2/4/8 B200" part of the metric is quoted on; it fits one GPU, so N = 1 runs the same matrix
and N > 1 is strong scaling of it (nnz-balanced row blocks + x halo exchange over NVLink).
A step is one y = A*x over the whole matrix.  The matrix is far larger than L2 (126 MB), so no
flush is needed between timed iterations.

One JSON line on stdout (rank 0).  `value` = device-timed (CUDA events, max over ranks) with
inputs resident in HBM; `e2e` = the same metric through the C-ABI call with HOST buffers
(hspmv_spmv_host: H2D of x, kernel, D2H of y inside the timed region).
`--impl reference` times the reference's own OpenMP spmv-csr (oracle/_ref, compiled from the
unmodified sources; the oracle port if that is missing) on the host cores, on a bounded sample
of the same matrix.
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
CPU_SAMPLE_ROWS = 4_000_000   # bounded sample for the CPU arms: first 4 M rows (25 z-planes of 400^2)


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


def alg_bytes(m, n, nnz, vs=8):
    """nnz*(V+4) + (m+1)*4 + n*V (x once) + m*V (y)."""
    return nnz * (vs + 4) + (m + 1) * 4 + n * vs + m * vs


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_lo, t_hi):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9 or not (t_lo - 0.05 <= ts <= t_hi + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v == "Active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


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


def cpu_sample_arrays(g, rows):
    """CSR arrays of rows [0,rows) of the stencil, fetched from the device generator when a GPU is
    present (fast), else built with numpy."""
    import numpy as np
    import torch
    rows = min(rows, g ** 3)
    n_cols = min(g ** 3, rows + g * g + g + 1)
    if torch.cuda.is_available():
        import heterogeneous_spmv_b200 as hs
        A = hs.Matrix.stencil(27, g, np.float64, 0, row_lo=0, row_hi=rows, col_shift=0, n_cols=n_cols)
        rp, ci, v = A.download()
        A.close()
    else:
        from heterogeneous_spmv_b200 import generators as G
        rp, ci, v = G.stencil27(g, row_lo=0, row_hi=rows)
    from heterogeneous_spmv_b200 import generators as G
    return rp, ci, v, G.vector_x(n_cols), n_cols


def cpu_time_sample(g, rows, warmup, iters):
    """(seconds per SpMV (min), nnz of the sample, kind).  Reference omp_spmv if oracle/_ref is
    there, else the oracle's OpenMP port."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    rp, ci, v, x, n_cols = cpu_sample_arrays(g, rows)
    if O.have_ref():
        t, _ = _ref_timed(O, rp, ci, v, x, warmup, iters)
        kind = "reference"
    else:
        _force_omp_threads()
        t, _ = O.csr_spmv_timed(rp, ci, v, x, warmup, iters)   # (min; the port has no mean output)
        kind = "port"
    return t, int(len(ci)), kind


def _ref_timed(O, rp, ci, v, x, warmup, iters):
    # the reference's omp_spmv only ever indexes x by column id, so a rectangular sample
    # (rows [0,R) with columns up to R + g^2+g+1) is a valid input for it
    import ctypes as C
    import numpy as np
    L = O.ref("csr", np.float64)
    _force_omp_threads()
    y = np.zeros(len(rp) - 1)
    fp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
    ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.ref_csr_spmv_timed.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, fp, fp, fp, C.c_int, C.c_int,
                                     C.POINTER(C.c_double)]
    L.ref_csr_spmv_timed.restype = C.c_double
    avg = C.c_double()
    L.ref_csr_spmv_timed(len(rp) - 1, len(x), len(ci), np.ascontiguousarray(rp, np.int32),
                         np.ascontiguousarray(ci, np.int32), v, x, y, warmup, iters, C.byref(avg))
    return float(avg.value), y      # mean over the timed calls, like the GPU arm's K-step mean


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation on the host cores."""
    if rank != 0:
        return
    cores = cpu_setup_threads(force=True)
    t, nnz, kind = cpu_time_sample(args.g, args.cpu_rows, args.warmup, args.steps)
    gf = 2.0 * nnz / t / 1e9
    sample = "rows [0,%d) of stencil27 g=%d: %d nnz; mean of %d timed omp_spmv calls after %d warm-ups" % (
        min(args.cpu_rows, args.g ** 3), args.g, nnz, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": "fp64 SpMV GFLOP/s", "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "stencil27_g%d_csr" % args.g, "n": args.g ** 3,
                   "nnz": stencil27_before(args.g ** 3, args.g), "format": "csr"},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def config2_cpu_reference(g2=100, sizes=(8, 16, 32, 64)):
    """BASELINE configs[1], CPU side: the compiled reference's CSRk_Graph::SpMV k = 2 (spmv-csrk/csrk.cpp:165-223) on
    the 27-point g2^3 stencil, best over a few super-row sizes, all host threads; None without oracle/_ref."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    if not O.have_ref():
        return None
    from heterogeneous_spmv_b200 import generators as G
    cores = cpu_setup_threads()
    rp, ci, v = G.stencil27(g2)
    x = G.vector_x(len(rp) - 1, np.float64)
    O.ref("csrk", np.float64)
    _force_omp_threads()
    best = None
    for s in sizes:
        R = O.RefCSRk(rp, ci, v, 2, [s])
        t, _ = R.spmv_timed(x, 5, 20)
        R.close()
        if best is None or t < best[0]:
            best = (t, s)
    return {"gflops": 2.0 * len(ci) / best[0] / 1e9, "ms": best[0] * 1e3, "sizes": [best[1]], "cores": cores,
            "kind": "reference", "sizes_tried": list(sizes)}


def config2_line(hs, np, no_cpu):
    """BASELINE configs[1]: CSR-k k = 2 on the 27-point 100^3 stencil with a tuned super-row size on one B200 (Band-k
    built in HBM, kernel shape autotuned, 20 launches back to back), beside the reference's OpenMP spmv-csrk."""
    g2 = 100
    n2, nnz2 = g2 ** 3, (3 * g2 - 2) ** 3
    best = None
    for s in (8, 16, 32):
        B = hs.Matrix.stencil(27, g2, np.float64, 0)
        B.build_csrk(2, [s])
        B.fill_x_synthetic(0)
        B.autotune(hs.ALGO_CSRK, 1)
        t = B.spmv_timed_batch(hs.ALGO_CSRK, warmup=5, iters=20)
        i = B.info()
        if best is None or t < best["ms"]:
            best = {"ms": t, "sizes": [s], "n_sr": i["n_sr"], "kernel": B.last_kernel()}
        B.close()
    best["gflops"] = 2.0 * nnz2 / (best["ms"] * 1e-3) / 1e9
    best["gbs"] = (nnz2 * 12 + (n2 + 1) * 4 + 2 * n2 * 8 + (best["n_sr"] + 1) * 4) / (best["ms"] * 1e-3) / 1e9
    out = {"workload": "stencil27_g100_csr2", "n": n2, "nnz": nnz2, "ours": best}
    if not no_cpu:
        out["cpu_reference"] = config2_cpu_reference(g2)
    return out


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import heterogeneous_spmv_b200 as hs
    from heterogeneous_spmv_b200 import dist as hd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the SpMV path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    hs.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = args.g
    n = g ** 3
    nnz_total = stencil27_before(n, g)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
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
    info = A.info()
    m_loc, nnz_loc = info["m"], info["nnz"]
    # x lives in the handle's own cudaMalloc'ed vector so that neighbours can map it (CUDA IPC)
    x_local = hd.wrap_device_ptr(A.x_ptr(), cmax - cmin + 1, torch.float64)
    A.fill_x_synthetic(cmin)                                       # generators.vector_x, global index
    y_local = torch.zeros(m_loc, dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def compute(a, b, x_ptr=None):
        if fmt_k == 1:
            A.spmv_range(a, b, algo, d_x=x_ptr or x_local.data_ptr(), d_y=y_local.data_ptr(), stream=stream)
        else:   # the CSR-k kernel walks whole super-rows: the shard runs as one launch once its halo is in
            assert (a, b) == (0, m_loc)
            A.spmv(algo, d_x=x_ptr or x_local.data_ptr(), d_y=y_local.data_ptr(), stream=stream)

    if world > 1:
        ranges = hd.gather_col_ranges((cmin, cmax))
        interior = A.interior_rows(lo - cmin, hi - cmin) if fmt_k == 1 else (0, 0)
        peer = fused = None
        x_bufs = [x_local]
        if args.exchange in ("peer", "peer-unfused"):
            # second x buffer (ping-pong by step parity, see PeerHalo), also cudaMalloc'ed so that it can be mapped
            x2_ptr = hs.peer_alloc((cmax - cmin + 1 + 16) * 8)
            x2 = hd.wrap_device_ptr(x2_ptr, cmax - cmin + 1, torch.float64)
            x2.copy_(x_local)
            x_bufs = [x_local, x2]
            peer = hd.PeerHalo(hs, [A.x_ptr(), x2_ptr], 8, row_lo, ranges, rank)
            if args.exchange == "peer" and fmt_k == 1:
                def fused(a, b, flags, n, step, timeout, x_ptr):
                    A.spmv_halo(a, b, d_x=x_ptr, d_y=y_local.data_ptr(), wait_flags=flags, n_wait=n,
                                wait_step=step, timed_out=timeout, stream=stream)
        sh = hd.ShardedSpMV(row_lo, ranges, rank, x_local, compute, interior, peer=peer, fused=fused)
        if args.no_exchange:   # diagnosis only: local kernels without the halo traffic (x is static)
            step = lambda: compute(0, m_loc)                  # noqa: E731
        else:
            step = lambda: sh.step(overlap=not args.no_overlap)   # noqa: E731
        halo_bytes = sh.halo_elems * 8
    else:
        step = lambda: compute(0, m_loc)                      # noqa: E731
        halo_bytes = 0

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

    if world > 1 and os.environ.get("HSPMV_SPIN_STATS") and sh.peer is not None and not args.no_exchange:
        sh.peer.trace = []
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n_warm = max(args.warmup, 3) if world == 1 else max(args.warmup, 10)   # NCCL P2P channels settle late
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

    # ---- e2e: the C-ABI call with HOST buffers, every step: x up from pinned memory, kernel, y down.
    # At N > 1 every rank uploads the window of x its rows reference ([cmin, cmax]: own slice + halo)
    # straight from host memory and downloads its rows of y -- the same pipelined call as at N = 1;
    # with x on the host there is nothing to exchange between the GPUs in this leg.
    win = cmax - cmin + 1
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
    t_wall1 = time.time()
    barrier()

    # ---- reduce over ranks
    stats = torch.tensor([ms, e2e_ms, float(launches), float(h2d_bytes)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, launches, h2d_bytes = float(mx[0]), float(mx[1]), int(sm[2]), int(sm[3])
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    if rank == 0:
        ms_step = ms / args.steps
        value = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9
        e2e_value = 2.0 * nnz_total / (e2e_ms / e2e_steps * 1e-3) / 1e9
        # roofline of the dominant kernel: at N = 1 the step IS one launch of it; at N > 1 it is the
        # interior launch (plus two thin boundary launches), so per-rank bytes / step time is reported
        b_launch = alg_bytes(m_loc, cmax - cmin + 1, nnz_loc)
        achieved = b_launch / (ms_step * 1e-3) / 1e9
        traffic = None
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
                       "exchange": ("none" if args.no_exchange else args.exchange) if world > 1 else "n/a",
                       "l2": "input (21.9 GB) >> L2 (126 MB): no flush needed",
                       "start": "host barrier + synchronize, then a stream-ordered all-reduce so all ranks' start "
                                "events coincide" if world > 1 else "synchronize"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": b_launch},
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": n * 8,
                    "call": "hspmv_spmv_host on every rank's window of x (%s: %s)" % (
                        e2e_mode, "H2D pieces / row-chunk kernels / D2H chunks overlapped" if e2e_mode == "pipelined"
                        else "H2D, kernel, D2H in sequence -- faster than the overlapped form at this N on this host"),
                    "steps": e2e_steps, "checksum": chk},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
        }
        if world == 1 and not args.no_csrk and fmt_k == 1:
            # the same matrix in the reference's own formats: Band-k (coarse graph + reorderA in HBM, RCM of the
            # coarse graph on the host), then the CSR-k kernel; 10 launches back to back after 3 warm-ups
            formats = {"csr": {"gflops": value, "ms": ms_step}}
            for k in (2, 3):
                B = hs.Matrix.stencil(27, g, np.float64, 0)
                sizes, _, _ = hs.auto_params(nnz_total, n, hs.ARCH_B200)
                t0 = time.time()
                B.build_csrk(k, sizes)
                tb = time.time() - t0
                B.fill_x_synthetic(0)
                t = B.spmv_timed_batch(hs.ALGO_CSRK, warmup=3, iters=10)
                i = B.info()
                formats["csr%d" % k] = {"gflops": 2.0 * nnz_total / (t * 1e-3) / 1e9, "ms": t, "sizes": sizes[: k - 1],
                                        "n_sr": i["n_sr"], "n_ssr": i["n_ssr"], "bandk_build_s": tb,
                                        "gbs": (alg_bytes(n, n, nnz_total) + (i["n_sr"] + i["n_ssr"] + 2) * 4) / (t * 1e-3) / 1e9,
                                        "kernel": B.last_kernel()}
                B.close()
            line["formats"] = formats
            try:    # an extra, never the reason for a missing bench line
                line["config2"] = config2_line(hs, np, args.no_cpu)
            except Exception as e:      # noqa: BLE001
                line["config2"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if world == 1 and not args.no_cpu:
            cores = cpu_setup_threads()
            t, nnz_s, kind = cpu_time_sample(g, args.cpu_rows, 5, 20)
            line["cpu_baseline"] = {
                "value": 2.0 * nnz_s / t / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
                "sample": "rows [0,%d) of the same matrix (%d nnz), omp_spmv, mean of 20 after 5 warm-ups, "
                          "OMP_SCHEDULE=static" % (min(args.cpu_rows, n), nnz_s)}
        print(json.dumps(line), flush=True)
    if world > 1:
        if sh.peer is not None:
            if sh.peer.timed_out():
                raise SystemExit("bench.py: a halo flag wait timed out (a peer never arrived)")
            sh.peer.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--g", type=int, default=G_DEFAULT, help="grid edge of the 27-point stencil (400 = BASELINE config 5)")
    ap.add_argument("--algo", default="stream", choices=["stream", "vector", "merge", "auto"])
    ap.add_argument("--cpu-rows", type=int, default=CPU_SAMPLE_ROWS)
    ap.add_argument("--format", default="csr", choices=["csr", "csr2", "csr3"],
                    help="csr (default, the headline), or the Band-k-reordered matrix as CSR-2 / CSR-3 (sharded at "
                         "super-row / super-super-row boundaries when N > 1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-csrk", action="store_true", help="skip the CSR-2 / CSR-3 timings of the same matrix (N = 1 only)")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "peer-unfused", "nccl"],
                    help="halo exchange: peer = copy-engine NVLink copies into IPC-mapped neighbour buffers + one fused "
                         "kernel per step that waits on the halo flags itself (default); peer-unfused = same transport, "
                         "separate interior/boundary launches; nccl = batch_isend_irecv")
    ap.add_argument("--no-exchange", action="store_true", help="diagnosis: skip the halo exchange (invalid as a result)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
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
