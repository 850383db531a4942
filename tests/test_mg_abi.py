"""Multi-GPU entry points of the C ABI (hspmv_mg_*, csrc/mg.cu; SURVEY.md 8b "multi-GPU init (device list) +
partition query"): through ctypes, through a plain C program, and through the drop-in drivers' --gpus flag.  Runs
with as many devices as are visible (1, 2, 3, 4); every step uses a new x, halos start as NaN, y is compared with
the oracle within the north_star bound."""
import os
import subprocess

import numpy as np
import pytest

from conftest import assert_within_bound

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIB = os.path.join(ROOT, "heterogeneous-spmv_b200", "lib")
BIN = os.path.join(ROOT, "heterogeneous-spmv_b200", "bin")


def _build_c(tmp):
    exe = os.path.join(str(tmp), "driver_mg")
    r = subprocess.run(["/usr/bin/gcc", "-O2", "-std=c99", "-Wall", "-Werror", "-I" + INC, os.path.join(ROOT, "tests", "compat", "driver_mg.c"),
                        "-o", exe, "-L" + LIB, "-lhspmv", "-Wl,-rpath," + LIB, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_using_the_mg_abi_compiles_and_links(tmp_path, hs):
    exe = _build_c(tmp_path)
    r = subprocess.run(["ldd", exe], capture_output=True, text=True)
    assert "libhspmv.so" in r.stdout and "not found" not in r.stdout


def _x_of_step(n, t):
    i = np.arange(n, dtype=np.int64)
    return 1.0 + ((i * (t + 3) + 11 * t) % 97) / 128.0


def _device_counts(hs):
    have = hs.device_count()
    return [d for d in (1, 2, 3, 4) if d <= have]


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 2, 3])
def test_mg_banded_matrix_changing_x(k, hs, oracle, gen):
    rp, ci, v = gen.stencil27(36, mode="varied")
    n = len(rp) - 1
    for nd in _device_counts(hs):
        M = hs.MultiGPU.from_csr(list(range(nd)), rp, ci, v, k=k, sup_row_sizes=[16, 8])
        info, part = M.info(), M.partition()
        assert info["n_dev"] == nd and info["k"] == k and not info["replicated"]
        assert part["row_lo"][0] == 0 and part["row_lo"][-1] == n and int(part["nnz"].sum()) == len(ci)
        if nd > 1:
            assert np.all(part["halo_elems"] > 0)
            assert part["nnz"].max() - part["nnz"].min() <= 0.1 * len(ci) / nd + 27 * 16 * 9       # nnz-balanced (CSR-k: up to a super-(super-)row)
        if k > 1:
            assert np.array_equal(M.perm(), oracle.bandk_build(rp, ci, k, [16, 8])["perm"])
        M.poison_halos()
        for t in range(1, 7):
            x = _x_of_step(n, t)
            M.set_x(x)
            M.spmv()
            y = M.get_y()
            assert_within_bound(y, oracle.csr_spmv(rp, ci, v, x), oracle.abs_rowsum(rp, ci, v, x), np.float64,
                                "mg k=%d devices=%d step %d [%s]" % (k, nd, t, M.shard_kernel(0)))
        # static x, several steps without set_x in between: both ping-pong buffers hold the owned slice
        ms = M.spmv_timed(2, 3)
        assert ms.shape == (3,) and np.all(ms > 0) and M.spmv_timed_batch(1, 4) > 0
        assert_within_bound(M.get_y(), oracle.csr_spmv(rp, ci, v, x), oracle.abs_rowsum(rp, ci, v, x), np.float64, "after timed runs")
        M.close()


@pytest.mark.gpu
def test_mg_power_law_matrix_runs_with_replicated_x(hs, oracle, gen):
    rp, ci, v = gen.rmat(14, 8)
    n = len(rp) - 1
    for nd in _device_counts(hs):
        M = hs.MultiGPU.from_csr(list(range(nd)), rp, ci, v)
        assert M.info()["replicated"] == (nd > 1)
        for t in (1, 2, 3):
            x = _x_of_step(n, t)
            M.set_x(x)
            M.spmv()
            assert_within_bound(M.get_y(), oracle.csr_spmv(rp, ci, v, x), oracle.abs_rowsum(rp, ci, v, x), np.float64,
                                "mg rmat devices=%d step %d" % (nd, t))
        M.close()


@pytest.mark.gpu
def test_mg_stencil_generated_in_hbm_and_fp32(hs, oracle, gen):
    g = 30
    rp, ci, v = gen.stencil27(g, mode="varied")
    n = g ** 3
    for dtype in (np.float64, np.float32):
        for nd in _device_counts(hs):
            M = hs.MultiGPU.stencil(list(range(nd)), 27, g, dtype, 1)
            x = gen.vector_x(n, dtype)
            M.poison_halos()
            M.set_x(x)
            M.spmv()
            vv = v.astype(dtype)
            assert_within_bound(M.get_y(), oracle.csr_spmv(rp, ci, vv, x), oracle.abs_rowsum(rp, ci, vv, x), dtype, "mg stencil nd=%d" % nd)
            M.close()


@pytest.mark.gpu
def test_c_program_and_drivers_on_all_visible_gpus(tmp_path, hs, gen):
    exe = _build_c(tmp_path)
    rp, ci, v = gen.stencil27(24, mode="varied")
    f0, f1 = str(tmp_path / "a0.mtx.rcm.csr"), str(tmp_path / "a1.mtx.rcm.csr")
    gen.write_csr_text(f0, rp, ci, v, base=0)
    gen.write_csr_text(f1, rp, ci, v, base=1)
    for nd in _device_counts(hs):
        for k in (1, 2, 3):
            r = subprocess.run([exe, "28", str(nd), str(k)], capture_output=True, text=True, timeout=120)
            assert r.returncode == 0 and "Number Wrong: 0" in r.stdout, r.stdout + r.stderr
        if nd == 1:
            continue
        r = subprocess.run([os.path.join(BIN, "cuda-spmv-csr", "spmv.exe"), f0, "5", "--gpus", str(nd)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "Number Wrong: 0" in r.stdout and "TimeMin:" in r.stdout and "Multi-GPU: %d devices" % nd in r.stdout, r.stdout + r.stderr
        for args in (["16"], ["16", "8"]):
            r = subprocess.run([os.path.join(BIN, "cuda-spmv-csrk", "cuda", "spmv"), f1, "5"] + args + ["--gpus", str(nd)], capture_output=True, text=True, timeout=120)
            assert r.returncode == 0 and "Number Wrong: 0" in r.stdout and "TimeAvg:" in r.stdout, r.stdout + r.stderr
