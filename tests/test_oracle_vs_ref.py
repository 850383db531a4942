"""Pins the oracle: our restatement (oracle/spmv_oracle.c, bandk_oracle.cpp) against the REAL
reference compiled from /root/reference (oracle/_ref) and against the committed golden vectors
that the same reference produced (tests/golden/make_golden.py).  CPU only."""
import glob
import os
import zlib

import numpy as np
import pytest

from conftest import assert_within_bound

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def load_case(path, gen, dtype):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(path), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name = os.path.basename(path)[:-4]
    g, kw, k, sizes = mg.CASES[name]
    rp, ci, v = getattr(gen, g)(dtype=dtype, **kw)
    return np.load(path), rp, ci, v, k, sizes


def test_appendix_b_known_answer(oracle, gen):
    """CRC-32s produced by the compiled reference at survey time."""
    rp, ci, v = gen.laplace2d_5pt(30, dtype=np.float32)
    o = oracle.bandk_build(rp, ci, 3, [7, 8])
    assert len(o["ssr_ptr"]) - 1 == 16 and len(o["sr_ptr"]) - 1 == 124
    assert list(o["ssr_ptr"]) == [0, 2, 11, 19, 27, 35, 43, 51, 59, 67, 75, 83, 91, 99, 107, 115, 124]
    assert list(o["perm"][:12]) == [863, 864, 865, 866, 867, 868, 869, 897, 898, 899, 842, 843]
    r2, c2, _ = oracle.bandk_reorder(o["perm"], rp, ci, v)
    got = ["%08x" % crc(a.astype(np.uint32)) for a in (o["ssr_ptr"], o["sr_ptr"], r2, c2, o["perm"])]
    assert got == ["a31ed77e", "d695efc5", "7196ebf4", "ee8c01be", "b23f1be8"]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_oracle_matches_golden(path, dtype, oracle, gen):
    g, rp, ci, v, k, sizes = load_case(path, gen, dtype)
    assert crc(rp) == g["crc_in_row_ptr"] and crc(ci) == g["crc_in_col"], "generator drifted from the golden inputs"
    tag = "f32" if dtype == np.float32 else "f64"
    o = oracle.bandk_build(rp, ci, k, sizes)
    assert np.array_equal(o["perm"], g["perm"])
    assert np.array_equal(o["sr_ptr"], g["sr_ptr"])
    if k == 3:
        assert np.array_equal(o["ssr_ptr"], g["ssr_ptr"])
    r2, c2, v2 = oracle.bandk_reorder(o["perm"], rp, ci, v)
    assert crc(r2) == g["crc_row_ptr"] and crc(c2) == g["crc_col"] and crc(v2) == g["crc_val_" + tag]
    x = gen.vector_x(len(rp) - 1, dtype)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    y_csr = oracle.csr_spmv(rp, ci, v, x)
    assert_within_bound(y_csr, g["y_csr_" + tag], scale, dtype, "csr restatement vs reference omp_spmv")
    maps = [o["sr_ptr"]] + ([o["ssr_ptr"]] if k == 3 else [])
    y_k = oracle.csrk_spmv(k, maps, r2, c2, v2, x[o["perm"]])
    assert_within_bound(y_k, g["y_csrk_" + tag], scale[o["perm"]], dtype, "csr-k restatement vs reference SpMV")
    # the reference's own check (spmv-csrk/spmv.cpp:203): y[i] == y_csr[perm[i]]
    assert_within_bound(y_k, y_csr[o["perm"]], scale[o["perm"]], dtype, "csr-k vs csr[perm]")


def _need_ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (only buildable where /root/reference exists)")


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["f32", "f64"])
def test_oracle_matches_live_reference(dtype, oracle, gen):
    """Fresh inputs (not in the golden set) through the reference .so and through the restatement."""
    _need_ref(oracle)
    cases = [(gen.stencil27(9, dtype=dtype, mode="varied"), 3, [12, 6]),
             (gen.laplace2d_5pt(41, dtype=dtype, mode="varied", rcm=True), 2, [24]),
             (gen.banded_shell(3000, dtype=dtype, band=90, seed=5), 3, [7, 5]),
             (gen.random_ragged(700, dtype=dtype, diag=True, seed=3), None, None)]
    for (rp, ci, v), k, sizes in cases:
        x = gen.vector_x(len(rp) - 1, dtype)
        scale = oracle.abs_rowsum(rp, ci, v, x)
        y = oracle.csr_spmv(rp, ci, v, x)
        assert_within_bound(y, oracle.ref_csr_spmv(rp, ci, v, x), scale, dtype, "omp_spmv")
        assert_within_bound(y, oracle.ref_csr_spmv(rp, ci, v, x, serial=True), scale, dtype, "test_spmv")
        assert_within_bound(oracle.csr_spmv(rp, ci, v, x, parallel=True), y, scale, dtype, "omp port")
        if k is None:
            continue
        R = oracle.RefCSRk(rp, ci, v, k, sizes)
        o = oracle.bandk_build(rp, ci, k, sizes)
        maps = R.maps()
        perm, r2, c2, v2 = R.matrix()
        assert np.array_equal(perm, o["perm"]) and np.array_equal(maps[0], o["sr_ptr"])
        if k == 3:
            assert np.array_equal(maps[1], o["ssr_ptr"])
        q = oracle.bandk_reorder(o["perm"], rp, ci, v)
        assert np.array_equal(q[0], r2) and np.array_equal(q[1], c2) and np.array_equal(q[2], v2)
        y_ref = R.spmv(x)
        omaps = [o["sr_ptr"]] + ([o["ssr_ptr"]] if k == 3 else [])
        assert_within_bound(oracle.csrk_spmv(k, omaps, r2, c2, v2, x[perm]), y_ref, scale[perm], dtype, "csrk")
        R.close()


def test_text_reader_matches_reference_reader(tmp_path, oracle, gen):
    _need_ref(oracle)
    rp, ci, v = gen.stencil27(6, dtype=np.float32, mode="varied")
    p0 = str(tmp_path / "a.mtx.rcm.csr")
    gen.write_csr_text(p0, rp, ci, v, base=0)
    m, n, r_rp, r_ci, r_v = oracle.ref_read_csr_f32(p0)
    om, on, o_rp, o_ci, o_v = oracle.read_csr_text(p0, 0)
    assert (m, n) == (om, on) == (216, 216)
    assert np.array_equal(r_rp, o_rp) and np.array_equal(r_ci, o_ci)
    assert np.array_equal(r_v, o_v.astype(np.float32)) and np.array_equal(r_v, v)
    p1 = str(tmp_path / "b.mtx.rcm.csr")
    gen.write_csr_text(p1, rp, ci, v, base=1)
    _, _, b_rp, b_ci, b_v = oracle.read_csr_text(p1, 1)
    assert np.array_equal(b_rp, rp) and np.array_equal(b_ci, ci) and np.array_equal(b_v.astype(np.float32), v)
