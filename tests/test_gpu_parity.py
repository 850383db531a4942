"""Parity of the CUDA path against the oracle, through the C ABI (ctypes -> libhspmv.so).
Run on the B200 box: python -m pytest tests -m gpu.

Bar (BASELINE.json north_star): index work bit-exact; |y - y_ref|_i <= 1e-12 * sum_j |a_ij x_j|
in fp64, 1e-5 in fp32 (conftest.assert_within_bound)."""
import glob
import os

import numpy as np
import pytest

from conftest import assert_within_bound

pytestmark = pytest.mark.gpu

DT = [np.float64, np.float32]
DT_IDS = ["f64", "f32"]


def check_csr(hs, oracle, gen, rp, ci, v, algos, x=None, opts=None, what=""):
    dtype = v.dtype
    n = len(rp) - 1
    x = gen.vector_x(n, dtype) if x is None else x.astype(dtype)
    y_ref = oracle.csr_spmv(rp, ci, v, x)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    A = hs.Matrix.from_csr(rp, ci, v)
    for k_, val in (opts or {}).items():
        A.set_option(k_, val)
    A.set_x(x)
    for algo in algos:
        A.set_y(np.full(n, np.nan, dtype))   # poison: rows a kernel forgets to write are caught
        A.spmv(algo)
        y = A.get_y()
        assert_within_bound(y, y_ref, scale, dtype, "%s %s [%s]" % (what, hs.ALGO_NAMES[algo], A.last_kernel()))
    A.close()


@pytest.mark.parametrize("dtype", DT, ids=DT_IDS)
def test_csr_regular_matrices(dtype, hs, oracle, gen):
    algos = [hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE, hs.ALGO_CSR_BLOCK, hs.ALGO_AUTO]
    for name, (rp, ci, v) in [("lap64rcm", gen.laplace2d_5pt(64, dtype=dtype, rcm=True, mode="varied")),
                              ("lap33", gen.laplace2d_5pt(33, dtype=dtype)),
                              ("st27_16", gen.stencil27(16, dtype=dtype, mode="varied")),
                              ("shell", gen.banded_shell(20000, dtype=dtype, band=400))]:
        check_csr(hs, oracle, gen, rp, ci, v, algos, what=name)


@pytest.mark.parametrize("dtype", DT, ids=DT_IDS)
def test_csr_ragged_empty_and_long_rows(dtype, hs, oracle, gen):
    algos = [hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE, hs.ALGO_CSR_BLOCK, hs.ALGO_AUTO]
    check_csr(hs, oracle, gen, *gen.random_ragged(3000, max_len=70, dtype=dtype), algos, what="ragged")
    check_csr(hs, oracle, gen, *gen.random_ragged(257, max_len=257, dtype=dtype, empty_frac=0.5), algos, what="ragged-half-empty")
    check_csr(hs, oracle, gen, *gen.rmat(13, 8, dtype=dtype), algos, what="rmat13")
    # leading / trailing empty rows, one row longer than the block-per-row threshold and than a merge tile
    n = 9000
    lens = np.zeros(n, np.int64)
    lens[5] = 8999
    lens[6:40] = 3
    lens[100] = 4100
    lens[8000:8990] = 1
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rng = np.random.default_rng(1)
    ci = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens]).astype(np.int32)
    v = (rng.integers(1, 64, size=len(ci)) / 64.0).astype(dtype)
    check_csr(hs, oracle, gen, rp, ci, v, algos, what="long-rows")
    # hub rows longer than one 16 K-entry segment of the long-row kernel (3 segments, 2, exactly 1)
    n2 = 60000
    lens2 = np.full(n2, 2, np.int64)
    lens2[7], lens2[30000], lens2[59999] = 40001, 16385, 16384
    rp2 = np.concatenate([[0], np.cumsum(lens2)]).astype(np.int32)
    ci2 = np.concatenate([np.sort(rng.choice(n2, size=l, replace=False)) for l in lens2]).astype(np.int32)
    v2 = (rng.integers(1, 64, size=len(ci2)) / 64.0).astype(dtype)
    check_csr(hs, oracle, gen, rp2, ci2, v2, [hs.ALGO_CSR_VECTOR, hs.ALGO_AUTO, hs.ALGO_CSR_MERGE, hs.ALGO_CSR_BLOCK], what="hub-rows")
    for vec in (1, 2, 4, 8, 16, 32):   # forced lane counts incl. mismatched ones
        check_csr(hs, oracle, gen, rp, ci, v, [hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_BLOCK], opts={hs.OPT_VEC: vec}, what="long-rows vec=%d" % vec)
    for cap in (256, 1024, 16384):     # CSR_BLOCK: block capacities below / above the long rows
        check_csr(hs, oracle, gen, rp, ci, v, [hs.ALGO_CSR_BLOCK], opts={hs.OPT_BLOCK_NNZ: cap}, what="long-rows block cap=%d" % cap)
        check_csr(hs, oracle, gen, *gen.random_ragged(3000, max_len=70, dtype=dtype), [hs.ALGO_CSR_BLOCK], opts={hs.OPT_BLOCK_NNZ: cap},
                  what="ragged block cap=%d" % cap)
    # masked single-round shapes (HSPMV_OPT_TAIL), one row per lane group and several
    for vec, tail, rpt in ((1, 12, 32), (2, 12, 16), (4, 12, 8), (4, 14, 24), (8, 8, 4), (8, 12, 12), (16, 8, 2), (32, 4, 1)):
        check_csr(hs, oracle, gen, rp, ci, v, [hs.ALGO_CSR_STREAM],
                  opts={hs.OPT_VEC: vec, hs.OPT_CHUNK: -1, hs.OPT_TAIL: tail, hs.OPT_ROWS_PER_TILE: rpt},
                  what="long-rows masked vec=%d tail=%d rpt=%d" % (vec, tail, rpt))
    # tiny capacity: every tile takes the direct (non-staged) path; tiny merge tile: rows span many tiles
    check_csr(hs, oracle, gen, rp, ci, v, [hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE],
              opts={hs.OPT_STAGE_NNZ: 64, hs.OPT_MERGE_TILE: 64, hs.OPT_ROWS_PER_TILE: 7}, what="long-rows tiny tiles")


@pytest.mark.parametrize("dtype", DT, ids=DT_IDS)
def test_csr_degenerate_shapes(dtype, hs, oracle, gen):
    algos = [hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE, hs.ALGO_CSR_BLOCK, hs.ALGO_AUTO]
    # 1x1, single dense row, all-empty matrix, diagonal
    check_csr(hs, oracle, gen, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.5], dtype), algos, what="1x1")
    n = 300
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = n
    check_csr(hs, oracle, gen, rp, np.arange(n, dtype=np.int32), np.full(n, 0.5, dtype), algos, what="one dense row")
    check_csr(hs, oracle, gen, np.zeros(n + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, dtype), algos, what="empty matrix")
    check_csr(hs, oracle, gen, np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), np.full(n, -1.5, dtype), algos, what="diagonal")


def test_x_with_inf_outside_the_row_pattern(hs, oracle, gen):
    """Masked lanes must be selected away, not multiplied by zero."""
    rp, ci, v = gen.random_ragged(600, max_len=9, dtype=np.float64, seed=11)
    x = gen.vector_x(600)
    unused = np.setdiff1d(np.arange(600), ci)
    assert unused.size > 0
    x[unused] = np.inf
    check_csr(hs, oracle, gen, rp, ci, v, [hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE], x=x, what="inf")
    # column 0 unused and x[0] = inf: idle slots of empty rows and of short rows park on index 0;
    # every masked shape, with rows longer than one round of the shape (the excess loop) as well
    rp2, ci2, v2 = gen.random_ragged(900, max_len=70, dtype=np.float64, seed=12)
    ci2 = (ci2 + 1).astype(np.int32)
    x2 = np.concatenate([[np.inf], gen.vector_x(900)])
    y_ref = oracle.csr_spmv(rp2, ci2, v2, x2)
    assert np.all(np.isfinite(y_ref))
    scale = oracle.abs_rowsum(rp2, ci2, v2, np.where(np.isfinite(x2), x2, 0.0))
    A = hs.Matrix.from_csr(rp2, ci2, v2, n_cols=901)
    A.set_x(x2)
    for opts in ({}, {hs.OPT_VEC: 4, hs.OPT_CHUNK: -1, hs.OPT_TAIL: 8, hs.OPT_ROWS_PER_TILE: 8},
                 {hs.OPT_VEC: 1, hs.OPT_CHUNK: -1, hs.OPT_TAIL: 12, hs.OPT_ROWS_PER_TILE: 32},
                 {hs.OPT_VEC: 2, hs.OPT_CHUNK: 5, hs.OPT_ROWS_PER_TILE: 16}, {hs.OPT_VEC: 8, hs.OPT_CHUNK: 3, hs.OPT_ROWS_PER_TILE: 4}):
        for k_ in (hs.OPT_VEC, hs.OPT_CHUNK, hs.OPT_TAIL, hs.OPT_ROWS_PER_TILE):
            A.set_option(k_, 0)
        for k_, val in opts.items():
            A.set_option(k_, val)
        for algo in (hs.ALGO_CSR_STREAM, hs.ALGO_CSR_VECTOR, hs.ALGO_AUTO):
            A.set_y(np.full(900, np.nan))
            A.spmv(algo)
            assert_within_bound(A.get_y(), y_ref, scale, np.float64, "inf at x[0] %s [%s]" % (opts, A.last_kernel()))


CSRK_CASES = [
    ("laplace2d_5pt", {"g": 64, "rcm": True, "mode": "varied"}, 2, [16]),
    ("laplace2d_5pt", {"g": 64, "rcm": True}, 3, [8, 8]),
    ("stencil27", {"g": 16, "mode": "varied"}, 2, [32]),
    ("stencil27", {"g": 16, "mode": "varied"}, 3, [20, 10]),
    ("stencil27", {"g": 16}, 3, [3, 2]),            # tiny groups
    ("stencil27", {"g": 12}, 2, [100000]),           # a single super-row, far over the stage capacity
    ("banded_shell", {"n": 20000, "band": 400}, 3, [9, 12]),
    ("banded_shell", {"n": 20000, "band": 400}, 2, [96]),
    ("random_ragged", {"n": 2000, "diag": True, "max_len": 50}, 3, [6, 5]),
]


@pytest.mark.parametrize("dtype", DT, ids=DT_IDS)
@pytest.mark.parametrize("case", CSRK_CASES, ids=["%s-k%d-%s" % (c[0], c[2], "x".join(map(str, c[3]))) for c in CSRK_CASES])
def test_csrk_matches_oracle(case, dtype, hs, oracle, gen):
    g, kw, k, sizes = case
    rp, ci, v = getattr(gen, g)(dtype=dtype, **kw)
    n = len(rp) - 1
    x = gen.vector_x(n, dtype)
    o = oracle.bandk_build(rp, ci, k, sizes)
    r2, c2, v2 = oracle.bandk_reorder(o["perm"], rp, ci, v)
    maps = [o["sr_ptr"]] + ([o["ssr_ptr"]] if k == 3 else [])
    y_ref_perm = oracle.csrk_spmv(k, maps, r2, c2, v2, x[o["perm"]])
    scale = oracle.abs_rowsum(rp, ci, v, x)

    # reference-named interface, as spmv-csrk/spmv.cpp:139-187 uses it
    A = hs.CSRk_Graph(n, n, len(ci), rp, ci, v, "SpMV", "", "HAND", False, k, sizes)
    A.putInCSRkFormat()
    got = A.A.csrk()
    assert np.array_equal(A.getPermutation(), o["perm"])
    assert np.array_equal(got["sr_ptr"], o["sr_ptr"])
    if k == 3:
        assert np.array_equal(got["ssr_ptr"], o["ssr_ptr"])
    assert A.getNumCoarsestRows() == len(maps[-1]) - 1
    d_rp, d_ci, d_v = A.A.download()      # the REORDERED matrix is what lives in HBM
    assert np.array_equal(d_rp, r2.astype(np.int32)) and np.array_equal(d_ci, c2.astype(np.int32)) and np.array_equal(d_v, v2)
    A.setX(x)
    y = np.full(n, np.nan, dtype)
    A.SpMV(y)
    assert_within_bound(y, y_ref_perm, scale[o["perm"]], dtype, "CSRk_Graph.SpMV [%s]" % A.A.last_kernel())
    # the reference driver's own check: y[i] vs y_csrser[perm[i]]
    y_csr = oracle.csr_spmv(rp, ci, v, x)
    assert_within_bound(y, y_csr[o["perm"]], scale[o["perm"]], dtype, "vs serial CSR")
    assert_within_bound(A.A.get_y(original_order=True), y_csr, scale, dtype, "original order")
    # every kernel knob; plain-CSR kernels must also work on the reordered matrix
    for vec in (1, 2, 4, 8, 16, 32):
        for warps, stages in ((4, 2), (8, 3), (16, 2)):
            A.A.set_option(hs.OPT_VEC, vec)
            A.A.set_option(hs.OPT_WARPS, warps)
            A.A.set_option(hs.OPT_STAGES, stages)
            A.SpMV(y)
            assert_within_bound(y, y_ref_perm, scale[o["perm"]], dtype, "knobs [%s]" % A.A.last_kernel())
    A.A.set_option(hs.OPT_VEC, 0)
    for algo in (hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE):
        A.A.spmv(algo)
        assert_within_bound(A.A.get_y(False), y_ref_perm, scale[o["perm"]], dtype, hs.ALGO_NAMES[algo])


def _arrow(n, hub_deg, band, seed=3):
    """symmetric band + one hub row/column of hub_deg entries + diagonal (long rows for the sorters)"""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    i = np.repeat(np.arange(n), band)
    j = np.clip(i + rng.integers(1, 3 * band, len(i)), 0, n - 1)
    hub = rng.choice(np.arange(1, n), hub_deg, replace=False)
    i = np.concatenate([i, np.full(hub_deg, n // 3)])
    j = np.concatenate([j, hub])
    a = sp.coo_matrix((np.ones(len(i)), (i, j)), shape=(n, n)).tocsr()
    a = (a + a.T + sp.identity(n)).tocsr()
    a.sort_indices()
    a.data[:] = 1.0 + (np.arange(a.nnz) % 7) * 0.25
    return a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(np.float64)


@pytest.mark.parametrize("case", [
    ("stencil27", {"g": 20, "mode": "varied"}, 2, [16]),
    ("stencil27", {"g": 20, "mode": "varied"}, 3, [16, 16]),
    ("laplace2d_5pt", {"g": 150, "rcm": True}, 3, [96, 8]),
    ("shell_blocks", {"n": 30000}, 3, [12, 6]),
    ("arrow", {"n": 6000, "hub_deg": 700, "band": 3}, 3, [8, 4]),     # 700-entry row: shared-memory sorters
    ("arrow", {"n": 6000, "hub_deg": 1500, "band": 3}, 2, [8]),      # > 1024-entry row: host takes over
    ("stencil27", {"g": 12}, 2, [400]),                              # 10.8 k entries per group: host takes over
], ids=lambda c: "%s-k%d-%s" % (c[0], c[2], "x".join(map(str, c[3]))))
def test_device_bandk_build_equals_host_build(case, hs, gen):
    """hspmv_build_csrk builds the level-1 coarse graph and A' = P A P^T in HBM; every
    array must equal what the host builder (HSPMV_OPT_BANDK_HOST, itself pinned to the reference) gives."""
    g, kw, k, sizes = case
    rp, ci, v = _arrow(**kw) if g == "arrow" else getattr(gen, g)(**kw)
    out = []
    for host in (0, 1):
        A = hs.Matrix.from_csr(rp, ci, v, keep_host=bool(host))
        A.set_option(hs.OPT_BANDK_HOST, host)
        l0 = hs.launch_count()
        A.build_csrk(k, sizes)
        out.append((A.csrk(), A.download(), hs.launch_count() - l0, A.row_stats()))
        x = gen.vector_x(len(rp) - 1, np.float64)
        A.set_x(x)
        A.spmv(hs.ALGO_CSRK)
        out[-1] += (A.get_y(original_order=True),)
    (cd, md, ld, sd, yd), (ch, mh, lh, sh, yh) = out
    for key in ch:
        assert np.array_equal(cd[key], ch[key]), key
    for a, b, name in zip(md, mh, ("row_ptr", "col_idx", "val")):
        assert np.array_equal(a, b), name
    assert np.array_equal(sd, sh) and np.array_equal(yd, yh)
    handed_over = (g == "arrow" and kw["hub_deg"] > 1024) or sizes[0] >= 400
    assert (ld > lh) != handed_over, "device builder launches kernels unless it hands over (%d vs %d)" % (ld, lh)


def test_ragged_rows_pick_the_masked_shape_and_batch_timer_runs(hs, oracle, gen):
    """rows of 35 +- 10 (config-3 stand-in): AUTO must choose the masked one-round shape sized for the
    90th-percentile row; hspmv_spmv_timed_batch must time launches without changing the result"""
    rp, ci, v = gen.shell_blocks(40000)
    lens = np.diff(rp)
    x = gen.vector_x(len(rp) - 1, np.float64)
    A = hs.Matrix.from_csr(rp, ci, v)
    A.set_x(x)
    ms = A.spmv_timed_batch(hs.ALGO_AUTO, warmup=2, iters=5)
    assert 0 < ms < 50
    k = A.last_kernel()
    p90 = int(np.sort(lens)[(len(lens) * 90 + 99) // 100 - 1])
    vec = 1
    while vec < 32 and p90 / vec > 12:
        vec *= 2
    assert "vec=%d,main=0," % vec in k, (k, p90)
    assert_within_bound(A.get_y(), oracle.csr_spmv(rp, ci, v, x), oracle.abs_rowsum(rp, ci, v, x), np.float64, k)


def test_autotune_keeps_results_and_is_not_slower(hs, oracle, gen):
    """hspmv_autotune picks among compiled kernel shapes by measurement"""
    for (rp, ci, v), k, sizes in ((gen.stencil27(40), 1, None), (gen.shell_blocks(60000), 2, [16]), (gen.laplace2d_5pt(300, rcm=True), 3, [96, 8])):
        n = len(rp) - 1
        x = gen.vector_x(n, np.float64)
        A = hs.Matrix.from_csr(rp, ci, v)
        if k > 1:
            A.build_csrk(k, sizes)
        A.set_x(x)
        before = min(A.spmv_timed_batch(hs.ALGO_AUTO, warmup=3, iters=10) for _ in range(3))
        kernel0 = A.last_kernel()
        tuned = A.autotune(hs.ALGO_AUTO, effort=2)
        after = min(A.spmv_timed_batch(hs.ALGO_AUTO, warmup=3, iters=10) for _ in range(3))
        assert after <= before * 1.10 + 0.002, (kernel0, before, A.last_kernel(), after, tuned)
        A.set_y(np.full(n, np.nan))
        A.spmv(hs.ALGO_AUTO)
        assert_within_bound(A.get_y(original_order=True), oracle.csr_spmv(rp, ci, v, x), oracle.abs_rowsum(rp, ci, v, x),
                            np.float64, "autotuned [%s]" % A.last_kernel())
        A.close()


def test_csrk_against_golden_reference_vectors(hs, oracle, gen):
    """y from the reference's own OpenMP CSR-k / CSR runs, stored in tests/golden."""
    import importlib.util
    gold = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("mg", os.path.join(gold, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for path in sorted(glob.glob(os.path.join(gold, "*.npz"))):
        g = np.load(path)
        fn, kw, k, sizes = mg.CASES[os.path.basename(path)[:-4]]
        for dtype, tag in ((np.float32, "f32"), (np.float64, "f64")):
            rp, ci, v = getattr(gen, fn)(dtype=dtype, **kw)
            x = gen.vector_x(len(rp) - 1, dtype)
            scale = oracle.abs_rowsum(rp, ci, v, x)
            A = hs.Matrix.from_csr(rp, ci, v)
            A.set_x(x)
            for algo in (hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE):
                A.spmv(algo)
                assert_within_bound(A.get_y(), g["y_csr_" + tag], scale, dtype, path + " csr")
            A.build_csrk(k, sizes)
            c = A.csrk()
            assert np.array_equal(c["perm"], g["perm"]) and np.array_equal(c["sr_ptr"], g["sr_ptr"])
            if k == 3:
                assert np.array_equal(c["ssr_ptr"], g["ssr_ptr"])
            A.set_x(x)
            A.spmv(hs.ALGO_CSRK)
            assert_within_bound(A.get_y(False), g["y_csrk_" + tag], scale[g["perm"]], dtype, path + " csrk")
            A.close()


def test_csr3_file_direct_load(tmp_path, hs, oracle, gen):
    rp, ci, v = gen.stencil27(12, mode="varied")
    o = oracle.bandk_build(rp, ci, 3, [20, 10])
    r2, c2, v2 = oracle.bandk_reorder(o["perm"], rp, ci, v)
    p = str(tmp_path / "m.csr3")
    hs.write_csr3(p, o["ssr_ptr"], o["sr_ptr"], r2, c2, v2)
    d = hs.read_csr3(p)
    A = hs.Matrix.from_csrk(d["sr_ptr"], d["row_ptr"], d["col_idx"], d["val"], ssr_ptr=d["ssr_ptr"])
    assert A.info()["k"] == 3
    x = gen.vector_x(len(rp) - 1)
    xp = x[o["perm"]]
    A.set_x(xp)                     # a .csr3 carries no permutation: vectors are in file order
    A.spmv(hs.ALGO_CSRK)
    y_ref = oracle.csrk_spmv(3, [o["sr_ptr"], o["ssr_ptr"]], r2, c2, v2, xp)
    assert_within_bound(A.get_y(), y_ref, oracle.abs_rowsum(rp, ci, v, x)[o["perm"]], np.float64, "csr3")
    B = hs.Matrix.from_csrk(d["sr_ptr"], d["row_ptr"], d["col_idx"], d["val"])   # k = 2 view
    B.set_x(xp)
    B.spmv(hs.ALGO_CSRK)
    assert_within_bound(B.get_y(), y_ref, oracle.abs_rowsum(rp, ci, v, x)[o["perm"]], np.float64, "csr2")


@pytest.mark.parametrize("dtype", DT, ids=DT_IDS)
def test_device_generators_match_numpy(dtype, hs, gen):
    for kind, g, fn in ((27, 9, gen.stencil27), (27, 2, gen.stencil27), (27, 1, gen.stencil27),
                        (5, 17, gen.laplace2d_5pt), (5, 2, gen.laplace2d_5pt)):
        for mode, mname in ((0, "classic"), (1, "varied")):
            rp, ci, v = fn(g, dtype=dtype, mode=mname)
            A = hs.Matrix.stencil(kind, g, dtype, mode)
            d = A.download()
            assert np.array_equal(d[0], rp) and np.array_equal(d[1], ci) and np.array_equal(d[2], v), (kind, g, mname)
            A.close()
    # a row block with localised columns (multi-GPU shard)
    full = gen.stencil27(10, dtype=dtype, mode="varied")
    lo, hi = 333, 777
    A = hs.Matrix.stencil(27, 10, dtype, 1, row_lo=lo, row_hi=hi)
    cmin, cmax = A.col_range()
    e0, e1 = full[0][lo], full[0][hi]
    assert (cmin, cmax) == (full[1][e0:e1].min(), full[1][e0:e1].max())
    A.close()
    B = hs.Matrix.stencil(27, 10, dtype, 1, row_lo=lo, row_hi=hi, col_shift=cmin, n_cols=cmax - cmin + 1)
    d = B.download()
    assert np.array_equal(d[0], full[0][lo:hi + 1] - e0) and np.array_equal(d[1], full[1][e0:e1] - cmin)
    B.fill_x_synthetic(cmin)
    B.spmv()
    import oracle as O
    y_ref = O.csr_spmv(full[0], full[1], full[2], gen.vector_x(1000, dtype))[lo:hi]
    sc = O.abs_rowsum(full[0], full[1], full[2], gen.vector_x(1000, dtype))[lo:hi]
    assert_within_bound(B.get_y(), y_ref, sc, dtype, "row block")


@pytest.mark.parametrize("k", [1, 2, 3])
def test_slice_rows_shards_of_a_csrk_matrix(k, hs, oracle, gen):
    """a rank's shard of a Band-k-reordered matrix = rows between super-(super-)row boundaries
    (hspmv_partition_rows with align_ptr), cut out in HBM with its column window; the shards' products,
    concatenated, equal the whole matrix's"""
    rp, ci, v = gen.stencil27(22, mode="varied")
    n = len(rp) - 1
    x = gen.vector_x(n)
    A = hs.Matrix.from_csr(rp, ci, v)
    align = None
    if k > 1:
        A.build_csrk(k, [12, 5])
        c = A.csrk()
        align = c["sr_ptr"] if k == 2 else c["sr_ptr"][c["ssr_ptr"]]
    A.set_x(x)                     # permuted inside for k > 1
    A.spmv(hs.ALGO_AUTO)
    y_full = A.get_y(original_order=False)
    xp = x[A.csrk()["perm"]] if k > 1 else x
    rp_dev = A.row_ptr()
    parts = 3
    row_lo = hs.partition_rows(rp_dev, parts, align_ptr=align)
    assert row_lo[0] == 0 and row_lo[-1] == n
    if align is not None:
        assert np.all(np.isin(row_lo, align))
    got = []
    for q in range(parts):
        lo, hi = int(row_lo[q]), int(row_lo[q + 1])
        wide = A.slice_rows(lo, hi)
        cmin, cmax = wide.col_range()
        wide.close()
        B = A.slice_rows(lo, hi, cmin, cmax - cmin + 1)
        i = B.info()
        assert i["m"] == hi - lo and i["n"] == cmax - cmin + 1 and i["k"] == (k if hi > lo else 1)
        assert np.array_equal(B.row_ptr(), rp_dev[lo:hi + 1] - rp_dev[lo])
        B.set_x(xp[cmin:cmax + 1])
        B.set_y(np.full(hi - lo, np.nan))
        B.spmv(hs.ALGO_AUTO)
        assert ("k=%d" % k) in B.last_kernel()
        got.append(B.get_y(original_order=False))
        if cmax - cmin + 1 > 3:
            with pytest.raises(hs.HspmvError):
                A.slice_rows(lo, hi, cmin + 1, cmax - cmin)       # window too narrow
        B.close()
    assert np.array_equal(np.concatenate(got), y_full)
    if k > 1:
        with pytest.raises(hs.HspmvError):
            A.slice_rows(int(align[1]) - 1 if align[1] > 1 else 1, n)   # not a super-row boundary


def test_merge_tile_table_device_equals_host(hs, gen):
    rp, ci, v = gen.rmat(12, 8)
    A = hs.Matrix.from_csr(rp, ci, v)
    A.set_option(hs.OPT_MERGE_TILE, 256)
    A.set_x(gen.vector_x(len(rp) - 1))
    A.spmv(hs.ALGO_CSR_MERGE)
    tile, table = A.merge_tile_rows()
    assert tile == 256 and np.array_equal(table, hs.merge_tile_rows_host(rp, 256))


def test_end_to_end_host_call_and_spmv_h_compat(hs, oracle, gen):
    rp, ci, v = gen.stencil27(14, mode="varied")
    n = len(rp) - 1
    x = gen.vector_x(n)
    y_ref = oracle.csr_spmv(rp, ci, v, x)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    # spmv.h:11-12 signature; second call hits the upload cache
    y = np.zeros(n)
    before = hs.launch_count()
    hs.omp_spmv(n, n, len(ci), rp, ci, v, x, y)
    assert_within_bound(y, y_ref, scale, np.float64, "omp_spmv")
    y2 = np.zeros(n)
    hs.omp_spmv(n, n, len(ci), rp, ci, v, 2 * x, y2)
    assert_within_bound(y2, 2 * y_ref, 2 * scale, np.float64, "omp_spmv cached")
    assert hs.launch_count() >= before + 2
    L = hs.lib()
    assert L.hspmv_omp_spmv_cache_size() == 1
    # the arrays are refilled IN PLACE (same addresses, same sizes) -- the cache must notice and upload again
    v[:] = 2.0 * v
    y3 = np.zeros(n)
    hs.omp_spmv(n, n, len(ci), rp, ci, v, x, y3)
    assert_within_bound(y3, 2 * y_ref, 2 * scale, np.float64, "omp_spmv after the values changed in place")
    assert L.hspmv_omp_spmv_cache_size() == 1
    v[:] = 0.5 * v
    # ... and the cache is bounded: six different matrices leave at most four resident
    keep = []
    for g_ in (5, 6, 7, 8, 9, 10):
        m_ = gen.stencil27(g_, mode="varied")
        keep.append(m_)
        xx = gen.vector_x(len(m_[0]) - 1)
        yy = np.zeros(len(xx))
        hs.omp_spmv(len(xx), len(xx), len(m_[1]), m_[0], m_[1], m_[2], xx, yy)
        assert_within_bound(yy, oracle.csr_spmv(m_[0], m_[1], m_[2], xx), oracle.abs_rowsum(m_[0], m_[1], m_[2], xx), np.float64, "omp_spmv g=%d" % g_)
    assert L.hspmv_omp_spmv_cache_size() <= 4
    L.hspmv_omp_spmv_cache_clear()
    assert L.hspmv_omp_spmv_cache_size() == 0
    # CSR-k handle: host x in original order -> host y in original order
    A = hs.Matrix.from_csr(rp, ci, v)
    A.build_csrk(3, [20, 10])
    assert_within_bound(A.spmv_host(x), y_ref, scale, np.float64, "spmv_host csrk")
    ms = A.spmv_timed(hs.ALGO_CSRK, warmup=2, iters=5, flush_l2=True)
    assert ms.shape == (5,) and np.all(ms > 0)
    A.set_option(hs.OPT_L2_PERSIST_X, 1)
    assert_within_bound(A.spmv_host(x), y_ref, scale, np.float64, "with L2 window")


def _stencil27_ones_expected(g, lo=0, hi=None):
    """A * 1 for the classic 27-point stencil (26 on the diagonal, -1 elsewhere): 27 - (points of the stencil
    inside the grid) = 27 - cx*cy*cz with c = 2 on a boundary plane, 3 inside"""
    hi = g ** 3 if hi is None else hi
    idx = np.arange(lo, hi, dtype=np.int64)
    out = np.ones(hi - lo, dtype=np.int64)
    for _ in range(3):
        c = idx % g
        out *= np.where((c == 0) | (c == g - 1), 2, 3) if g > 1 else 1
        idx //= g
    return (27 - out).astype(np.float64)


@pytest.mark.parametrize("fmt", ["csr", "csr2", "csr3"])
def test_config5_full_size_closed_form(fmt, hs, gen):
    """BASELINE config 5 at FULL size (27-point 400^3: 64 M rows, 1.72 G nnz, generated in HBM): with x = 1 every
    y_i is a small integer known in closed form, so the whole 64 M-entry result is compared exactly; CSR-2 / CSR-3 go
    through the Band-k build in HBM and come back in the original row order."""
    g = 400
    n = g ** 3
    A = hs.Matrix.stencil(27, g, np.float64, 0)
    info = A.info()
    assert info["m"] == n and info["nnz"] == 1719374392 == (3 * g - 2) ** 3
    if fmt != "csr":
        k = int(fmt[-1])
        sizes, _, _ = hs.auto_params(info["nnz"], n, hs.ARCH_B200)
        A.build_csrk(k, sizes)
        c = A.csrk()
        assert np.array_equal(np.sort(c["perm"]), np.arange(n, dtype=np.uint32))            # a permutation
        assert c["sr_ptr"][0] == 0 and c["sr_ptr"][-1] == n and np.all(np.diff(c["sr_ptr"].astype(np.int64)) > 0)
        if k == 3:
            assert c["ssr_ptr"][0] == 0 and c["ssr_ptr"][-1] == len(c["sr_ptr"]) - 1
    A.set_x(np.ones(n))
    A.set_y(np.full(n, np.nan))
    A.spmv(hs.ALGO_AUTO)
    y = A.get_y(original_order=True)
    want = _stencil27_ones_expected(g)
    assert np.array_equal(y, want), "%d rows differ [%s]" % (int(np.count_nonzero(y != want)), A.last_kernel())
    assert float(y.sum()) == 27.0 * n - info["nnz"]
    A.close()


def test_linearity_and_scaling_properties_at_config2_size(hs, gen):
    """Size-independent properties at BASELINE config-2 size (27-pt 100^3, 26.5 M nnz), generated
    in HBM: A(ax+by) = aAx + bAy, classic stencil rows sum to {<=0 interior = 0}, all kernels agree."""
    g = 100
    A = hs.Matrix.stencil(27, g, np.float64, 0)
    n = g ** 3
    assert A.info()["nnz"] == 298 ** 3
    x1 = gen.vector_x(n)
    x2 = np.cos(np.arange(n) * 0.001)
    ys = {}
    for algo in (hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE):
        ys[algo] = A.spmv_host(x1, algo)
    ref = ys[hs.ALGO_CSR_VECTOR]
    for algo, y in ys.items():
        assert np.max(np.abs(y - ref)) <= 1e-12 * 27 * 26 * 2, hs.ALGO_NAMES[algo]
    y2 = A.spmv_host(x2, hs.ALGO_CSR_STREAM)
    y12 = A.spmv_host(3.0 * x1 - 0.5 * x2, hs.ALGO_CSR_STREAM)
    assert np.max(np.abs(y12 - (3.0 * ys[hs.ALGO_CSR_STREAM] - 0.5 * y2))) < 1e-9
    ones = A.spmv_host(np.ones(n), hs.ALGO_CSR_STREAM)
    i = np.arange(n)
    xx, yy, zz = i % g, (i // g) % g, i // (g * g)
    interior = (xx > 0) & (xx < g - 1) & (yy > 0) & (yy < g - 1) & (zz > 0) & (zz < g - 1)
    assert np.all(ones[interior] == 0.0) and np.all(ones[~interior] > 0)
    # spot rows against the numpy generator + oracle
    import oracle as O
    rp, ci, v = gen.stencil27(g, row_lo=500000, row_hi=500200)
    y_rows = O.csr_spmv(rp, ci, v, x1)
    assert np.max(np.abs(y_rows - ys[hs.ALGO_CSR_STREAM][500000:500200])) < 1e-10


def test_row_ranges_and_interior_rows(hs, oracle, gen):
    """hspmv_spmv_range / hspmv_interior_rows: what the multi-GPU layer launches per shard."""
    import partition_oracle as PO
    g = 14
    rp, ci, v = gen.stencil27(g, mode="varied")
    n = len(rp) - 1
    row_lo = hs.partition_rows(rp, 3)
    x = gen.vector_x(n)
    y_full = oracle.csr_spmv(rp, ci, v, x)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    for r in range(3):
        lo, hi = int(row_lo[r]), int(row_lo[r + 1])
        probe = hs.Matrix.stencil(27, g, np.float64, 1, row_lo=lo, row_hi=hi)
        cmin, cmax = probe.col_range()
        assert (cmin, cmax) == hs.block_col_range(rp, ci, lo, hi)
        B = hs.Matrix.stencil(27, g, np.float64, 1, row_lo=lo, row_hi=hi, col_shift=cmin, n_cols=cmax - cmin + 1)
        b_rp, b_ci, _ = B.download()
        a, b = B.interior_rows(lo - cmin, hi - cmin)
        assert (a, b) == PO.interior_rows(b_rp, b_ci, lo - cmin, hi - cmin)
        B.set_x(x[cmin:cmax + 1])
        for algo in (hs.ALGO_CSR_STREAM, hs.ALGO_CSR_VECTOR, hs.ALGO_AUTO):
            B.set_y(np.full(hi - lo, np.nan))
            B.spmv_range(a, b, algo)                       # interior first, as ShardedSpMV does
            y = B.get_y()
            assert np.all(np.isnan(y[:a])) and np.all(np.isnan(y[b:])) and not np.any(np.isnan(y[a:b]))
            B.spmv_range(0, a, algo)
            B.spmv_range(b, hi - lo, algo)
            assert_within_bound(B.get_y(), y_full[lo:hi], scale[lo:hi], np.float64, "shard %d %s" % (r, hs.ALGO_NAMES[algo]))
        with pytest.raises(hs.HspmvError):
            B.spmv_range(0, 5, hs.ALGO_CSR_MERGE)


def test_pipelined_host_call_equals_sequential(hs, gen):
    """hspmv_spmv_host on a large plain-CSR handle overlaps H2D(x) / kernel / D2H(y) over row
    chunks; the result must be bit-identical to the strictly sequential call, and right."""
    import torch
    import oracle as O
    for kind, g in ((27, 170), (5, 2300)):
        A = hs.Matrix.stencil(kind, g, np.float64, 1)
        n = A.info()["n"]
        assert n * 8 >= 32 << 20
        x = torch.from_numpy(gen.vector_x(n) * np.cos(np.arange(n) * 1e-3)).pin_memory()
        y1 = torch.empty(n, dtype=torch.float64).pin_memory()
        y2 = torch.empty(n, dtype=torch.float64).pin_memory()
        y1.fill_(float("nan"))
        A.spmv_host_ptr(x.data_ptr(), y1.data_ptr())                  # pipelined
        A.set_option(hs.OPT_NO_PIPELINE, 1)
        A.spmv_host_ptr(x.data_ptr(), y2.data_ptr())                  # sequential
        assert torch.equal(y1, y2), (kind, g)
        # pageable buffers take the same path without the overlap
        A.set_option(hs.OPT_NO_PIPELINE, 0)
        y3 = A.spmv_host(x.numpy().copy())
        assert np.array_equal(y3, y2.numpy())
        # spot rows against the generator + oracle (rows around a chunk boundary)
        lo = n // 4 - 50
        fn = gen.stencil27 if kind == 27 else None
        if fn is not None:
            rp, ci, v = fn(g, mode="varied", row_lo=lo, row_hi=lo + 100)
            ref = O.csr_spmv(rp, ci, v, x.numpy())
            sc = O.abs_rowsum(rp, ci, v, x.numpy())
            assert_within_bound(y1.numpy()[lo:lo + 100], ref, sc, np.float64, "pipelined rows")
        A.close()


@pytest.mark.parametrize("dtype", DT, ids=DT_IDS)
def test_tolerance_is_exercised_by_rough_values(dtype, hs, oracle, gen):
    """The generator values are dyadic, so sums are exact and every kernel matches the oracle bit
    for bit.  With arbitrary values and x the kernels' FMA / lane-tree summation order differs from
    the oracle's serial one: the difference must be non-zero somewhere and inside the bound."""
    rng = np.random.default_rng(42)
    rp, ci, _ = gen.shell_blocks(50000, dofs=5)
    n = len(rp) - 1
    v = rng.standard_normal(len(ci)).astype(dtype)
    x = rng.standard_normal(n).astype(dtype)
    y_ref = oracle.csr_spmv(rp, ci, v, x)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    A = hs.Matrix.from_csr(rp, ci, v)
    A.set_x(x)
    differs = 0
    for algo in (hs.ALGO_CSR_STREAM, hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_MERGE):
        A.spmv(algo)
        y = A.get_y()
        assert_within_bound(y, y_ref, scale, dtype, hs.ALGO_NAMES[algo])
        differs += int(np.count_nonzero(y != y_ref))
    A.build_csrk(3, hs.auto_params(len(ci), n, hs.ARCH_B200)[0])
    A.set_x(x)
    A.spmv(hs.ALGO_CSRK)
    y = A.get_y(True)
    assert_within_bound(y, y_ref, scale, dtype, "csrk")
    differs += int(np.count_nonzero(y != y_ref))
    assert differs > 0, "suspicious: rough inputs reproduced the serial sum bit for bit everywhere"
    # and an fp32 accumulation of fp64 data would NOT pass: the bound is tight enough to tell
    if dtype == np.float64:
        y32 = oracle.csr_spmv(rp, ci, v.astype(np.float32), x.astype(np.float32)).astype(np.float64)
        assert np.count_nonzero(np.abs(y32 - y_ref) > 1e-12 * scale) > n // 2


def test_fused_halo_launch_single_gpu(hs, oracle, gen):
    """hspmv_spmv_halo: interior | low | high ranges in one launch, with the in-kernel flag wait
    satisfied by a front-end flag write on the same stream (what a neighbour GPU would do)."""
    import torch
    g = 18
    rp, ci, v = gen.stencil27(g, mode="varied")
    n = len(rp) - 1
    x = gen.vector_x(n)
    y_ref = oracle.csr_spmv(rp, ci, v, x)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    A = hs.Matrix.from_csr(rp, ci, v)
    A.set_x(x)
    flags = hs.peer_alloc(4 * 64)      # ints [0,2): data flags; [8, 8 + HSPMV_HALO_DIAG_WORDS): timeout word + spin diagnostics
    st = torch.cuda.current_stream().cuda_stream
    for a, b in ((g * g + g + 1, n - g * g - g - 1), (0, n), (0, 0), (n, n), (100, 101)):
        for n_wait, step in ((0, 0), (2, 7)):
            A.set_y(np.full(n, np.nan))
            if n_wait:
                hs.stream_write(st, flags, step)            # two "neighbours" arrive
                hs.stream_write(st, flags + 4, step + 1)
            A.spmv_halo(a, b, wait_flags=flags if n_wait else None, n_wait=n_wait, wait_step=step,
                        timed_out=(flags + 32) if n_wait else None, stream=st)
            assert_within_bound(A.get_y(), y_ref, scale, np.float64, "halo launch [%d,%d) wait=%d" % (a, b, n_wait))
    from heterogeneous_spmv_b200 import dist as hd
    assert int(hd.wrap_device_ptr(flags + 32, 1, torch.int32).item()) == 0      # no timeout recorded
    hs.stream_wait_ge(st, flags, 7)                                               # already satisfied: returns
    torch.cuda.synchronize()
    hs.peer_free(flags)


def test_hub_column_relabel_keeps_the_product(hs, oracle, gen):
    """hspmv_relabel_hub_columns: columns renumbered by popularity (x permuted on the way in, y untouched)."""
    for dtype in (np.float64, np.float32):
        rp, ci, v = gen.rmat(14, 8, dtype=dtype)
        n = len(rp) - 1
        x = gen.vector_x(n, dtype)
        y_ref = oracle.csr_spmv(rp, ci, v, x)
        scale = oracle.abs_rowsum(rp, ci, v, x)
        A = hs.Matrix.from_csr(rp, ci, v)
        perm = A.relabel_hub_columns()
        cnt = np.bincount(ci, minlength=n)
        assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
        assert np.all(np.diff(cnt[perm].astype(np.int64)) <= 0)                          # most referenced first
        same = np.diff(cnt[perm].astype(np.int64)) == 0
        assert np.all(np.diff(perm.astype(np.int64))[same] > 0)                          # ties by column id
        d_rp, d_ci, d_v = A.download()
        inv = np.empty(n, np.int64)
        inv[perm] = np.arange(n)
        assert np.array_equal(d_rp, rp) and np.array_equal(d_ci, inv[ci]) and np.array_equal(d_v, v)
        A.set_x(x)
        for algo in (hs.ALGO_AUTO, hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_MERGE, hs.ALGO_CSR_BLOCK):
            A.set_y(np.full(n, np.nan, dtype))
            A.spmv(algo)
            assert_within_bound(A.get_y(), y_ref, scale, dtype, "hub relabel %s [%s]" % (hs.ALGO_NAMES[algo], A.last_kernel()))
        assert_within_bound(A.spmv_host(x), y_ref, scale, dtype, "hub relabel, host call")
        with pytest.raises(hs.HspmvError):
            A.relabel_hub_columns()
        A.close()


def test_graph_replay_equals_plain_launches(hs, oracle, gen):
    """hspmv_spmv_graph: N products as one CUDA-graph replay -- same y, launch counter counts the captured kernels, options drop the graph."""
    for maker, algo in ((lambda: gen.laplace2d_5pt(64, rcm=True, mode="varied"), hs.ALGO_AUTO), (lambda: gen.rmat(13, 8), hs.ALGO_AUTO),
                        (lambda: gen.stencil27(14, mode="varied"), hs.ALGO_CSR_BLOCK)):
        rp, ci, v = maker()
        n = len(rp) - 1
        x = gen.vector_x(n)
        A = hs.Matrix.from_csr(rp, ci, v)
        A.set_x(x)
        A.spmv(algo)
        y_plain = A.get_y()
        A.set_y(np.full(n, np.nan))
        before = hs.launch_count()
        A.spmv_graph(algo, iters=4)                         # one uncaptured product (tables, module load), capture, first replay
        first = hs.launch_count() - before
        assert np.array_equal(A.get_y(), y_plain)
        A.set_y(np.full(n, np.nan))
        before = hs.launch_count()
        A.spmv_graph(algo, iters=4)                         # replay of the cached graph
        per_replay = hs.launch_count() - before
        assert np.array_equal(A.get_y(), y_plain)
        assert per_replay >= 4 and per_replay % 4 == 0 and first == per_replay + per_replay // 4
        assert A.spmv_graph_timed(algo, iters=8, reps=2) > 0
        A.set_option(hs.OPT_VEC, 4)                         # drops the graph; the next call captures again
        A.spmv_graph(algo, iters=2)
        assert_within_bound(A.get_y(), oracle.csr_spmv(rp, ci, v, x), oracle.abs_rowsum(rp, ci, v, x), np.float64, "graph after option")
        A.close()
