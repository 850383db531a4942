"""Sparse triangular solve on CSR-k, k = 2, level-set ordering.

CPU: the oracle's literal restatement (oracle/sts_oracle.cpp) against the COMPILED REFERENCE (ctor "STS"/"LS"/k=2,
putInCSRkFormat, incomplete_choloskey, compute_b, lowerSTS through oracle/ref_shim_csrk.cpp) -- permutation, pack
bounds, L, b and x bit for bit -- and the product's O(nnz log d) level finder against both.
GPU: hspmv_build_sts / compute_b / solve against the reference: index work and b bit-exact, x within
|x - x_ref| <= 1e-10 (fp64) / 1e-4 (fp32) * max|x_ref| (the packs are summed in a different order than the reference's
serial loop, and a triangular solve carries rounding differences along its dependency chains)."""
import numpy as np
import pytest


def _cases(gen):
    return [("st27_9", gen.stencil27(9, mode="varied")), ("lap40rcm", gen.laplace2d_5pt(40, rcm=True, mode="varied")),
            ("lap17", gen.laplace2d_5pt(17)), ("shell", gen.shell_blocks(6000)), ("st27_24", gen.stencil27(24, mode="varied")),
            ("banded", gen.banded_shell(4000, band=200)), ("ragged_diag_sym", _sym(gen.random_ragged(1500, max_len=12, diag=True)))]


def _sym(m):
    import scipy.sparse as sp
    rp, ci, v = m
    n = len(rp) - 1
    A = sp.csr_matrix((np.abs(v) + 1.0, ci, rp), shape=(n, n))
    A = (A + A.T + 40.0 * sp.identity(n)).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), (np.round(A.data * 64) / 64).astype(np.float64)


def _ref_or_oracle(oracle, rp, ci, v):
    """dict(perm, pack_ptr, l_rp, l_ci, l_val, b, x): from the compiled reference when it is there, else the oracle"""
    if oracle.have_ref():
        R = oracle.RefSTS(rp, ci, v)
        a = R.arrays()
        a["x"] = R.solve()
        a["_ref"] = R
        return a
    lv = oracle.sts_levels(rp, ci)
    r2, c2, v2 = oracle.bandk_reorder(lv["perm"], rp, ci, v)
    l_rp, l_ci, l_val = oracle.sts_lower(r2, c2, v2)
    b = oracle.sts_b(l_rp, l_ci, l_val)
    return {"perm": lv["perm"], "pack_ptr": lv["pack_ptr"], "l_rp": l_rp, "l_ci": l_ci, "l_val": l_val, "b": b,
            "x": oracle.sts_solve(lv["pack_ptr"], l_rp, l_ci, l_val, b)}


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_restatement_equals_the_compiled_reference(dtype, oracle, gen):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for name, (rp, ci, v) in _cases(gen):
        v = v.astype(dtype)
        R = oracle.RefSTS(rp, ci, v)
        a = R.arrays()
        lv = oracle.sts_levels(rp, ci)
        assert np.array_equal(lv["perm"], a["perm"]) and np.array_equal(lv["pack_ptr"], a["pack_ptr"]), name
        assert np.array_equal(np.sort(a["perm"]), np.arange(len(rp) - 1, dtype=np.uint32))
        r2, c2, v2 = oracle.bandk_reorder(lv["perm"], rp, ci, v)
        l_rp, l_ci, l_val = oracle.sts_lower(r2, c2, v2)
        assert np.array_equal(l_rp, a["l_rp"]) and np.array_equal(l_ci, a["l_ci"]) and np.array_equal(l_val, a["l_val"]), name
        b = oracle.sts_b(l_rp, l_ci, l_val)
        assert np.array_equal(b, a["b"]), name
        assert np.array_equal(oracle.sts_solve(lv["pack_ptr"], l_rp, l_ci, l_val, b), R.solve()), name
        # a right-hand side that is not L * 1
        b2 = (1.0 + (np.arange(len(b)) % 13) / 16.0).astype(dtype)
        R.set_b(b2)
        assert np.array_equal(oracle.sts_solve(lv["pack_ptr"], l_rp, l_ci, l_val, b2), R.solve()), name


def test_level_sets_of_the_library_are_bit_exact(hs, oracle, gen):
    for name, (rp, ci, v) in _cases(gen):
        got = hs.sts_levels_host(rp, ci)
        want = oracle.sts_levels(rp, ci)
        assert np.array_equal(got["perm"], want["perm"]) and np.array_equal(got["pack_ptr"], want["pack_ptr"]), name
        # rows of one pack are mutually independent in the reordered matrix, and every dependency points to an earlier pack
        inv = np.empty(len(rp) - 1, np.int64)
        inv[got["perm"]] = np.arange(len(rp) - 1)
        pack_of = np.searchsorted(got["pack_ptr"], np.arange(len(rp) - 1), side="right") - 1
        rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        a, b = inv[rows], inv[ci]
        low = b < a
        assert np.all(pack_of[b[low]] < pack_of[a[low]]), name


def test_level_search_rejects_missing_diagonal_and_unsymmetric_structure(hs):
    rp = np.array([0, 1, 3, 4], np.int32)
    with pytest.raises(hs.HspmvError):
        hs.sts_levels_host(rp, np.array([0, 0, 2, 2], np.int32))          # row 1 has no diagonal
    with pytest.raises(hs.HspmvError):                                      # row 2 depends on row 1, which never names row 2
        hs.sts_levels_host(np.array([0, 1, 2, 4], np.int32), np.array([0, 1, 1, 2], np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_sts_on_gpu_against_the_reference(dtype, hs, oracle, gen):
    tol = 1e-10 if dtype == np.float64 else 1e-4
    for name, (rp, ci, v) in _cases(gen):
        v = v.astype(dtype)
        n = len(rp) - 1
        want = _ref_or_oracle(oracle, rp, ci, v)
        # the reference-named interface
        G = hs.CSRk_Graph(n, n, len(ci), rp, ci, v, "STS", "LS", "HAND", False, 2, [0])
        G.putInCSRkFormat()
        G.incomplete_choloskey()
        got = G.A.sts()
        assert np.array_equal(got["perm"], want["perm"]) and np.array_equal(got["pack_ptr"], want["pack_ptr"]), name
        assert np.array_equal(G.getPermutation(), want["perm"])
        d_rp, d_ci, d_v = G.A.download()                   # the REORDERED matrix lives in HBM
        l_rp, l_ci, l_val = oracle.sts_lower(d_rp, d_ci, d_v)
        assert np.array_equal(l_rp, want["l_rp"]) and np.array_equal(l_ci, want["l_ci"]) and np.array_equal(l_val, want["l_val"]), name
        G.compute_b()
        assert np.array_equal(G.A.sts_get_b(), want["b"]), name
        for barriers in (2, 1, 0):          # the barrier-free solve, the level-scheduled one, the library's choice
            G.A.set_option(hs.OPT_STS_BARRIERS, barriers)
            for _ in range(2):              # twice: the ready flags carry an epoch
                G.lowerSTS()
            x = G.A.sts_get_x()
            assert barriers == 0 or ("sts_flags" if barriers == 2 else "sts_solve") in G.A.last_kernel()
            assert not G.A.sts_status()["timed_out"]
            assert np.all(np.isfinite(x)), name
            assert np.max(np.abs(x.astype(np.float64) - want["x"])) <= tol * np.max(np.abs(want["x"])), "%s [%s]" % (name, G.A.last_kernel())
            assert abs(G.checkError()) <= tol * n
        G.A.set_option(hs.OPT_STS_BARRIERS, 0)
        # another right-hand side, given in the ORIGINAL numbering; x back in the original numbering
        b2 = (1.0 + (np.arange(n) % 13) / 16.0).astype(dtype)
        G.A.sts_set_b(b2, original_order=True)
        G.A.sts_solve()
        x2 = G.A.sts_get_x(original_order=True)
        x2_ref = np.empty(n, dtype)
        x2_ref[want["perm"]] = oracle.sts_solve(want["pack_ptr"], want["l_rp"], want["l_ci"], want["l_val"], b2[want["perm"]])
        assert np.max(np.abs(x2.astype(np.float64) - x2_ref)) <= tol * np.max(np.abs(x2_ref)), name
        # residual of the solve itself: L x = b
        xk = G.A.sts_get_x()
        Lx = np.add.reduceat(want["l_val"].astype(np.float64) * xk[want["l_ci"]], want["l_rp"][:-1].astype(np.int64))
        assert np.max(np.abs(Lx - b2[want["perm"]])) <= 1e3 * tol * np.max(np.abs(b2)), name
        ms = G.A.sts_solve_timed(1, 3)
        assert ms.shape == (3,) and np.all(ms > 0)
        # the handle still multiplies: y = A x in the original numbering
        xx = gen.vector_x(n, dtype)
        y = G.A.spmv_host(xx)
        y_ref = oracle.csr_spmv(rp, ci, v, xx)
        assert np.max(np.abs(y - y_ref)) <= (1e-12 if dtype == np.float64 else 1e-5) * np.max(oracle.abs_rowsum(rp, ci, v, xx)), name
        G.A.close()


def test_direct_level_computation_equals_the_literal_search(hs, oracle, monkeypatch):
    """csrc/sts.cpp has two finders: the level-by-level search (the reference's decisions, literally) and, for
    structurally symmetric matrices, levels straight from 'one more than the deepest lower neighbour' plus the order in
    which the previous level's scan first meets each vertex.  Random symmetric matrices (hubs, disconnected pieces, long
    chains): both finders == oracle; matrices with extra one-sided UPPER entries (accepted by the reference: only lower
    entries need a mirror) must take the literal path and still match."""
    import scipy.sparse as sp
    rng = np.random.default_rng(99)
    for trial in range(80):
        n = int(rng.integers(2, 500))
        e = int(n * float(rng.choice([0.3, 1.0, 3.0, 8.0])))
        a = sp.coo_matrix((np.ones(e), (rng.integers(0, n, e), rng.integers(0, n, e))), shape=(n, n)).tocsr()
        if trial % 6 == 0:                                   # a hub
            spokes = rng.choice(n, size=max(1, n // 3), replace=False)
            a = a + sp.coo_matrix((np.ones(len(spokes)), (np.full(len(spokes), int(rng.integers(0, n))), spokes)), shape=(n, n)).tocsr()
        if trial % 5 == 0:                                   # a chain: as many levels as vertices
            a = a + sp.diags([np.ones(n - 1)], [1], shape=(n, n)).tocsr()
        a = (a + a.T + sp.identity(n)).tocsr()
        if trial % 4 == 3:                                   # one-sided upper entries: (g, v) with g < v and no (v, g)
            g = rng.integers(0, n, 5)
            v = rng.integers(0, n, 5)
            keep = g < v
            extra = sp.coo_matrix((np.ones(int(keep.sum())), (g[keep], v[keep])), shape=(n, n)).tocsr()
            a = (a + extra).tocsr()
        a.sort_indices()
        rp, ci = a.indptr.astype(np.int32), a.indices.astype(np.int32)
        want = oracle.sts_levels(rp, ci)
        monkeypatch.delenv("HSPMV_STS_LEVELS", raising=False)
        got = hs.sts_levels_host(rp, ci)
        monkeypatch.setenv("HSPMV_STS_LEVELS", "literal")
        lit = hs.sts_levels_host(rp, ci)
        for key in ("perm", "pack_ptr"):
            assert np.array_equal(got[key], want[key]), (trial, key, "direct")
            assert np.array_equal(lit[key], want[key]), (trial, key, "literal")


def test_sts_golden_vectors_from_the_compiled_reference(hs, oracle, gen):
    """tests/golden/sts/*.npz were produced by the compiled reference's STS path (tests/golden/make_golden_sts.py): level
    permutation, pack bounds, L, b and the solution for a right-hand side that is not L * 1.  The oracle restatement
    reproduces all of it bit for bit, and the product's level finder (both forms) the permutation and the packs --
    without /root/reference or oracle/_ref."""
    import glob
    import importlib.util
    import os
    import zlib
    gold = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("mgs", os.path.join(gold, "make_golden_sts.py"))
    mgs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgs)
    paths = sorted(glob.glob(os.path.join(gold, "sts", "*.npz")))
    assert len(paths) == len(mgs.CASES) >= 4

    def crc(a):
        return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))

    for path in paths:
        g = np.load(path)
        fn, kw = mgs.CASES[os.path.basename(path)[:-4]]
        for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
            rp, ci, v = getattr(gen, fn)(dtype=dt, **kw)
            assert crc(rp) == g["crc_in_row_ptr"] and crc(ci) == g["crc_in_col"], "generator drifted from the golden inputs"
            lv = oracle.sts_levels(rp, ci)
            assert np.array_equal(lv["perm"], g["perm"]) and np.array_equal(lv["pack_ptr"], g["pack_ptr"]), path
            got = hs.sts_levels_host(rp, ci)
            assert np.array_equal(got["perm"], g["perm"]) and np.array_equal(got["pack_ptr"], g["pack_ptr"]), path
            r2, c2, v2 = oracle.bandk_reorder(lv["perm"], rp, ci, v)
            l_rp, l_ci, l_val = oracle.sts_lower(r2, c2, v2)
            assert crc(l_rp) == g["crc_l_rp"] and crc(l_ci) == g["crc_l_ci"] and crc(l_val) == g["crc_l_val_" + tag], (path, tag)
            b = oracle.sts_b(l_rp, l_ci, l_val)
            assert crc(b) == g["crc_b_" + tag], (path, tag)
            assert bool(np.all(oracle.sts_solve(lv["pack_ptr"], l_rp, l_ci, l_val, b) == 1)) == bool(g["x_ones_is_exact_" + tag])
            x = oracle.sts_solve(lv["pack_ptr"], l_rp, l_ci, l_val, mgs.rhs(len(rp) - 1, dt))
            assert np.array_equal(x, g["x_" + tag]), (path, tag)
