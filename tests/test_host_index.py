"""CPU tests of the product's host side (no device needed): Band-k builder, reorderA, text I/O,
heuristics, partition / bins / merge tiles -- bit-exact against the oracle."""
import glob
import os
import re
import subprocess

import numpy as np
import pytest

import partition_oracle as PO


def test_library_exports_every_declared_symbol(hs):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "hspmv.h")).read()
    declared = set(re.findall(r"\b(hspmv_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("hspmv_matrix")
    assert len(declared) >= 35
    out = subprocess.check_output(["nm", "-D", "--defined-only", hs.LIB_PATH]).decode()
    exported = set(re.findall(r" T (hspmv_[a-z0-9_]+)", out))
    assert declared <= exported, "declared but not exported: %s" % sorted(declared - exported)
    L = hs.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.hspmv_version() >= 100


def test_no_cpu_fallback_without_device(hs, gen):
    """On a box without a usable sm_100 device the compute entry points fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rp, ci, v = gen.laplace2d_5pt(8)
    with pytest.raises(hs.HspmvError):
        hs.Matrix.from_csr(rp, ci, v)


CASES = [
    ("laplace2d_5pt", {"g": 30}, 3, [7, 8]),
    ("laplace2d_5pt", {"g": 50, "rcm": True}, 2, [16]),
    ("laplace2d_5pt", {"g": 50, "rcm": True}, 3, [8, 8]),
    ("stencil27", {"g": 14}, 2, [32]),
    ("stencil27", {"g": 14, "mode": "varied"}, 3, [20, 10]),
    ("stencil27", {"g": 10}, 3, [1, 1]),
    ("stencil27", {"g": 10}, 2, [100000]),     # one super-row
    ("banded_shell", {"n": 6000, "band": 150}, 3, [9, 12]),
    ("banded_shell", {"n": 6000, "band": 150}, 2, [96]),
    ("random_ragged", {"n": 500, "diag": True}, 3, [5, 4]),
]


# The RCM of the coarse graph has two walks in the product (csrc/bandk.cpp): the serial queue walk and the level-
# synchronous one that large graphs get.  HSPMV_RCM / HSPMV_RCM_CUTOFF force either on any graph, so every Band-k test
# below runs as "auto" (serial at these sizes), with the level-synchronous walk forking on every level (cutoff 1), and
# with it forking only on levels of >= 7 vertices (mixed serial / forked levels inside one search).
RCM_MODES = [None, ("parallel", 1), ("parallel", 7), "relabel"]


@pytest.fixture(params=RCM_MODES, ids=["rcm-auto", "rcm-levelsync-cut1", "rcm-levelsync-cut7", "relabel-then-coarsen"])
def rcm_mode(request, monkeypatch):
    """... and the k = 3 builder has two forms of "coarsen the renumbered level-1 graph": the renumbering applied on the
    fly (default) and the literal two steps, relabel() then coarsen (HSPMV_BANDK_RELABEL=1)."""
    if request.param == "relabel":
        monkeypatch.setenv("HSPMV_BANDK_RELABEL", "1")
    elif request.param is not None:
        monkeypatch.setenv("HSPMV_RCM", request.param[0])
        monkeypatch.setenv("HSPMV_RCM_CUTOFF", str(request.param[1]))
    return request.param


@pytest.mark.parametrize("case", CASES, ids=["%s-k%d-%s" % (c[0], c[2], "x".join(map(str, c[3]))) for c in CASES])
def test_bandk_host_bit_exact(case, hs, oracle, gen, rcm_mode):
    g, kw, k, sizes = case
    rp, ci, v = getattr(gen, g)(**kw)
    a = hs.bandk_host(rp, ci, k, sizes)
    b = oracle.bandk_build(rp, ci, k, sizes)
    for key in b:
        assert np.array_equal(a[key], b[key]), key
    assert sorted(a["perm"].tolist()) == list(range(len(rp) - 1))
    assert a["sr_ptr"][0] == 0 and a["sr_ptr"][-1] == len(rp) - 1
    for dt in (np.float32, np.float64):
        r1 = hs.reorder_host(a["perm"], rp, ci, v.astype(dt))
        r2 = oracle.bandk_reorder(b["perm"], rp, ci, v.astype(dt))
        for p, q in zip(r1, r2):
            assert np.array_equal(p, q)


def test_bandk_random_symmetric_graphs_against_oracle_and_reference(hs, oracle, rcm_mode):
    """60 random structurally symmetric matrices with a diagonal (disconnected pieces, hubs, many equal
    multiplicities -> RCM tie order): product builder == oracle restatement == the compiled reference."""
    import scipy.sparse as sp
    rng = np.random.default_rng(2024)
    for trial in range(60):
        n = int(rng.integers(2, 400))
        density = float(rng.choice([0.002, 0.01, 0.05, 0.2]))
        a = sp.random(n, n, density=density, random_state=int(rng.integers(1 << 30)), format="csr")
        if trial % 5 == 0:                                   # a hub row/column
            h = int(rng.integers(0, n))
            a = a.tolil()
            a[h, rng.choice(n, size=max(1, n // 3), replace=False)] = 1.0
            a = a.tocsr()
        if trial % 7 == 0:                                   # two blocks that never touch
            half = n // 2
            a = sp.block_diag([a[:half, :half], a[half:, half:]], format="csr")
        a = (a + a.T + sp.identity(n)).tocsr()
        a.sort_indices()
        rp, ci = a.indptr.astype(np.int32), a.indices.astype(np.int32)
        k = int(rng.integers(2, 4))
        sizes = [int(rng.integers(1, 12)), int(rng.integers(1, 8))][: k - 1]
        got = hs.bandk_host(rp, ci, k, sizes)
        want = oracle.bandk_build(rp, ci, k, sizes)
        for key in want:
            assert np.array_equal(got[key], want[key]), (trial, key, n, k, sizes)
        if oracle.have_ref() and trial % 4 == 0 and rcm_mode != RCM_MODES[2]:
            R = oracle.RefCSRk(rp, ci, np.ones(len(ci), np.float32), k, sizes)
            maps = R.maps()
            perm = R.matrix()[0]
            R.close()
            assert np.array_equal(perm, got["perm"]) and np.array_equal(maps[0], got["sr_ptr"]), (trial, "reference")
            if k == 3:
                assert np.array_equal(maps[1], got["ssr_ptr"]), (trial, "reference ssr")


def test_bandk_golden_vectors(hs, gen, rcm_mode):
    import importlib.util
    gold = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("mg", os.path.join(gold, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    for path in sorted(glob.glob(os.path.join(gold, "*.npz"))):
        g = np.load(path)
        fn, kw, k, sizes = mg.CASES[os.path.basename(path)[:-4]]
        rp, ci, v = getattr(gen, fn)(**kw)
        a = hs.bandk_host(rp, ci, k, sizes)
        assert np.array_equal(a["perm"], g["perm"]) and np.array_equal(a["sr_ptr"], g["sr_ptr"]), path
        if k == 3:
            assert np.array_equal(a["ssr_ptr"], g["ssr_ptr"]), path


def test_level_synchronous_rcm_equals_serial_walk_on_larger_graphs(hs, gen, monkeypatch):
    """The two RCM walks on graphs whose levels are wide enough to be cut into several chunks, with odd thread counts,
    on a graph of many components and on one with a hub: permutation and maps identical."""
    import ctypes as C
    import scipy.sparse as sp
    rng = np.random.default_rng(7)
    n = 30000
    e = int(0.6 * n)                                                         # near the percolation threshold: pieces of every size
    a = sp.coo_matrix((np.ones(e), (rng.integers(0, n, e), rng.integers(0, n, e))), shape=(n, n)).tocsr()
    many_pieces = (a + a.T + sp.identity(n)).tocsr()                         # thousands of small components + a giant one
    spokes = rng.choice(n, size=n // 4, replace=False)
    star = sp.coo_matrix((np.ones(len(spokes)), (np.full(len(spokes), 5), spokes)), shape=(n, n)).tocsr()
    hub = (many_pieces + star + star.T).tocsr()
    graphs = [gen.stencil27(36)[:2], gen.laplace2d_5pt(300, rcm=True)[:2]]
    for m in (many_pieces, hub):
        m.sort_indices()
        graphs.append((m.indptr.astype(np.int32), m.indices.astype(np.int32)))
    gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
    before = gomp.omp_get_max_threads()
    try:
        for gi, (rp, ci) in enumerate(graphs):
            for k, sizes in ((2, [3]), (3, [2, 3])):
                monkeypatch.setenv("HSPMV_RCM", "serial")
                want = hs.bandk_host(rp, ci, k, sizes)
                for threads, cutoff in ((3, 1), (5, 64), (before, 300)):
                    gomp.omp_set_num_threads(threads)
                    monkeypatch.setenv("HSPMV_RCM", "parallel")
                    monkeypatch.setenv("HSPMV_RCM_CUTOFF", str(cutoff))
                    got = hs.bandk_host(rp, ci, k, sizes)
                    for key in want:
                        assert np.array_equal(got[key], want[key]), (gi, k, threads, cutoff, key)
    finally:
        gomp.omp_set_num_threads(before)


def test_bandk_rejects_directed_graph_without_upper_entries(hs):
    # a "star" stored one-sided: every row references column 0 only, so every coarse vertex but
    # the first has no upper/diagonal entry -- the reference would read out of bounds
    n = 64
    rp = np.arange(n + 1, dtype=np.int32)
    ci = np.zeros(n, dtype=np.int32)
    with pytest.raises(hs.HspmvError):
        hs.bandk_host(rp, ci, 2, [4])
    import oracle as O
    with pytest.raises(RuntimeError):
        O.bandk_build(rp, ci, 2, [4])


def test_heuristics_reproduce_reference_formulas(hs):
    # 30x30 Laplacian, Volta formula -> {7, 8} (reformat-auto prints "78")
    sizes, vec, bdy = hs.auto_params(4380, 900, hs.ARCH_VOLTA)
    assert sizes == [7, 8] and vec == 0 and bdy == 12
    # helpers/params.txt style spot checks of the Ampere bands (spmv-auto-ampere.cu:130-158)
    import math
    for nnz, n in [(4996000, 1000000), (26463592, 1000000), (52500000, 1500000), (10, 1), (5000, 50)]:
        d = nnz / n
        ssrs = math.floor(0.5 + (9.175 - 1.32 * math.log(d)))
        srs = math.floor(0.5 + (20.5 - 3.5 * math.log(d)))
        vec = 0
        if 8 < d <= 16:
            vec, srs = 4, ssrs * 4
        elif 16 < d <= 32:
            vec = 8
            ssrs = math.floor(ssrs * 2.5 + 0.5)
            srs = ssrs * 3
        elif 32 < d <= 64:
            vec = 16
            ssrs *= 2
            srs = ssrs * 2
        elif d > 64:
            vec = 32
            ssrs = math.floor(ssrs * 2.7 + 0.5)
            srs = math.floor(ssrs / 4 + 0.5)
        got, gvec, _ = hs.auto_params(nnz, n, hs.ARCH_AMPERE)
        assert got == [max(ssrs, 1), max(srs, 1)] and gvec == vec
    sizes, vec, _ = hs.auto_params(26463592, 1000000, hs.ARCH_B200)
    assert vec == 2 and sizes == [16, 16]          # 27-pt stencil: 16 rows x 27 = 432 nnz per super-row


def test_text_io_round_trip(tmp_path, hs, gen, oracle):
    rp, ci, v = gen.stencil27(7, mode="varied")
    for base in (0, 1):
        for dt in (np.float32, np.float64):
            p = str(tmp_path / ("m%d.mtx.rcm.csr" % base))
            hs.my_write_csr(p, rp, ci, v.astype(dt), index_base=base)
            # our writer is readable by the reference-grammar reader restated in the oracle
            m, n, o_rp, o_ci, o_v = oracle.read_csr_text(p, base)
            assert np.array_equal(o_rp, rp) and np.array_equal(o_ci, ci) and np.array_equal(o_v, v)
            m2, n2, nnz2, r, c, w = hs.my_read_csr(p, index_base=base, dtype=dt)
            assert (m2, n2, nnz2) == (343, 343, len(ci)) and w.dtype == dt
            assert np.array_equal(r, rp) and np.array_equal(c, ci) and np.array_equal(w, v.astype(dt))
    # generator's python writer (helpers/converter.m layout) is read identically
    p = str(tmp_path / "py.mtx.csr")
    gen.write_csr_text(p, rp, ci, v, base=0)
    _, _, _, r, c, w = hs.my_read_csr(p, 0, np.float64)
    assert np.array_equal(r, rp) and np.array_equal(c, ci) and np.array_equal(w, v)
    # exotic number spellings fall back to strtod
    with open(str(tmp_path / "e.csr"), "w") as f:
        f.write("2 2 3\n0 2 3 \n0 1 1 \n1e-3 -2.5E+2 0.1234567890123456789 \n")
    _, _, _, r, c, w = hs.my_read_csr(str(tmp_path / "e.csr"), 0, np.float64)
    assert w.tolist() == [1e-3, -250.0, 0.1234567890123456789]
    with pytest.raises(hs.HspmvError):
        hs.my_read_csr(str(tmp_path / "missing.csr"))
    with open(str(tmp_path / "short.csr"), "w") as f:
        f.write("2 2 3\n0 2 3 \n0 1 \n")
    with pytest.raises(hs.HspmvError):
        hs.my_read_csr(str(tmp_path / "short.csr"))


def test_parallel_parser_and_writer_on_a_multi_megabyte_file(tmp_path, hs, gen):
    """the body is cut into one piece per host thread; every piece must land in the
    right array slots, and the parallel writer must produce the bytes of fprintf("%d "/"%f ")."""
    rp, ci, v = gen.stencil27(48, mode="varied")          # 110 k rows, 2.8 M nnz, ~40 MB of text
    v = v + np.arange(len(v)) % 1000 * 1e-3                # non-dyadic decimals
    p = str(tmp_path / "big.mtx.rcm.csr")
    hs.my_write_csr(p, rp, ci, v, index_base=1)
    assert os.path.getsize(p) > 8 << 20
    with open(p) as f:
        head = f.readline().split()
        l_rp, l_ci, l_v = f.readline(), f.readline(), f.readline()
    assert head == [str(len(rp) - 1), str(len(rp) - 1), str(len(ci))]
    assert l_rp == " ".join(str(t + 1) for t in rp.tolist()) + " \n"
    assert l_ci[:4000] == (" ".join(str(t + 1) for t in ci[:2000].tolist()) + " ")[:4000]
    assert l_v.split() == ["%f" % t for t in v.tolist()]
    m, n, nnz, r, c, w = hs.my_read_csr(p, index_base=1, dtype=np.float64)
    assert np.array_equal(r, rp) and np.array_equal(c, ci)
    assert np.array_equal(w, np.array([float("%f" % t) for t in v.tolist()]))
    # a token that is not a number, in the middle of some thread's piece
    data = open(p, "rb").read()
    mid = data.index(b" ", len(data) // 2)
    with open(str(tmp_path / "bad.csr"), "wb") as f:
        f.write(data[:mid] + b" 12x" + data[mid:])
    with pytest.raises(hs.HspmvError):
        hs.my_read_csr(str(tmp_path / "bad.csr"), index_base=1)
    with open(str(tmp_path / "cut.csr"), "wb") as f:
        f.write(data[: len(data) - 4096])
    with pytest.raises(hs.HspmvError):
        hs.my_read_csr(str(tmp_path / "cut.csr"), index_base=1)


def test_reader_accepts_any_whitespace_layout(tmp_path, hs, gen):
    """the reference readers are fscanf("%d ") loops: any mix of blanks, tabs, CR/LF and blank lines between
    tokens is the same file; tokens beyond the announced counts are ignored"""
    rng = np.random.default_rng(99)
    seps = [" ", "  ", "\t", "\n", "\r\n", " \n ", "\n\n", "\t \t"]
    for trial in range(25):
        rp, ci, v = gen.random_ragged(int(rng.integers(1, 60)), max_len=int(rng.integers(1, 12)), seed=int(rng.integers(1 << 30)))
        n = len(rp) - 1
        v = np.round(v * rng.choice([1.0, 1e3, 1e-3]), 6)
        base = int(rng.integers(0, 2))
        toks = [str(n), str(n), str(len(ci))] + [str(t + base) for t in rp] + [str(t + base) for t in ci]
        toks += [rng.choice(["%f", "%.6f", "%g", "%.3e", "%r"]) % float(t) for t in v]
        text = rng.choice(["", "\n", "  "]) + "".join(t + rng.choice(seps) for t in toks)
        if trial % 3 == 0:
            text += " 7 8 9 trailing junk is never reached"
        p = str(tmp_path / ("w%d.csr" % trial))
        with open(p, "w", newline="") as f:
            f.write(text)
        m2, n2, nnz2, r, c, w = hs.my_read_csr(p, index_base=base)
        assert (m2, n2, nnz2) == (n, n, len(ci)) and np.array_equal(r, rp) and np.array_equal(c, ci)
        want = np.array([float(t) for t in toks[3 + n + 1 + len(ci):]])
        assert np.array_equal(w, want), trial


@pytest.mark.parametrize("parts", [1, 2, 5, 16])
def test_partition_rule_on_random_row_pointers(parts, hs):
    rng = np.random.default_rng(parts)
    for _ in range(30):
        m = int(rng.integers(1, 400))
        lens = rng.integers(0, 30, m) * (rng.random(m) < 0.8)
        if rng.random() < 0.3:
            lens[int(rng.integers(0, m))] = 5000                       # one dominant row
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        got = hs.partition_rows(rp, parts)
        assert np.array_equal(got, PO.partition_rows(rp, parts))
        assert got[0] == 0 and got[-1] == m and np.all(np.diff(got) >= 0)
        step = int(rng.integers(1, 9))
        align = np.unique(np.concatenate([np.arange(0, m, step), [m]])).astype(np.uint32)
        got = hs.partition_rows(rp, parts, align_ptr=align)
        assert np.array_equal(got, PO.partition_rows(rp, parts, align)) and np.all(np.isin(got, align))


def test_halo_plan_on_random_partitions():
    """dist.halo_plan against the independent restatement, plus its two global properties: what r receives
    from q is exactly what q sends to r, and the received pieces tile the part of r's window it does not own"""
    import importlib
    hd = importlib.import_module("heterogeneous_spmv_b200.dist")
    rng = np.random.default_rng(4)
    for _ in range(200):
        P = int(rng.integers(1, 9))
        n = int(rng.integers(P, 500))
        cuts = np.sort(rng.choice(np.arange(1, n), size=P - 1, replace=False)) if P > 1 else np.zeros(0, np.int64)
        row_lo = [0] + [int(c) for c in cuts] + [n]
        ranges = []
        for r in range(P):
            lo, hi = row_lo[r], row_lo[r + 1]
            reach = int(rng.integers(0, n))                           # band, or everything (power-law)
            ranges.append((max(0, lo - reach), min(n - 1, hi - 1 + reach)))
        plans = [hd.halo_plan(row_lo, ranges, r) for r in range(P)]
        for r in range(P):
            assert plans[r] == PO.halo_plan(row_lo, ranges, r)
            recv, send = plans[r]
            for q, a, b in recv:
                assert (r, a, b) in plans[q][1]                       # q sends exactly that piece to r
            got = sorted((a, b) for _, a, b in recv)
            cmin, cmax = ranges[r]
            want = [(a, b) for a, b in ((cmin, row_lo[r]), (row_lo[r + 1], cmax + 1)) if a < b]
            flat = [i for a, b in got for i in range(a, b)]
            assert flat == [i for a, b in want for i in range(a, b)]


def test_binary_side_file_cache(tmp_path, hs, gen):
    """second and later reads of a text matrix come from <file>.hspmvbin; a changed text
    file, another index base or another dtype invalidate it"""
    rp, ci, v = gen.stencil27(9, mode="varied")
    p = str(tmp_path / "c.mtx.rcm.csr")
    hs.my_write_csr(p, rp, ci, v, index_base=1)
    a = hs.my_read_csr(p, index_base=1, cache=True)
    assert a[-1] is False and os.path.exists(p + ".hspmvbin")
    b = hs.my_read_csr(p, index_base=1, cache=True)
    assert b[-1] is True
    for u, w in zip(a[:6], b[:6]):
        assert np.array_equal(u, w)
    assert np.array_equal(b[3], rp) and np.array_equal(b[4], ci) and np.array_equal(b[5], v)
    assert hs.my_read_csr(p, index_base=1, dtype=np.float32, cache=True)[-1] is False      # other dtype: re-parsed, rewritten
    assert hs.my_read_csr(p, index_base=1, dtype=np.float32, cache=True)[-1] is True
    v2 = v.copy()
    v2[0] += 1.0
    hs.my_write_csr(p, rp, ci, v2, index_base=1)                                            # text changed (mtime/size)
    os.utime(p, ns=(os.stat(p).st_atime_ns, os.stat(p).st_mtime_ns + 5_000_000))
    c = hs.my_read_csr(p, index_base=1, dtype=np.float32, cache=True)
    assert c[-1] is False and c[5][0] == np.float32(v2[0])
    with open(p + ".hspmvbin", "r+b") as f:                                                 # damaged side file
        f.truncate(100)
    d = hs.my_read_csr(p, index_base=1, dtype=np.float32, cache=True)
    assert d[-1] is False and np.array_equal(d[4], ci)
    with pytest.raises(hs.HspmvError):
        hs.my_read_csr(str(tmp_path / "missing.csr"), cache=True)


def test_fixed_point_formatter_matches_printf(tmp_path, hs):
    """put_real6 (csrc/io.cpp) must print what fprintf("%f ") prints, ties and huge values included."""
    rng = np.random.default_rng(5)
    k = rng.integers(-10**9, 10**9, 20000).astype(np.float64)
    vals = np.concatenate([
        k * 5e-7,                                   # decimal ties x.xxxxxx5 (inexact in binary: either side)
        k * 5e-7 * (1 + 2.0 ** -52), k * 5e-7 * (1 - 2.0 ** -52),
        np.array([0.0, -0.0, 1e-7, -1e-7, 4.9999999e-7, 5e-7, 5.0000001e-7, 1.5e-6, 2.5e-6, 0.9999995, 0.99999949,
                  123456.7890125, 3.9e9, 4.1e9, 1e15, -1e15, 1e22, 2.0 ** 52 + 1, 1 / 3, 2 / 3, 1e-300]),
        rng.standard_normal(20000) * 10.0 ** rng.integers(-8, 9, 20000),
    ])
    n = len(vals)
    rp = np.array([0, n], dtype=np.int32)
    ci = np.zeros(n, dtype=np.int32)
    p = str(tmp_path / "fmt.csr")
    hs.my_write_csr(p, rp, ci, vals, index_base=0)
    toks = open(p).read().split("\n")[3].split()
    want = ["%f" % t for t in vals.tolist()]
    assert toks == want


def test_csr3_round_trip_and_layout(tmp_path, hs, gen):
    rp, ci, v = gen.laplace2d_5pt(30, dtype=np.float32)
    o = hs.bandk_host(rp, ci, 3, [7, 8])
    r2, c2, v2 = hs.reorder_host(o["perm"], rp, ci, v)
    p = str(tmp_path / "lap30.csr3")
    hs.write_csr3(p, o["ssr_ptr"], o["sr_ptr"], r2, c2, v2)
    text = open(p).read()
    # header of SURVEY Appendix B: "16 124 900 900 4380"
    assert text.split("\n")[0].split() == ["16", "124", "900", "900", "4380"]
    body = text.split("\n", 1)[1].split()
    assert body[:17] == [str(t) for t in o["ssr_ptr"]]
    assert body[-1] == "%.6f" % v2[-1]
    d = hs.read_csr3(p, np.float32)
    assert np.array_equal(d["ssr_ptr"], o["ssr_ptr"]) and np.array_equal(d["sr_ptr"], o["sr_ptr"])
    assert np.array_equal(d["row_ptr"], r2) and np.array_equal(d["col_idx"], c2) and np.array_equal(d["val"], v2)


@pytest.mark.parametrize("parts", [1, 2, 3, 8])
def test_partition_bit_exact(parts, hs, gen):
    for rp, ci, v in (gen.stencil27(12), gen.random_ragged(3000, max_len=60), gen.laplace2d_5pt(40, rcm=True)):
        a = hs.partition_rows(rp, parts)
        b = PO.partition_rows(rp, parts)
        assert np.array_equal(a, b)
        assert a[0] == 0 and a[-1] == len(rp) - 1 and np.all(np.diff(a) >= 0)
        nnz = rp[a[1:]] - rp[a[:-1]]
        assert nnz.sum() == rp[-1]
        if parts > 1 and rp[-1] > 0:
            assert nnz.max() - nnz.min() <= 2 * int(np.diff(rp).max()) + 1
        for p in range(parts):
            assert hs.block_col_range(rp, ci, a[p], a[p + 1]) == PO.col_range(rp, ci, a[p], a[p + 1])
    # snapping to super-row boundaries
    rp, ci, v = gen.stencil27(12)
    o = hs.bandk_host(rp, ci, 2, [32])
    r2, _, _ = hs.reorder_host(o["perm"], rp, ci, v)
    a = hs.partition_rows(r2.astype(np.int32), 4, align_ptr=o["sr_ptr"])
    assert np.array_equal(a, PO.partition_rows(r2, 4, o["sr_ptr"]))
    assert set(a.tolist()) <= set(o["sr_ptr"].tolist())


def test_row_length_percentile(hs, gen):
    for rp, ci, v in (gen.stencil27(9), gen.random_ragged(5000, max_len=80), gen.shell_blocks(20000), gen.rmat(10, 8)):
        for pct in (0, 1, 50, 90, 99, 100):
            assert hs.row_len_percentile(rp, pct) == PO.row_len_percentile(rp, pct), pct
    assert hs.row_len_percentile(np.array([0, 3], np.int32), 90) == 3


def test_bins_and_tiles_bit_exact(hs, gen):
    lens = np.array([0, 1, 4, 5, 8, 9, 16, 17, 32, 33, 64, 65, 4096, 4097, 0, 3], dtype=np.int64)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rows, bp = hs.bin_rows_host(rp)
    orows, obp = PO.bin_rows(rp)
    assert np.array_equal(rows, orows) and np.array_equal(bp, obp)
    assert PO.bin_of_len(lens).tolist() == [0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 0, 1]
    for rp in (gen.random_ragged(5000, max_len=70)[0], gen.stencil27(9)[0], rp):
        for tile in (4, 64, 4096):
            assert np.array_equal(hs.merge_tile_rows_host(rp, tile), PO.tile_rows(rp, tile))


def test_block_rows_bit_exact(hs, gen):
    """CSR_BLOCK row blocks: library rule == oracle/partition_oracle.py:block_rows, and the invariants the kernel needs."""
    lens = np.array([0, 0, 5, 300, 1, 2, 3, 0, 700, 4, 4, 4, 4, 4, 4, 0], dtype=np.int64)
    rps = [np.concatenate([[0], np.cumsum(lens)]).astype(np.int32), gen.random_ragged(5000, max_len=70)[0], gen.stencil27(9)[0],
           gen.rmat(10, 8)[0], np.zeros(1, np.int32), np.zeros(8, np.int32)]
    for rp in rps:
        m = len(rp) - 1
        for cap, max_rows in ((256, 256), (2048, 2048), (64, 3), (16384, 2048)):
            b = hs.block_rows_host(rp, cap, max_rows)
            assert np.array_equal(b, PO.block_rows(rp, cap, max_rows)), (m, cap, max_rows)
            if m == 0:
                continue
            assert b[0] == 0 and b[-1] == m and np.all(np.diff(b) > 0)
            nnz = rp[b[1:]].astype(np.int64) - rp[b[:-1]]
            rows = np.diff(b)
            assert np.all((nnz <= cap) | (rows == 1)) and np.all(rows <= max_rows)


def test_bandk_grouping_forward_scan_equals_the_literal_rule(hs):
    """handCoarsen's greedy grouping (csrk.cpp:1278-1292) two ways: the row-by-row loop and the forward scan of
    row_ptr the product uses -- same starts and same group of every row, for random row lengths (empty rows, empty
    tails, rows longer than the threshold) and every small threshold, T = 0 included."""
    rng = np.random.default_rng(5)
    cases = 0
    for trial in range(300):
        n = int(rng.integers(1, 200))
        kind = trial % 4
        if kind == 0:
            lens = rng.integers(0, 4, n)
        elif kind == 1:
            lens = rng.integers(0, 60, n)
        elif kind == 2:
            lens = np.where(rng.random(n) < 0.7, 0, rng.integers(1, 9, n))
        else:
            lens = np.full(n, int(rng.integers(0, 5)))
        if trial % 5 == 0:
            lens[int(rng.integers(0, n)):] = 0                       # trailing empty rows: no closing start
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
        for T in [0, 1, 2, 3, int(rng.integers(1, 100)), int(rp[-1]) + 1]:
            a_start, a_grp = hs.bandk_group_rows_host(rp, T, literal=True)
            b_start, b_grp = hs.bandk_group_rows_host(rp, T, literal=False)
            assert np.array_equal(a_start, b_start), (trial, T, lens.tolist())
            assert np.array_equal(a_grp, b_grp), (trial, T, lens.tolist())
            cases += 1
    assert cases == 1800
    # the literal rule itself, restated in Python, on one case
    rp = np.array([0, 3, 3, 10, 11, 11, 12, 20], np.uint32)
    start, grp = hs.bandk_group_rows_host(rp, 4, literal=True)
    acc, g, want_s, want_g = 0, 0, [0], []
    for v in range(7):
        ln = int(rp[v + 1] - rp[v])
        if acc < 4:
            acc += ln
        else:
            g, acc = g + 1, ln
            want_s.append(v)
        want_g.append(g)
    want_s.append(7)
    assert start.tolist() == want_s and grp.tolist() == want_g


def test_row_walks_do_not_depend_on_the_thread_count(hs):
    """bin_rows / row_len_percentile run on all host cores above 65536 rows: same answer as the numpy rule with 1, 3
    and all threads (rows of a bin stay in ascending order)."""
    import ctypes as C
    rng = np.random.default_rng(3)
    m = 150_001
    lens = rng.integers(0, 70, m)
    lens[rng.integers(0, m, 50)] = 5000
    lens[rng.integers(0, m, 5000)] = 0
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    want_rows, want_bp = PO.bin_rows(rp)
    gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
    before = gomp.omp_get_max_threads()
    try:
        for nt in (1, 3, before):
            gomp.omp_set_num_threads(nt)
            rows, bp = hs.bin_rows_host(rp)
            assert np.array_equal(rows, want_rows) and np.array_equal(bp, want_bp), nt
            for pct in (50, 90, 100):
                assert hs.row_len_percentile(rp, pct) == PO.row_len_percentile(rp, pct), (nt, pct)
    finally:
        gomp.omp_set_num_threads(before)


def test_pipelined_host_call_plan_properties(hs):
    """The host rules behind hspmv_spmv_host's pipeline (partition.cpp: pipe_chunks / pipe_pieces): equal row chunks
    cover [0, m) in order with 32 B-aligned bounds, never more than 32; x pieces cover [0, n) in order; every column a
    chunk references lies in a piece the chunk waits for; and for a banded matrix chunk c waits for piece c only (the
    point of cutting the pieces at the chunks' column ranges)."""
    rng = np.random.default_rng(11)
    sizes = [4_194_304, 4_200_001, 8_000_000, 13_824_000, 64_000_000, 200_000_000, 2_147_483_000] + [int(v) for v in rng.integers(4_194_304, 300_000_000, 20)]
    for m in sizes:
        for vs in (8, 4):
            if m * vs < (32 << 20):
                continue
            row_lo = hs.pipe_plan_host(m, m, vs)
            C = len(row_lo) - 1
            assert 4 <= C <= 32 and row_lo[0] == 0 and row_lo[-1] == m, (m, vs, C)
            d = np.diff(row_lo)
            assert np.all(d > 0) and np.all(row_lo[1:-1] % 4 == 0), (m, vs)
            assert d.max() - d.min() <= 8 and C == min(32, max(4, m * vs // (4 << 20))), (m, vs, C)     # equal chunks of ~4 MB, at most 32
            # banded matrix: chunk c references columns up to its last row + a half bandwidth
            for half_band in (0, 1, 160_401, m // 50):
                cmax = np.minimum(row_lo[1:] - 1 + half_band, m - 1)
                rl, x_lo, need = hs.pipe_plan_host(m, m, vs, cmax)
                assert np.array_equal(rl, row_lo)
                assert x_lo[0] == 0 and x_lo[-1] == m and np.all(np.diff(x_lo) >= 0) and np.all((x_lo % 4 == 0) | (x_lo == m))
                for c in range(C):
                    assert cmax[c] < x_lo[need[c] + 1], (m, vs, half_band, c)          # everything chunk c reads has been uploaded
                    assert need[c] <= c, (m, vs, half_band, c, need[c])                # ... by piece c at the latest
            # arbitrary column ranges (unsorted, empty chunks, a chunk that references the last column)
            for trial in range(5):
                cmax = rng.integers(-1, m, C)
                cmax[rng.integers(0, C)] = m - 1
                cmax[rng.integers(0, C)] = -1
                _, x_lo, need = hs.pipe_plan_host(m, m, vs, cmax)
                assert x_lo[0] == 0 and x_lo[-1] == m and np.all(np.diff(x_lo) >= 0)
                for c in range(C):
                    assert 0 <= need[c] < C
                    assert cmax[c] < x_lo[need[c] + 1], (m, vs, trial, c)
                    if need[c] > 0:
                        assert x_lo[need[c]] <= cmax[c], (m, vs, trial, c)            # and no later piece than necessary


def test_hub_column_order_is_a_stable_sort_by_descending_count(hs):
    """hspmv_relabel_hub_columns' host rule (partition.cpp: hub_order): columns by descending reference count, ascending
    id inside a count -- against numpy's stable sort, for power-law counts, ties, all-equal counts, one count larger
    than n (histograms larger than the data: single-core path), and with 1, 3 and all threads above 65536 columns."""
    import ctypes as C
    rng = np.random.default_rng(21)

    def check(cnt, tag):
        perm, inv = hs.hub_order_host(cnt)
        want = np.argsort(-cnt.astype(np.int64), kind="stable").astype(np.uint32)
        assert np.array_equal(perm, want), tag
        assert np.array_equal(inv[perm], np.arange(len(cnt), dtype=np.uint32)), tag

    check(np.zeros(0, np.uint32), "empty")
    check(np.array([5], np.uint32), "one")
    for trial in range(40):
        n = int(rng.integers(1, 3000))
        kind = trial % 4
        if kind == 0:
            cnt = rng.integers(0, 4, n)
        elif kind == 1:
            cnt = np.minimum(4.0 / rng.random(n) ** 2, 3e5).astype(np.int64)          # power law
        elif kind == 2:
            cnt = np.full(n, int(rng.integers(0, 9)))
        else:
            cnt = rng.integers(0, 50, n)
            cnt[int(rng.integers(0, n))] = (10 * n + 7) if trial % 8 == 3 else 2_000_000_000    # a hub referenced more often than there are columns (far more: comparison-sort path)
        check(cnt.astype(np.uint32), (trial, kind))
    gomp = C.CDLL("libgomp.so.1", mode=C.RTLD_GLOBAL)
    before = gomp.omp_get_max_threads()
    try:
        big = np.minimum(4.0 / rng.random(200_003) ** 2, 2.4e5).astype(np.uint32)
        flat = rng.integers(0, 3, 150_000).astype(np.uint32)
        for nt in (1, 3, before):
            gomp.omp_set_num_threads(nt)
            check(big, ("big", nt))
            check(flat, ("flat", nt))
    finally:
        gomp.omp_set_num_threads(before)
