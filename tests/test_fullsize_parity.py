"""Full-size, NON-BLIND parity for all five BASELINE.json configs (VERDICT r1 "weak #2, #3").

Every case uses row- and column-dependent matrix values (generators.value_of 'varied' / R-MAT values) and the
non-constant x of generators.vector_x -- x = 1 with -1/26 values preserves row sums under any column-index error,
which is exactly the blind spot SURVEY.md section 4 warns about.  The whole y is compared, within the north_star
bound |y - y_ref|_i <= tol * sum_j |a_ij x_j| (1e-12 fp64, 1e-5 fp32):

  cfg1-cfg3  CSR (AUTO / STREAM / VECTOR), CSR-2, CSR-3, fp64 and fp32, against the COMPILED REFERENCE
             (oracle/_ref: omp_spmv, CSRk_Graph::SpMV -- like spmv-csrk/spmv.cpp:193-211 but with x non-constant);
             Band-k index work bit-exact against the reference's maps and permutation
  cfg4       R-MAT scale 24 (torch generator): whole y of AUTO / VECTOR / MERGE against the oracle's OpenMP CSR loop
  cfg5       27-point 400^3 generated in HBM with varied values: the stream kernel against the vector kernel on all
             64 M rows, and against the oracle on three windows of 1 M rows (start, the N = 2 partition boundary, end);
             CSR-2 / CSR-3 the same through Band-k built in HBM
"""
import numpy as np
import pytest

from conftest import assert_within_bound, TOL

pytestmark = pytest.mark.gpu


def _ref_csr(O, rp, ci, v, x):
    return O.ref_csr_spmv(rp, ci, v, x) if O.have_ref() else O.csr_spmv(rp, ci, v, x, parallel=True)


CFG = {
    "cfg1": lambda G: G.laplace2d_5pt(1000, mode="varied", rcm=True),
    "cfg2": lambda G: G.stencil27(100, mode="varied"),
    "cfg3": lambda G: G.shell_blocks(1_500_000),
}
_cache = {}


def _matrix(name, G):
    if name not in _cache:
        _cache.clear()                     # one full-size matrix in host memory at a time
        _cache[name] = CFG[name](G)
    return _cache[name]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_csr_whole_y_against_the_compiled_reference(name, dtype, hs, oracle, gen):
    rp, ci, v = _matrix(name, gen)
    v = v.astype(dtype)
    n = len(rp) - 1
    x = gen.vector_x(n, dtype)
    y_ref = _ref_csr(oracle, rp, ci, v, x)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    A = hs.Matrix.from_csr(rp, ci, v)
    A.set_x(x)
    for algo in (hs.ALGO_AUTO, hs.ALGO_CSR_STREAM, hs.ALGO_CSR_VECTOR):
        A.set_y(np.full(n, np.nan, dtype))
        A.spmv(algo)
        assert_within_bound(A.get_y(), y_ref, scale, dtype, "%s %s [%s]" % (name, hs.ALGO_NAMES[algo], A.last_kernel()))
    A.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3"])
def test_csrk_whole_y_and_index_work_against_the_compiled_reference(name, k, dtype, hs, oracle, gen):
    rp, ci, v = _matrix(name, gen)
    v = v.astype(dtype)
    n = len(rp) - 1
    x = gen.vector_x(n, dtype)
    sizes, _, _ = hs.auto_params(len(ci), n, hs.ARCH_B200)
    A = hs.Matrix.from_csr(rp, ci, v)
    A.build_csrk(k, sizes)
    c = A.csrk()
    A.set_x(x)
    A.set_y(np.full(n, np.nan, dtype))
    A.spmv(hs.ALGO_CSRK)
    y_perm = A.get_y(original_order=False)
    y_orig = A.get_y(original_order=True)
    kernel = A.last_kernel()
    A.close()
    if oracle.have_ref():
        R = oracle.RefCSRk(rp, ci, v, k, sizes)
        maps = R.maps()
        perm, _, _, _ = R.matrix()
        y_ref_perm = R.spmv(x)              # CSRk_Graph::SpMV on the reordered matrix, permuted row order
        R.close()
        assert np.array_equal(c["perm"], perm), "permBigG differs from the reference"
        assert np.array_equal(c["sr_ptr"], maps[0])
        if k == 3:
            assert np.array_equal(c["ssr_ptr"], maps[1])
    else:
        o = oracle.bandk_build(rp, ci, k, sizes)
        assert all(np.array_equal(c[key], o[key]) for key in o)
        y_ref_perm = oracle.csr_spmv(rp, ci, v, x, parallel=True)[c["perm"]]
    scale = oracle.abs_rowsum(rp, ci, v, x)
    assert_within_bound(y_perm, y_ref_perm, scale[c["perm"]], dtype, "%s csr-%d permuted order [%s]" % (name, k, kernel))
    assert_within_bound(y_orig, _ref_csr(oracle, rp, ci, v, x), scale, dtype, "%s csr-%d original order" % (name, k))


def test_cfg4_rmat24_whole_y(hs, oracle, gen):
    """BASELINE configs[3]: R-MAT scale 24, ~256 M nonzeros, hub rows of 10^5 entries, empty rows."""
    rp, ci, v = gen.rmat_torch(24, 16)
    import torch
    torch.cuda.empty_cache()
    n = len(rp) - 1
    assert n == 1 << 24 and 200_000_000 < len(ci) < 270_000_000
    x = gen.vector_x(n)
    y_ref = oracle.csr_spmv(rp, ci, v, x, parallel=True)
    scale = oracle.abs_rowsum(rp, ci, v, x)
    A = hs.Matrix.from_csr(rp, ci, v, keep_host=False)
    A.set_x(x)
    for algo in (hs.ALGO_AUTO, hs.ALGO_CSR_VECTOR, hs.ALGO_CSR_MERGE):
        A.set_y(np.full(n, np.nan))
        A.spmv(algo)
        assert_within_bound(A.get_y(), y_ref, scale, np.float64, "cfg4 %s [%s]" % (hs.ALGO_NAMES[algo], A.last_kernel()))
    A.close()


@pytest.mark.parametrize("fmt", ["csr", "csr2", "csr3"])
def test_cfg5_full_size_varied_values_and_x(fmt, hs, oracle, gen):
    """BASELINE configs[4] at FULL size, not blind: varied values, vector_x."""
    import bench
    g = 400
    n = g ** 3
    A = hs.Matrix.stencil(27, g, np.float64, 1)                 # mode 1 = generators.value_of('varied')
    assert A.info()["nnz"] == (3 * g - 2) ** 3
    A.fill_x_synthetic(0)
    # sum_j |a_ij x_j| >= 32 + 7 on every row (diagonal >= 32, >= 7 off-diagonals of magnitude >= 1, x >= 1)
    uniform = TOL[np.dtype(np.float64)] * 39.0
    A.set_y(np.full(n, np.nan))
    A.spmv(hs.ALGO_CSR_VECTOR)
    y_vec = A.get_y()
    perm = None
    if fmt == "csr":
        A.set_y(np.full(n, np.nan))
        A.spmv(hs.ALGO_CSR_STREAM)
        y = A.get_y()
    else:
        k = int(fmt[-1])
        sizes, _, _ = hs.auto_params(A.info()["nnz"], n, hs.ARCH_B200)
        A.build_csrk(k, sizes)
        perm = A.csrk()["perm"]
        assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
        A.set_x(gen.vector_x(n))                                # set_x applies the permutation (setX, csrk.h:327-335)
        A.set_y(np.full(n, np.nan))
        A.spmv(hs.ALGO_CSRK)
        y = A.get_y(original_order=True)
    kernel = A.last_kernel()
    A.close()
    assert np.all(np.isfinite(y))
    worst = float(np.max(np.abs(y - y_vec)))
    assert worst <= uniform, "%s vs the vector kernel on all %d rows: max |diff| %g > %g [%s]" % (fmt, n, worst, uniform, kernel)
    cut = bench.partition_stencil(g, 2)[1]
    W = 1_000_000
    for lo in (0, cut - W // 2, n - W):
        rp, ci, v = oracle.stencil27(g, lo, lo + W, mode="varied")
        c0 = int(ci.min())
        x = gen.vector_x(int(ci.max()) + 1)[c0:]
        y_ref = oracle.csr_spmv(rp, ci - c0, v, x, parallel=True)
        scale = oracle.abs_rowsum(rp, ci - c0, v, x)
        assert_within_bound(y[lo:lo + W], y_ref, scale, np.float64, "cfg5 %s rows [%d,%d) vs the oracle [%s]" % (fmt, lo, lo + W, kernel))
