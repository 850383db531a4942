"""TEST INFRASTRUCTURE ONLY -- ctypes front-end to the CPU oracle.

Two layers, both test-only (see oracle/spmv_oracle.c header):

* ``liboracle.so``  -- OUR restatement of the reference's algorithms (spmv_oracle.c,
  bandk_oracle.cpp).  Always available after ``make -C oracle``.
* ``_ref/libref_*.so`` -- the REAL reference (unmodified /root/reference sources compiled by
  ``make -C oracle ref``).  Built in the CPU container; travels to the GPU box prebuilt.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = {}

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _fp(dtype):
    return np.ctypeslib.ndpointer(dtype=dtype, flags="C_CONTIGUOUS")


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError("oracle supports float32/float64 only, got %s" % dtype)


def build(ref=True, quiet=True):
    """Compile liboracle.so and, when /root/reference is present, the _ref libraries."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE], stdout=out, stderr=out)
    if ref and os.path.isdir("/root/reference/spmv-csrk"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=out, stderr=out)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _LIB = C.CDLL(path)
        _LIB.csr_spmv_omp_timed_f32.restype = C.c_double
        _LIB.csr_spmv_omp_timed_f64.restype = C.c_double
    return _LIB


def have_ref():
    return all(os.path.exists(os.path.join(HERE, "_ref", "libref_%s_%s.so" % (k, s)))
               for k in ("csr", "csrk") for s in ("f32", "f64"))


def ref(kind, dtype):
    """kind in {'csr','csrk'} -> CDLL of the compiled reference, or raises."""
    key = (kind, _sfx(dtype))
    if key not in _REF:
        path = os.path.join(HERE, "_ref", "libref_%s_%s.so" % key)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` in the build container)")
        L = C.CDLL(path)
        if kind == "csrk":
            L.ref_csrk_create.restype = C.c_void_p
            L.ref_csrk_num_coarsest.restype = C.c_long
            L.ref_csrk_map_len.restype = C.c_long
            L.ref_csrk_spmv_timed.restype = C.c_double
        else:
            L.ref_csr_spmv_timed.restype = C.c_double
        _REF[key] = L
    return _REF[key]


# --------------------------------------------------------------------------- generators
def stencil27(g, row_lo=0, row_hi=None, dtype=np.float64, mode="classic"):
    """rows [row_lo,row_hi) of the 27-point g^3 stencil (same arrays as generators.stencil27, built
    by the C closed form in parallel): (rp int32 rebased, ci int32 GLOBAL ids, val)."""
    n = g ** 3
    row_hi = n if row_hi is None else row_hi
    L = lib()
    L.gen_stencil27_before.restype = C.c_long
    L.gen_stencil27_before.argtypes = [C.c_long, C.c_long]
    nnz = L.gen_stencil27_before(row_hi, g) - L.gen_stencil27_before(row_lo, g)
    rp = np.empty(row_hi - row_lo + 1, np.int32)
    ci = np.empty(nnz, np.int32)
    val = np.empty(nnz, dtype)
    fn = getattr(L, "gen_stencil27_" + _sfx(dtype))
    fn.argtypes = [C.c_long, C.c_long, C.c_long, C.c_int, i32p, i32p, _fp(dtype)]
    fn(g, row_lo, row_hi, 0 if mode == "classic" else 1, rp, ci, val)
    return rp, ci, val


# --------------------------------------------------------------------------- restatement
def csr_spmv(rp, ci, v, x, parallel=False):
    """y = A x, one accumulator per row, left to right (spmv-csr/spmv.c:68-90 / :92-114)."""
    v = np.ascontiguousarray(v)
    sfx = _sfx(v.dtype)
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=v.dtype)
    m = len(rp) - 1
    y = np.zeros(m, dtype=v.dtype)
    fn = getattr(lib(), ("csr_spmv_omp_" if parallel else "csr_spmv_serial_") + sfx)
    fn.argtypes = [C.c_long, i32p, i32p, _fp(v.dtype), _fp(v.dtype), _fp(v.dtype)]
    fn(m, rp, ci, v, x, y)
    return y


def csr_spmv_timed(rp, ci, v, x, warmup=5, iters=20):
    """min wall seconds of the OpenMP port (the loop of spmv-csr/spmv.c:164-181)."""
    v = np.ascontiguousarray(v)
    sfx = _sfx(v.dtype)
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=v.dtype)
    m = len(rp) - 1
    y = np.zeros(m, dtype=v.dtype)
    fn = getattr(lib(), "csr_spmv_omp_timed_" + sfx)
    fn.argtypes = [C.c_long, i32p, i32p, _fp(v.dtype), _fp(v.dtype), _fp(v.dtype), C.c_int, C.c_int]
    return float(fn(m, rp, ci, v, x, y, warmup, iters)), y


def abs_rowsum(rp, ci, v, x):
    """sum_j |a_ij x_j| in double: the scale of the north_star tolerance."""
    v = np.ascontiguousarray(v)
    sfx = _sfx(v.dtype)
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=v.dtype)
    m = len(rp) - 1
    out = np.zeros(m, dtype=np.float64)
    fn = getattr(lib(), "abs_rowsum_" + sfx)
    fn.argtypes = [C.c_long, i32p, i32p, _fp(v.dtype), _fp(v.dtype), f64p]
    fn(m, rp, ci, v, x, out)
    return out


def csrk_spmv(k, maps, rp, ci, v, xp):
    """CSR-k loop nest (csrk.cpp:165-285) on an already reordered matrix and permuted x.
    maps = [sr_ptr] (k=2) or [sr_ptr, ssr_ptr] (k=3)."""
    v = np.ascontiguousarray(v)
    sfx = _sfx(v.dtype)
    rp = np.ascontiguousarray(rp, dtype=np.uint32)
    ci = np.ascontiguousarray(ci, dtype=np.uint32)
    xp = np.ascontiguousarray(xp, dtype=v.dtype)
    y = np.zeros(len(rp) - 1, dtype=v.dtype)
    fp = _fp(v.dtype)
    if k == 2:
        sr = np.ascontiguousarray(maps[0], dtype=np.uint32)
        fn = getattr(lib(), "csrk2_spmv_" + sfx)
        fn.argtypes = [C.c_long, u32p, u32p, u32p, fp, fp, fp]
        fn(len(sr) - 1, sr, rp, ci, v, xp, y)
    elif k == 3:
        sr = np.ascontiguousarray(maps[0], dtype=np.uint32)
        ssr = np.ascontiguousarray(maps[1], dtype=np.uint32)
        fn = getattr(lib(), "csrk3_spmv_" + sfx)
        fn.argtypes = [C.c_long, u32p, u32p, u32p, u32p, fp, fp, fp]
        fn(len(ssr) - 1, ssr, sr, rp, ci, v, xp, y)
    else:
        raise ValueError("k must be 2 or 3")
    return y


def read_csr_text(path, base):
    m, n, nnz = C.c_int(), C.c_int(), C.c_int()
    rp, ci, v = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    L = lib()
    rc = L.read_csr_text(path.encode(), int(base), C.byref(m), C.byref(n), C.byref(nnz),
                         C.byref(rp), C.byref(ci), C.byref(v))
    if rc:
        raise IOError("oracle read_csr_text(%s) failed with code %d" % (path, rc))
    try:
        out = (m.value, n.value,
               np.ctypeslib.as_array(rp, (m.value + 1,)).copy(),
               np.ctypeslib.as_array(ci, (max(nnz.value, 1),))[: nnz.value].copy(),
               np.ctypeslib.as_array(v, (max(nnz.value, 1),))[: nnz.value].copy())
    finally:
        for p in (rp, ci, v):
            L.oracle_free(p)
    return out


def bandk_build(rp, ci, k, sizes):
    """Band-k (csrk.cpp:841-1067).  Returns dict(perm, sr_ptr[, ssr_ptr])."""
    rp = np.ascontiguousarray(rp, dtype=np.uint32)
    ci = np.ascontiguousarray(ci, dtype=np.uint32)
    n = len(rp) - 1
    perm = np.zeros(n, dtype=np.uint32)
    map1 = np.zeros(n + 1, dtype=np.uint32)
    map2 = np.zeros(n + 1, dtype=np.uint32)
    n1, n2 = C.c_long(0), C.c_long(0)
    sz = (C.c_int * max(1, k - 1))(*[int(s) for s in sizes[: k - 1]])
    fn = lib().bandk_oracle_build
    fn.argtypes = [C.c_long, C.c_long, u32p, u32p, C.c_int, C.POINTER(C.c_int), u32p, u32p,
                   C.POINTER(C.c_long), u32p, C.POINTER(C.c_long)]
    rc = fn(n, len(ci), rp, ci, k, sz, perm, map1, C.byref(n1), map2, C.byref(n2))
    if rc:
        raise RuntimeError("bandk_oracle_build failed with code %d" % rc)
    out = {"perm": perm, "sr_ptr": map1[: n1.value + 1].copy()}
    if k == 3:
        out["ssr_ptr"] = map2[: n2.value + 1].copy()
    return out


def bandk_reorder(perm, rp, ci, v):
    """reorderA (csrk.cpp:548-676): returns (rp', ci', v') of P A P^T with sorted rows."""
    v = np.ascontiguousarray(v)
    sfx = _sfx(v.dtype)
    rp = np.ascontiguousarray(rp, dtype=np.uint32)
    ci = np.ascontiguousarray(ci, dtype=np.uint32)
    perm = np.ascontiguousarray(perm, dtype=np.uint32)
    n = len(rp) - 1
    r2 = np.zeros(n + 1, dtype=np.uint32)
    c2 = np.zeros(len(ci), dtype=np.uint32)
    v2 = np.zeros(len(ci), dtype=v.dtype)
    fn = getattr(lib(), "bandk_oracle_reorder_" + sfx)
    fp = _fp(v.dtype)
    fn.argtypes = [C.c_long, u32p, u32p, u32p, fp, u32p, u32p, fp]
    fn(n, perm, rp, ci, v, r2, c2, v2)
    return r2, c2, v2


# --------------------------------------------------------------------------- STS restatement (sts_oracle.cpp)
def sts_levels(rp, ci):
    """find_levels + levels sorted by size (csrk.cpp:2704-2816, 1612-1666): dict(perm new->old, pack_ptr)."""
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    n = len(rp) - 1
    perm = np.zeros(n, np.uint32)
    pack = np.zeros(n + 2, np.int32)
    npk = C.c_int(0)
    fn = lib().sts_oracle_levels
    fn.argtypes = [C.c_long, C.c_long, i32p, i32p, u32p, i32p, C.POINTER(C.c_int)]
    rc = fn(n, len(ci), rp, ci, perm, pack, C.byref(npk))
    if rc:
        raise RuntimeError("sts_oracle_levels: the level search stalled (no stored diagonal, or unsymmetric structure)")
    return {"perm": perm, "pack_ptr": pack[: npk.value + 1].copy()}


def sts_lower(rp, ci, v):
    """L (column <= row) of an already reordered matrix (csrk.cpp:708-789): (l_rp, l_ci, l_val)."""
    v = np.ascontiguousarray(v)
    rp = np.ascontiguousarray(rp, dtype=np.uint32)
    ci = np.ascontiguousarray(ci, dtype=np.uint32)
    n = len(rp) - 1
    l_rp = np.zeros(n + 1, np.uint32)
    l_ci = np.zeros(len(ci), np.uint32)
    pos = np.zeros(len(ci), np.int64)
    fn = lib().sts_oracle_lower
    fn.restype = C.c_long
    fn.argtypes = [C.c_long, u32p, u32p, u32p, u32p, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
    cnt = fn(n, rp, ci, l_rp, l_ci, pos)
    if cnt < 0:
        raise RuntimeError("sts_oracle_lower: a row has no diagonal entry")
    return l_rp, l_ci[:cnt].copy(), v[pos[:cnt]].copy()


def sts_b(l_rp, l_ci, l_val):
    """compute_b (csrk.cpp:791-809): b = L * 1."""
    l_val = np.ascontiguousarray(l_val)
    b = np.zeros(len(l_rp) - 1, l_val.dtype)
    fn = getattr(lib(), "sts_oracle_b_" + _sfx(l_val.dtype))
    fn.argtypes = [C.c_long, u32p, u32p, _fp(l_val.dtype), _fp(l_val.dtype)]
    fn(len(l_rp) - 1, np.ascontiguousarray(l_rp, np.uint32), np.ascontiguousarray(l_ci, np.uint32), l_val, b)
    return b


def sts_solve(pack_ptr, l_rp, l_ci, l_val, b):
    """lowerSTS, k == 2 (csrk.cpp:79-89): x with L x = b, packs in order."""
    l_val = np.ascontiguousarray(l_val)
    b = np.ascontiguousarray(b, dtype=l_val.dtype)
    x = np.zeros(len(l_rp) - 1, l_val.dtype)
    pack_ptr = np.ascontiguousarray(pack_ptr, dtype=np.int32)
    fn = getattr(lib(), "sts_oracle_solve_" + _sfx(l_val.dtype))
    fp = _fp(l_val.dtype)
    fn.argtypes = [C.c_long, C.c_int, i32p, u32p, u32p, fp, fp, fp]
    fn(len(l_rp) - 1, len(pack_ptr) - 1, pack_ptr, np.ascontiguousarray(l_rp, np.uint32), np.ascontiguousarray(l_ci, np.uint32), l_val, b, x)
    return x


# --------------------------------------------------------------------------- real reference
class RefSTS:
    """The reference's STS-2 path with LS ordering (ctor "STS"/"LS"/k=2, putInCSRkFormat, incomplete_choloskey,
    compute_b), via oracle/_ref."""

    def __init__(self, rp, ci, v):
        v = np.ascontiguousarray(v)
        self.dtype = v.dtype
        self.L = ref("csrk", v.dtype)
        self.n = len(rp) - 1
        rp = np.ascontiguousarray(rp, dtype=np.uint32)
        ci = np.ascontiguousarray(ci, dtype=np.uint32)
        self.L.ref_sts_create.restype = C.c_void_p
        self.L.ref_sts_create.argtypes = [C.c_long, C.c_long, u32p, u32p, _fp(v.dtype)]
        self.h = C.c_void_p(self.L.ref_sts_create(self.n, len(ci), rp, ci, v))
        self.L.ref_sts_nnz_l.restype = C.c_long
        self.L.ref_sts_nnz_l.argtypes = [C.c_void_p]
        self.L.ref_sts_num_packs.argtypes = [C.c_void_p]
        self.L.ref_sts_solve_timed.restype = C.c_double

    def arrays(self):
        npk = self.L.ref_sts_num_packs(self.h)
        nl = self.L.ref_sts_nnz_l(self.h)
        perm = np.zeros(self.n, np.uint32)
        pack = np.zeros(npk + 1, np.int32)
        l_rp = np.zeros(self.n + 1, np.uint32)
        l_ci = np.zeros(nl, np.uint32)
        l_val = np.zeros(nl, self.dtype)
        b = np.zeros(self.n, self.dtype)
        fp = _fp(self.dtype)
        self.L.ref_sts_get.argtypes = [C.c_void_p, u32p, i32p, u32p, u32p, fp, fp]
        self.L.ref_sts_get(self.h, perm, pack, l_rp, l_ci, l_val, b)
        return {"perm": perm, "pack_ptr": pack, "l_rp": l_rp, "l_ci": l_ci, "l_val": l_val, "b": b}

    def set_b(self, b):
        b = np.ascontiguousarray(b, dtype=self.dtype)
        self.L.ref_sts_set_b.argtypes = [C.c_void_p, _fp(self.dtype)]
        self.L.ref_sts_set_b(self.h, b)

    def solve(self):
        x = np.zeros(self.n, self.dtype)
        self.L.ref_sts_solve.argtypes = [C.c_void_p, _fp(self.dtype)]
        self.L.ref_sts_solve(self.h, x)
        return x

    def solve_timed(self, warmup=3, iters=10):
        self.L.ref_sts_solve_timed.argtypes = [C.c_void_p, C.c_int, C.c_int]
        return float(self.L.ref_sts_solve_timed(self.h, warmup, iters))


class RefCSRk:
    """The reference's CSRk_Graph (ctor + putInCSRkFormat), via oracle/_ref."""

    def __init__(self, rp, ci, v, k, sizes):
        v = np.ascontiguousarray(v)
        self.dtype = v.dtype
        self.L = ref("csrk", v.dtype)
        self.n = len(rp) - 1
        self.k = k
        rp = np.ascontiguousarray(rp, dtype=np.uint32)
        ci = np.ascontiguousarray(ci, dtype=np.uint32)
        self.nnz = len(ci)
        sz = (C.c_int * max(1, k - 1))(*[int(s) for s in sizes[: k - 1]])
        self.L.ref_csrk_create.argtypes = [C.c_long, C.c_long, u32p, u32p, _fp(v.dtype), C.c_int,
                                           C.POINTER(C.c_int)]
        self.h = C.c_void_p(self.L.ref_csrk_create(self.n, self.nnz, rp, ci, v, k, sz))

    def maps(self):
        """[sr_ptr] for k=2, [sr_ptr, ssr_ptr] for k=3."""
        out = []
        for level in range(1, self.k):
            self.L.ref_csrk_map_len.argtypes = [C.c_void_p, C.c_int]
            ln = self.L.ref_csrk_map_len(self.h, level)
            a = np.zeros(ln + 1, dtype=np.uint32)
            self.L.ref_csrk_get_map.argtypes = [C.c_void_p, C.c_int, u32p]
            self.L.ref_csrk_get_map(self.h, level, a)
            out.append(a)
        return out

    def matrix(self):
        perm = np.zeros(self.n, dtype=np.uint32)
        r = np.zeros(self.n + 1, dtype=np.uint32)
        c = np.zeros(self.nnz, dtype=np.uint32)
        v = np.zeros(self.nnz, dtype=self.dtype)
        self.L.ref_csrk_get_matrix.argtypes = [C.c_void_p, u32p, u32p, u32p, _fp(self.dtype)]
        self.L.ref_csrk_get_matrix(self.h, perm, r, c, v)
        return perm, r, c, v

    def spmv(self, x):
        """y in permuted row order: y[i] = (A x)[perm[i]] (spmv-csrk/spmv.cpp:203)."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.zeros(self.n, dtype=self.dtype)
        self.L.ref_csrk_setx.argtypes = [C.c_void_p, _fp(self.dtype)]
        self.L.ref_csrk_setx(self.h, x)
        self.L.ref_csrk_spmv.argtypes = [C.c_void_p, _fp(self.dtype)]
        self.L.ref_csrk_spmv(self.h, y)
        return y

    def spmv_timed(self, x, warmup=5, iters=20):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.zeros(self.n, dtype=self.dtype)
        self.L.ref_csrk_setx.argtypes = [C.c_void_p, _fp(self.dtype)]
        self.L.ref_csrk_setx(self.h, x)
        self.L.ref_csrk_spmv_timed.argtypes = [C.c_void_p, _fp(self.dtype), C.c_int, C.c_int]
        return float(self.L.ref_csrk_spmv_timed(self.h, y, warmup, iters)), y

    def close(self):
        if self.h:
            self.L.ref_csrk_destroy.argtypes = [C.c_void_p]
            self.L.ref_csrk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ref_csr_spmv(rp, ci, v, x, serial=False):
    """The reference's omp_spmv (spmv-csr/spmv.c:92-114) or test_spmv (:68-90)."""
    v = np.ascontiguousarray(v)
    L = ref("csr", v.dtype)
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=v.dtype)
    m = len(rp) - 1
    y = np.zeros(m, dtype=v.dtype)
    fn = L.test_spmv if serial else L.omp_spmv
    fp = _fp(v.dtype)
    fn.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, fp, fp, fp]
    fn.restype = None
    fn(m, m, len(ci), rp, ci, v, x, y)
    return y


def ref_csr_spmv_timed(rp, ci, v, x, warmup=5, iters=20):
    v = np.ascontiguousarray(v)
    L = ref("csr", v.dtype)
    rp = np.ascontiguousarray(rp, dtype=np.int32)
    ci = np.ascontiguousarray(ci, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=v.dtype)
    m = len(rp) - 1
    y = np.zeros(m, dtype=v.dtype)
    fp = _fp(v.dtype)
    L.ref_csr_spmv_timed.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, fp, fp, fp, C.c_int,
                                     C.c_int, C.POINTER(C.c_double)]
    return float(L.ref_csr_spmv_timed(m, len(x), len(ci), rp, ci, v, x, y, warmup, iters, None)), y


def ref_read_csr_f32(path):
    """The reference's own 0-based fp32 reader (spmv-csr/spmv.c:11-57)."""
    L = ref("csr", np.float32)
    m, n, nnz = C.c_int(), C.c_int(), C.c_int()
    rp, ci, v = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_float)()
    L.my_read_csr.restype = None
    L.my_read_csr(path.encode(), C.byref(m), C.byref(n), C.byref(nnz), C.byref(rp), C.byref(ci),
                  C.byref(v))
    return (m.value, n.value,
            np.ctypeslib.as_array(rp, (m.value + 1,)).copy(),
            np.ctypeslib.as_array(ci, (nnz.value,)).copy(),
            np.ctypeslib.as_array(v, (nnz.value,)).copy())
