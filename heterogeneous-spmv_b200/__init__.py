"""heterogeneous-spmv_b200 -- Python face of libhspmv.so (include/hspmv.h).

The product is the C-ABI library built from csrc/ (hand-written sm_100a kernels + C++ host index
work).  This module only binds it with ctypes and mirrors, name for name, the entry points the
reference's drivers use, so that the parity tests read like the reference's own drivers:

    my_read_csr / omp_spmv               spmv-csr/spmv.h:8-12, spmv-csr/spmv.c:11-57,92-114
    CSRk_Graph(...).putInCSRkFormat()    spmv-csrk/csrk.h:309-338, cuda-spmv-csrk/cuda/csrk.cuh:341-348
      .setX .setY .getY .SpMV .getPermutation .getNumCoarsestRows

There is no CPU fallback: if the library is missing, or no sm_100 device is present, the
compute calls raise.  (The directory name has a hyphen, as the task prescribes; import it as
`heterogeneous_spmv_b200` through the shim at the repo root.)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# HSPMV_LIB: load another build of the SAME library (A/B timing of two kernel versions)
LIB_PATH = os.environ.get("HSPMV_LIB") or os.path.join(HERE, "lib", "libhspmv.so")

F32, F64 = 0, 1
ALGO_AUTO, ALGO_CSR_VECTOR, ALGO_CSR_STREAM, ALGO_CSR_MERGE, ALGO_CSRK, ALGO_CSR_BLOCK = 0, 1, 2, 3, 4, 5
ALGO_NAMES = {0: "auto", 1: "csr_vector", 2: "csr_stream", 3: "csr_merge", 4: "csrk", 5: "csr_block"}
(OPT_VEC, OPT_WARPS, OPT_STAGES, OPT_STAGE_NNZ, OPT_CTAS_PER_SM, OPT_ROWS_PER_TILE,
 OPT_L2_PERSIST_X, OPT_MERGE_TILE, OPT_THREADS, OPT_CHUNK, OPT_NO_PIPELINE, OPT_BANDK_HOST, OPT_TAIL, OPT_BLOCK_NNZ, OPT_STS_BARRIERS, OPT_STS_GROUPS) = range(1, 17)
ARCH_VOLTA, ARCH_AMPERE, ARCH_MI100, ARCH_B200 = 0, 1, 2, 3

_lib = None


class HspmvError(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/ into lib/libhspmv.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc")], stdout=out, stderr=out)


def lib():
    """The loaded C-ABI library.  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HspmvError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU or PyTorch fallback for the SpMV path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.hspmv_last_error.restype = C.c_char_p
        L.hspmv_last_kernel.restype = C.c_char_p
        L.hspmv_launch_count.restype = C.c_int64
        L.hspmv_x_device.restype = C.c_void_p
        L.hspmv_y_device.restype = C.c_void_p
        for name in ("hspmv_x_device", "hspmv_y_device", "hspmv_last_kernel", "hspmv_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc:
        raise HspmvError("hspmv error %d: %s" % (rc, lib().hspmv_last_error().decode()))


def _dt(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return F64
    if dtype == np.float32:
        return F32
    raise TypeError("value type must be float32 or float64")


def _np_dt(code):
    return np.float64 if code == F64 else np.float32


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def row_len_percentile(row_ptr, pct):
    rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
    lib().hspmv_row_len_percentile_host.restype = C.c_int64
    return int(lib().hspmv_row_len_percentile_host(C.c_int64(len(rp) - 1), _p(rp), int(pct)))


def launch_count():
    return int(lib().hspmv_launch_count())


def device_info():
    sm, l2, hbm, cc = C.c_int(), C.c_int64(), C.c_int64(), C.c_int()
    _check(lib().hspmv_device_info(C.byref(sm), C.byref(l2), C.byref(hbm), C.byref(cc)))
    return {"sm_count": sm.value, "l2_bytes": l2.value, "hbm_bytes": hbm.value, "cc": cc.value}


def set_device(i):
    _check(lib().hspmv_set_device(int(i)))


# ------------------------------------------------------------------------------ file formats
def my_read_csr(fname, index_base=0, dtype=np.float64, cache=False):
    """(m, n, nnz, row_start, col_idx, val) -- the reference's my_read_csr.  index_base=0 is the
    spmv-csr / cuda-spmv-csr reader, index_base=1 the spmv-csrk / cuda-spmv-csrk reader.
    cache=True goes through the binary side file <fname>.hspmvbin and appends cache_hit."""
    m, n, nnz = C.c_int64(), C.c_int64(), C.c_int64()
    rp, ci, val = C.c_void_p(), C.c_void_p(), C.c_void_p()
    hit = C.c_int(0)
    if cache:
        _check(lib().hspmv_read_csr_cached(fname.encode(), int(index_base), _dt(dtype), C.byref(m), C.byref(n),
                                           C.byref(nnz), C.byref(rp), C.byref(ci), C.byref(val), C.byref(hit)))
    else:
        _check(lib().hspmv_read_csr_text(fname.encode(), int(index_base), _dt(dtype), C.byref(m), C.byref(n),
                                         C.byref(nnz), C.byref(rp), C.byref(ci), C.byref(val)))
    try:
        row = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int32)), (m.value + 1,)).copy()
        col = np.ctypeslib.as_array(C.cast(ci, C.POINTER(C.c_int32)), (nnz.value + 1,))[: nnz.value].copy()
        ct = C.c_double if np.dtype(dtype) == np.float64 else C.c_float
        v = np.ctypeslib.as_array(C.cast(val, C.POINTER(ct)), (nnz.value + 1,))[: nnz.value].copy()
    finally:
        for p in (rp, ci, val):
            lib().hspmv_free_host(p)
    if cache:
        return m.value, n.value, nnz.value, row, col, v, bool(hit.value)
    return m.value, n.value, nnz.value, row, col, v


def my_write_csr(fname, row_start, col_idx, val, index_base=0, n_cols=None):
    row_start = np.ascontiguousarray(row_start, dtype=np.int32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    val = np.ascontiguousarray(val)
    m = len(row_start) - 1
    _check(lib().hspmv_write_csr_text(fname.encode(), int(index_base), _dt(val.dtype), C.c_int64(m),
                                      C.c_int64(m if n_cols is None else n_cols), C.c_int64(len(col_idx)),
                                      _p(row_start), _p(col_idx), _p(val)))


def read_csr3(fname, dtype=np.float64):
    """`.csr3` -> dict(ssr_ptr, sr_ptr, row_ptr, col_idx, val, m, n) (reformat-csr-to-csr3/stats.c:10-79)."""
    v = [C.c_int64() for _ in range(5)]
    ptrs = [C.c_void_p() for _ in range(5)]
    _check(lib().hspmv_read_csr3_text(fname.encode(), _dt(dtype), *[C.byref(t) for t in v],
                                      *[C.byref(t) for t in ptrs]))
    n_ssr, n_sr, m, n, nnz = [t.value for t in v]
    try:
        def arr(p, ct, cnt):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), (cnt + 1,))[:cnt].copy()
        out = {
            "ssr_ptr": arr(ptrs[0], C.c_uint32, n_ssr + 1), "sr_ptr": arr(ptrs[1], C.c_uint32, n_sr + 1),
            "row_ptr": arr(ptrs[2], C.c_uint32, m + 1), "col_idx": arr(ptrs[3], C.c_uint32, nnz),
            "val": arr(ptrs[4], C.c_double if np.dtype(dtype) == np.float64 else C.c_float, nnz),
            "m": m, "n": n,
        }
    finally:
        for p in ptrs:
            lib().hspmv_free_host(p)
    return out


def write_csr3(fname, ssr_ptr, sr_ptr, row_ptr, col_idx, val, n_cols=None):
    ssr_ptr = np.ascontiguousarray(ssr_ptr, dtype=np.uint32)
    sr_ptr = np.ascontiguousarray(sr_ptr, dtype=np.uint32)
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    val = np.ascontiguousarray(val)
    m = len(row_ptr) - 1
    _check(lib().hspmv_write_csr3_text(fname.encode(), _dt(val.dtype), C.c_int64(len(ssr_ptr) - 1),
                                       C.c_int64(len(sr_ptr) - 1), C.c_int64(m),
                                       C.c_int64(m if n_cols is None else n_cols), C.c_int64(len(col_idx)),
                                       _p(ssr_ptr), _p(sr_ptr), _p(row_ptr), _p(col_idx), _p(val)))


# ------------------------------------------------------------------------------ host index work
def bandk_host(row_ptr, col_idx, k, sup_row_sizes):
    """Band-k without a device: dict(perm, sr_ptr[, ssr_ptr]) -- putInCSRkFormat's index work."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    n = len(row_ptr) - 1
    perm = np.zeros(n, np.uint32)
    sr = np.zeros(n + 1, np.uint32)
    ssr = np.zeros(n + 1, np.uint32)
    n_sr, n_ssr = C.c_int64(), C.c_int64()
    sizes = (C.c_int * 2)(*(list(sup_row_sizes) + [0])[:2])
    _check(lib().hspmv_bandk_host(C.c_int64(n), C.c_int64(len(col_idx)), _p(row_ptr), _p(col_idx), int(k),
                                  sizes, _p(perm), _p(sr), C.byref(n_sr), _p(ssr), C.byref(n_ssr)))
    out = {"perm": perm, "sr_ptr": sr[: n_sr.value + 1].copy()}
    if k == 3:
        out["ssr_ptr"] = ssr[: n_ssr.value + 1].copy()
    return out


def reorder_host(perm, row_ptr, col_idx, val):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    perm = np.ascontiguousarray(perm, dtype=np.uint32)
    val = np.ascontiguousarray(val)
    n = len(row_ptr) - 1
    r2, c2, v2 = np.zeros(n + 1, np.uint32), np.zeros(len(col_idx), np.uint32), np.zeros_like(val)
    _check(lib().hspmv_reorder_host(C.c_int64(n), C.c_int64(len(col_idx)), _p(perm), _p(row_ptr), _p(col_idx),
                                    _p(val), _dt(val.dtype), _p(r2), _p(c2), _p(v2)))
    return r2, c2, v2


def auto_params(nnz, n_rows, arch=ARCH_B200):
    """(sizes[2], veclevel, block_dim_y): the spmv-auto-* heuristics (arch 0..2) or the B200 retune."""
    sizes = (C.c_int * 2)()
    vec, bdy = C.c_int(), C.c_int()
    _check(lib().hspmv_auto_params(C.c_int64(nnz), C.c_int64(n_rows), int(arch), sizes, C.byref(vec), C.byref(bdy)))
    return [sizes[0], sizes[1]], vec.value, bdy.value


def partition_rows(row_ptr, n_parts, align_ptr=None):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    out = np.zeros(n_parts + 1, np.int64)
    if align_ptr is not None:
        align_ptr = np.ascontiguousarray(align_ptr, dtype=np.uint32)
        ap, na = _p(align_ptr), len(align_ptr) - 1
    else:
        ap, na = None, 0
    _check(lib().hspmv_partition_rows(C.c_int64(len(row_ptr) - 1), _p(row_ptr), int(n_parts), ap, C.c_int64(na), _p(out)))
    return out


def block_col_range(row_ptr, col_idx, row_lo, row_hi):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    lo, hi = C.c_int64(), C.c_int64()
    _check(lib().hspmv_block_col_range(_p(row_ptr), _p(col_idx), C.c_int64(row_lo), C.c_int64(row_hi), C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def bin_rows_host(row_ptr):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    m = len(row_ptr) - 1
    rows = np.zeros(max(m, 1), np.int32)
    bp = np.zeros(9, np.int64)
    _check(lib().hspmv_bin_rows_host(C.c_int64(m), _p(row_ptr), _p(rows), _p(bp)))
    return rows[:m], bp


def block_rows_host(row_ptr, cap, max_rows):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    m = len(row_ptr) - 1
    out = np.zeros(m + 2, np.int32)
    nb = C.c_int64()
    _check(lib().hspmv_block_rows_host(C.c_int64(m), _p(row_ptr), C.c_int64(cap), C.c_int64(max_rows), _p(out), C.byref(nb)))
    return out[: nb.value + 1].copy() if m > 0 else out[:1].copy()


def bandk_group_rows_host(row_ptr, threshold, literal=False):
    """Band-k's greedy grouping of consecutive rows (csrk.cpp:1278-1292): (start[], grp[n]); literal=True runs the
    row-by-row statement of the rule, False the forward scan hspmv_build_csrk uses."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    n = len(row_ptr) - 1
    start = np.zeros(n + 2, np.uint32)
    grp = np.zeros(max(n, 1), np.uint32)
    ns = C.c_int64()
    _check(lib().hspmv_bandk_group_rows_host(C.c_int64(n), _p(row_ptr), C.c_uint32(int(threshold)), int(bool(literal)), _p(start),
                                             C.byref(ns), _p(grp)))
    return start[: ns.value].copy(), grp[:n].copy()


def hub_order_host(count):
    """Column order of relabel_hub_columns for the reference counts `count`: (perm new->old, inv old->new)."""
    count = np.ascontiguousarray(count, dtype=np.uint32)
    n = len(count)
    perm = np.zeros(max(n, 1), np.uint32)
    inv = np.zeros(max(n, 1), np.uint32)
    _check(lib().hspmv_hub_order_host(C.c_int64(n), _p(count), _p(perm), _p(inv)))
    return perm[:n].copy(), inv[:n].copy()


def pipe_plan_host(m, n, vsize=8, chunk_cmax=None):
    """Plan of the pipelined host call (hspmv_spmv_host): row chunk bounds; with chunk_cmax also (x piece bounds, need)."""
    row_lo = np.zeros(66, np.int64)
    nc = C.c_int()
    if chunk_cmax is None:
        _check(lib().hspmv_pipe_plan_host(C.c_int64(m), C.c_int64(n), int(vsize), None, _p(row_lo), C.byref(nc), None, None))
        return row_lo[: nc.value + 1].copy()
    cm = np.ascontiguousarray(chunk_cmax, dtype=np.int64)
    x_lo = np.zeros(66, np.int64)
    need = np.zeros(65, np.int32)
    _check(lib().hspmv_pipe_plan_host(C.c_int64(m), C.c_int64(n), int(vsize), _p(cm), _p(row_lo), C.byref(nc), _p(x_lo), _p(need)))
    return row_lo[: nc.value + 1].copy(), x_lo[: nc.value + 1].copy(), need[: nc.value].copy()


def sts_levels_host(row_ptr, col_idx):
    """Level sets of the triangular solve (kernel "STS", order "LS", k = 2) without a device: dict(perm new->old, pack_ptr)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    n = len(row_ptr) - 1
    perm = np.zeros(n, np.uint32)
    pack = np.zeros(n + 2, np.int32)
    npk = C.c_int64()
    _check(lib().hspmv_sts_levels_host(C.c_int64(n), C.c_int64(len(col_idx)), _p(row_ptr), _p(col_idx), _p(perm), _p(pack), C.byref(npk)))
    return {"perm": perm, "pack_ptr": pack[: npk.value + 1].copy()}


def merge_tile_rows_host(row_ptr, tile):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
    m = len(row_ptr) - 1
    n_tiles = (int(row_ptr[m]) + tile - 1) // tile
    out = np.zeros(n_tiles + 1, np.int32)
    _check(lib().hspmv_merge_tile_rows_host(C.c_int64(m), _p(row_ptr), C.c_int64(tile), _p(out)))
    return out


# ------------------------------------------------------------------------------ NVLink peer exchange
def peer_export(dptr):
    h = (C.c_ubyte * 64)()
    _check(lib().hspmv_peer_export(C.c_void_p(dptr), h))
    return bytes(h)


def peer_open(handle):
    p = C.c_void_p()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    _check(lib().hspmv_peer_open(buf, C.byref(p)))
    return p.value


def peer_close(dptr):
    _check(lib().hspmv_peer_close(C.c_void_p(dptr)))


def peer_alloc(nbytes):
    p = C.c_void_p()
    _check(lib().hspmv_peer_alloc(C.c_int64(nbytes), C.byref(p)))
    return p.value


def peer_free(dptr):
    _check(lib().hspmv_peer_free(C.c_void_p(dptr)))


def memcpy_async(dst, src, nbytes, stream):
    _check(lib().hspmv_memcpy_async(C.c_void_p(dst), C.c_void_p(src), C.c_int64(nbytes), C.c_void_p(stream)))


def stream_write(stream, addr, value):
    _check(lib().hspmv_stream_write(C.c_void_p(stream), C.c_void_p(addr), int(value)))


def fill_x_segments(d_x, dtype, segments, global_offset, scale, stream=None):
    """x_i = vector_x(global_offset + i) * scale on up to 8 ranges [(lo, hi), ...] of a device vector, one launch."""
    segs = [(int(a), int(b)) for a, b in segments]
    lo = (C.c_int64 * max(len(segs), 1))(*[a for a, _ in segs])
    hi = (C.c_int64 * max(len(segs), 1))(*[b for _, b in segs])
    _check(lib().hspmv_fill_x_segments(C.c_void_p(d_x), _dt(dtype), len(segs), lo, hi, C.c_int64(global_offset),
                                       C.c_double(scale), C.c_void_p(stream)))


def stream_wait_ge(stream, addr, value):
    _check(lib().hspmv_stream_wait_ge(C.c_void_p(stream), C.c_void_p(addr), int(value)))


# ------------------------------------------------------------------------------ device matrices
class Matrix:
    """A device-resident matrix handle (hspmv_matrix*)."""

    def __init__(self, handle, dtype):
        self.h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.dtype = np.dtype(dtype)

    # -- constructors
    @classmethod
    def from_csr(cls, row_ptr, col_idx, val, n_cols=None, index_base=0, keep_host=True):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
        val = np.ascontiguousarray(val)
        m = len(row_ptr) - 1
        h = C.c_void_p()
        _check(lib().hspmv_create_csr(C.byref(h), C.c_int64(m), C.c_int64(m if n_cols is None else n_cols),
                                      C.c_int64(len(col_idx)), _p(row_ptr), _p(col_idx), _p(val), _dt(val.dtype),
                                      int(index_base), int(keep_host)))
        return cls(h, val.dtype)

    @classmethod
    def from_csrk(cls, sr_ptr, row_ptr, col_idx, val, ssr_ptr=None, n_cols=None):
        sr_ptr = np.ascontiguousarray(sr_ptr, dtype=np.uint32)
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
        val = np.ascontiguousarray(val)
        m = len(row_ptr) - 1
        if ssr_ptr is not None:
            ssr_ptr = np.ascontiguousarray(ssr_ptr, dtype=np.uint32)
        h = C.c_void_p()
        _check(lib().hspmv_create_csrk(C.byref(h), C.c_int64(0 if ssr_ptr is None else len(ssr_ptr) - 1),
                                       C.c_int64(len(sr_ptr) - 1), C.c_int64(m),
                                       C.c_int64(m if n_cols is None else n_cols), C.c_int64(len(col_idx)),
                                       None if ssr_ptr is None else _p(ssr_ptr), _p(sr_ptr), _p(row_ptr),
                                       _p(col_idx), _p(val), _dt(val.dtype)))
        return cls(h, val.dtype)

    @classmethod
    def stencil(cls, kind, g, dtype=np.float64, mode=0, row_lo=0, row_hi=-1, col_shift=0, n_cols=0):
        h = C.c_void_p()
        _check(lib().hspmv_create_stencil(C.byref(h), int(kind), C.c_int64(g), _dt(dtype), int(mode),
                                          C.c_int64(row_lo), C.c_int64(row_hi), C.c_int64(col_shift),
                                          C.c_int64(n_cols)))
        return cls(h, dtype)

    # -- queries
    def info(self):
        m, n, nnz, n_sr, n_ssr = (C.c_int64() for _ in range(5))
        dt, k = C.c_int(), C.c_int()
        _check(lib().hspmv_matrix_info(self.h, C.byref(m), C.byref(n), C.byref(nnz), C.byref(dt), C.byref(k),
                                       C.byref(n_sr), C.byref(n_ssr)))
        return {"m": m.value, "n": n.value, "nnz": nnz.value, "dtype": dt.value, "k": k.value,
                "n_sr": n_sr.value, "n_ssr": n_ssr.value}

    def download(self):
        i = self.info()
        rp = np.zeros(i["m"] + 1, np.int32)
        ci = np.zeros(i["nnz"], np.int32)
        v = np.zeros(i["nnz"], self.dtype)
        _check(lib().hspmv_download_csr(self.h, _p(rp), _p(ci), _p(v)))
        return rp, ci, v

    def slice_rows(self, row_lo, row_hi, col_lo=0, n_cols=0):
        """rows [row_lo,row_hi) as a new handle, columns [col_lo, col_lo+n_cols) relocated to 0 (n_cols=0: all)"""
        out = C.c_void_p()
        _check(lib().hspmv_slice_rows(self.h, C.c_int64(row_lo), C.c_int64(row_hi), C.c_int64(col_lo), C.c_int64(n_cols),
                                      C.byref(out)))
        return Matrix(out, self.dtype)

    def row_ptr(self):
        rp = np.zeros(self.info()["m"] + 1, np.int32)
        _check(lib().hspmv_download_csr(self.h, _p(rp), None, None))
        return rp

    def row_stats(self):
        out = np.zeros(10, np.int64)
        _check(lib().hspmv_row_stats(self.h, _p(out)))
        return out

    def col_range(self):
        lo, hi = C.c_int64(), C.c_int64()
        _check(lib().hspmv_col_range_device(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def interior_rows(self, own_col_lo, own_col_hi):
        a, b = C.c_int64(), C.c_int64()
        _check(lib().hspmv_interior_rows(self.h, C.c_int64(own_col_lo), C.c_int64(own_col_hi), C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- Band-k
    def build_csrk(self, k, sup_row_sizes):
        sizes = (C.c_int * 2)(*(list(sup_row_sizes) + [0])[:2])
        _check(lib().hspmv_build_csrk(self.h, int(k), sizes))

    def build_phases(self):
        """wall seconds of the last build_csrk by phase (device builder; zeros when the host builder ran)"""
        out = (C.c_double * 6)()
        _check(lib().hspmv_build_phases(self.h, out, 6))
        names = ["host_prep", "coarse_graph_device", "rcm_levels_host", "reorder_device", "maps_stats", "total"]
        return {k: float(v) for k, v in zip(names, out)}

    def csrk(self):
        i = self.info()
        perm = np.zeros(i["m"], np.uint32)
        sr = np.zeros(i["n_sr"] + 1, np.uint32)
        ssr = np.zeros(i["n_ssr"] + 1, np.uint32)
        _check(lib().hspmv_csrk_get(self.h, _p(perm), _p(sr), _p(ssr) if i["k"] == 3 else None))
        out = {"perm": perm, "sr_ptr": sr}
        if i["k"] == 3:
            out["ssr_ptr"] = ssr
        return out

    def relabel_hub_columns(self):
        """power-law matrices: columns renumbered by popularity in HBM; returns perm (new column -> old column)"""
        perm = np.zeros(self.info()["n"], np.uint32)
        _check(lib().hspmv_relabel_hub_columns(self.h, _p(perm)))
        return perm

    # -- STS (triangular solve on the level-set-ordered matrix)
    def build_sts(self):
        _check(lib().hspmv_build_sts(self.h))

    def sts(self):
        npk = C.c_int64()
        _check(lib().hspmv_sts_get(self.h, None, None, C.byref(npk)))
        perm = np.zeros(self.info()["m"], np.uint32)
        pack = np.zeros(npk.value + 1, np.int32)
        _check(lib().hspmv_sts_get(self.h, _p(perm), _p(pack), C.byref(npk)))
        return {"perm": perm, "pack_ptr": pack}

    def sts_compute_b(self):
        _check(lib().hspmv_sts_compute_b(self.h))

    def sts_set_b(self, b, original_order=False):
        b = np.ascontiguousarray(b, dtype=self.dtype)
        assert len(b) == self.info()["m"]
        _check(lib().hspmv_sts_set_b(self.h, _p(b), int(original_order)))

    def sts_get_b(self, original_order=False):
        b = np.zeros(self.info()["m"], self.dtype)
        _check(lib().hspmv_sts_get_b(self.h, _p(b), int(original_order)))
        return b

    def sts_solve(self, stream=None):
        _check(lib().hspmv_sts_solve(self.h, C.c_void_p(stream)))

    def sts_status(self):
        t = C.c_int()
        _check(lib().hspmv_sts_status(self.h, C.byref(t)))
        return {"timed_out": bool(t.value)}

    def sts_get_x(self, original_order=False):
        x = np.zeros(self.info()["m"], self.dtype)
        _check(lib().hspmv_sts_get_x(self.h, _p(x), int(original_order)))
        return x

    def sts_solve_timed(self, warmup=3, iters=10):
        ms = np.zeros(iters, np.float32)
        _check(lib().hspmv_sts_solve_timed(self.h, int(warmup), int(iters), _p(ms)))
        return ms

    # -- vectors / hot path
    def set_option(self, opt, value):
        _check(lib().hspmv_set_option(self.h, int(opt), C.c_int64(int(value))))

    def set_x(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        assert len(x) == self.info()["n"]
        _check(lib().hspmv_set_x(self.h, _p(x)))

    def set_y(self, y):
        y = np.ascontiguousarray(y, dtype=self.dtype)
        assert len(y) == self.info()["m"]
        _check(lib().hspmv_set_y(self.h, _p(y)))

    def fill_x_synthetic(self, offset=0):
        _check(lib().hspmv_fill_x_synthetic(self.h, C.c_int64(offset)))

    def get_y(self, original_order=True):
        y = np.zeros(self.info()["m"], self.dtype)
        _check(lib().hspmv_get_y(self.h, _p(y), int(original_order)))
        return y

    def x_ptr(self):
        return lib().hspmv_x_device(self.h)

    def y_ptr(self):
        return lib().hspmv_y_device(self.h)

    def spmv(self, algo=ALGO_AUTO, d_x=None, d_y=None, stream=None):
        _check(lib().hspmv_spmv(self.h, int(algo), C.c_void_p(d_x), C.c_void_p(d_y), C.c_void_p(stream)))

    def spmv_range(self, row_lo, row_hi, algo=ALGO_AUTO, d_x=None, d_y=None, stream=None):
        _check(lib().hspmv_spmv_range(self.h, int(algo), C.c_int64(row_lo), C.c_int64(row_hi), C.c_void_p(d_x),
                                      C.c_void_p(d_y), C.c_void_p(stream)))

    def spmv_halo(self, interior_lo, interior_hi, d_x=None, d_y=None, wait_flags=None, n_wait=0, wait_step=0,
                  timed_out=None, stream=None):
        _check(lib().hspmv_spmv_halo(self.h, C.c_int64(interior_lo), C.c_int64(interior_hi), C.c_void_p(d_x),
                                     C.c_void_p(d_y), C.c_void_p(wait_flags), int(n_wait), int(wait_step),
                                     C.c_void_p(timed_out), C.c_void_p(stream)))

    def spmv_host_timed(self, x_ptr, y_ptr, iters, algo=ALGO_AUTO):
        """iters x (H2D x, kernel, D2H y) on raw host addresses; returns total device-event ms."""
        ms = C.c_float()
        _check(lib().hspmv_spmv_host_timed(self.h, int(algo), C.c_void_p(x_ptr), C.c_void_p(y_ptr), int(iters),
                                           C.byref(ms)))
        return float(ms.value)

    def spmv_timed(self, algo=ALGO_AUTO, warmup=3, iters=20, flush_l2=False):
        ms = np.zeros(iters, np.float32)
        # flush_l2: False / True, or 2 = write + read flush (L2 left full of clean foreign lines)
        _check(lib().hspmv_spmv_timed(self.h, int(algo), int(warmup), int(iters), int(flush_l2), _p(ms)))
        return ms

    def spmv_timed_batch(self, algo=ALGO_AUTO, warmup=5, iters=20):
        """ms per launch of `iters` back-to-back launches between one pair of events"""
        ms = C.c_float()
        _check(lib().hspmv_spmv_timed_batch(self.h, int(algo), int(warmup), int(iters), C.byref(ms)))
        return ms.value

    def spmv_graph(self, algo=ALGO_AUTO, iters=10, stream=None):
        """`iters` products as one CUDA-graph replay (captured on first use)"""
        _check(lib().hspmv_spmv_graph(self.h, int(algo), int(iters), C.c_void_p(stream)))

    def spmv_graph_timed(self, algo=ALGO_AUTO, iters=20, reps=5):
        ms = C.c_float()
        _check(lib().hspmv_spmv_graph_timed(self.h, int(algo), int(iters), int(reps), C.byref(ms)))
        return ms.value

    def autotune(self, algo=ALGO_AUTO, effort=1):
        """times the kernel shapes around the heuristic choice, keeps the fastest; returns its ms per launch"""
        ms = C.c_float()
        _check(lib().hspmv_autotune(self.h, int(algo), int(effort), C.byref(ms)))
        return ms.value

    def spmv_host(self, x, algo=ALGO_AUTO, out=None):
        """End-to-end call with host buffers (setX + SpMV + getY)."""
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.zeros(self.info()["m"], self.dtype) if out is None else out
        _check(lib().hspmv_spmv_host(self.h, int(algo), _p(x), _p(y)))
        return y

    def spmv_host_ptr(self, x_ptr, y_ptr, algo=ALGO_AUTO):
        """Same, on raw host addresses (e.g. pinned torch tensors' data_ptr())."""
        _check(lib().hspmv_spmv_host(self.h, int(algo), C.c_void_p(x_ptr), C.c_void_p(y_ptr)))

    def last_kernel(self):
        return lib().hspmv_last_kernel(self.h).decode()

    def merge_tile_rows(self):
        tile, n_tiles = C.c_int64(), C.c_int64()
        _check(lib().hspmv_merge_tile_rows_device(self.h, C.byref(tile), C.byref(n_tiles), None))
        out = np.zeros(n_tiles.value + 1, np.int32)
        _check(lib().hspmv_merge_tile_rows_device(self.h, C.byref(tile), C.byref(n_tiles), _p(out)))
        return tile.value, out

    def close(self):
        if self.h:
            lib().hspmv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGPU:
    """hspmv_mg_*: the matrix row-sharded over several devices of ONE process (csrc/mg.cu).  x and y are host
    vectors in the ORIGINAL numbering; halo entries of x only ever travel GPU to GPU inside spmv()."""

    def __init__(self, handle, dtype):
        self.h = handle
        self.dtype = np.dtype(dtype)

    @staticmethod
    def _mg_check(rc):
        if rc:
            lib().hspmv_mg_last_error.restype = C.c_char_p
            raise HspmvError("hspmv_mg error %d: %s" % (rc, lib().hspmv_mg_last_error().decode()))

    @staticmethod
    def _sizes(k, sizes):
        return (C.c_int * 2)(*(list(sizes or []) + [0, 0])[:2]) if k > 1 else None

    @classmethod
    def from_csr(cls, devices, row_ptr, col_idx, val, n_cols=None, index_base=0, k=1, sup_row_sizes=None):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int32)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
        val = np.ascontiguousarray(val)
        m = len(row_ptr) - 1
        dev = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        cls._mg_check(lib().hspmv_mg_create_csr(C.byref(h), dev, len(devices), C.c_int64(m), C.c_int64(m if n_cols is None else n_cols),
                                                C.c_int64(len(col_idx)), _p(row_ptr), _p(col_idx), _p(val), _dt(val.dtype),
                                                int(index_base), int(k), cls._sizes(k, sup_row_sizes)))
        return cls(h, val.dtype)

    @classmethod
    def stencil(cls, devices, kind, g, dtype=np.float64, mode=0, k=1, sup_row_sizes=None):
        dev = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        cls._mg_check(lib().hspmv_mg_create_stencil(C.byref(h), dev, len(devices), int(kind), C.c_int64(g), _dt(dtype), int(mode),
                                                    int(k), cls._sizes(k, sup_row_sizes)))
        return cls(h, dtype)

    def info(self):
        nd, dt, k, rep = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        m, n, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        self._mg_check(lib().hspmv_mg_info(self.h, C.byref(nd), C.byref(m), C.byref(n), C.byref(nnz), C.byref(dt), C.byref(k), C.byref(rep)))
        return {"n_dev": nd.value, "m": m.value, "n": n.value, "nnz": nnz.value, "dtype": dt.value, "k": k.value,
                "replicated": bool(rep.value)}

    def partition(self):
        nd = self.info()["n_dev"]
        dev = (C.c_int * nd)()
        row_lo = np.zeros(nd + 1, np.int64)
        cmin, cmax, nnz, halo = (np.zeros(nd, np.int64) for _ in range(4))
        self._mg_check(lib().hspmv_mg_partition_query(self.h, dev, _p(row_lo), _p(cmin), _p(cmax), _p(nnz), _p(halo)))
        return {"devices": list(dev), "row_lo": row_lo, "col_min": cmin, "col_max": cmax, "nnz": nnz, "halo_elems": halo}

    def perm(self):
        p = np.zeros(self.info()["m"], np.uint32)
        self._mg_check(lib().hspmv_mg_get_perm(self.h, _p(p)))
        return p

    def shard_kernel(self, d=0):
        lib().hspmv_mg_shard.restype = C.c_void_p
        return lib().hspmv_last_kernel(C.c_void_p(lib().hspmv_mg_shard(self.h, int(d)))).decode()

    def set_x(self, x):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        assert len(x) == self.info()["n"]
        self._mg_check(lib().hspmv_mg_set_x(self.h, _p(x)))

    def poison_halos(self):
        self._mg_check(lib().hspmv_mg_poison_halos(self.h))

    def spmv(self):
        self._mg_check(lib().hspmv_mg_spmv(self.h))

    def sync(self):
        self._mg_check(lib().hspmv_mg_sync(self.h))

    def get_y(self):
        y = np.zeros(self.info()["m"], self.dtype)
        self._mg_check(lib().hspmv_mg_get_y(self.h, _p(y)))
        return y

    def spmv_timed(self, warmup=3, iters=10):
        ms = np.zeros(iters, np.float32)
        self._mg_check(lib().hspmv_mg_spmv_timed(self.h, int(warmup), int(iters), _p(ms)))
        return ms

    def spmv_timed_batch(self, warmup=3, iters=10):
        ms = C.c_float()
        self._mg_check(lib().hspmv_mg_spmv_timed_batch(self.h, int(warmup), int(iters), C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            lib().hspmv_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    c = C.c_int()
    _check(lib().hspmv_device_count(C.byref(c)))
    return c.value


# ------------------------------------------------------------------------------ reference-named API
def omp_spmv(m, n, nnz, row_start, col_idx, val, x, y):
    """Drop-in for spmv-csr/spmv.h:11-12: host CSR in, host y out; the matrix is uploaded once and
    cached by array identity (hspmv_omp_spmv_f64/f32)."""
    val = np.ascontiguousarray(val)
    row_start = np.ascontiguousarray(row_start, dtype=np.int32)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.int32)
    x = np.ascontiguousarray(x, dtype=val.dtype)
    assert y.dtype == val.dtype and y.flags.c_contiguous
    fn = lib().hspmv_omp_spmv_f64 if val.dtype == np.float64 else lib().hspmv_omp_spmv_f32
    _check(fn(int(m), int(n), int(nnz), _p(row_start), _p(col_idx), _p(val), _p(x), _p(y)))
    return y


class CSRk_Graph:
    """Mirror of the reference's class (spmv-csrk/csrk.h:253-345; CUDA flavour csrk.cuh:341-348).
    Same constructor arguments, same method names, same ownership (the arrays are copied)."""

    def __init__(self, nRows, nCols, nnz, rVec, cVec, value, kernelCalled="SpMV", inOrderType="",
                 inCoarsenType="HAND", isTuned=False, inK=2, inSupRowSizes=(8,)):
        if kernelCalled not in ("SpMV", "STS"):
            raise HspmvError("kernel must be SpMV or STS")
        if kernelCalled == "STS" and (int(inK) != 2 or inOrderType != "LS"):
            raise HspmvError("STS is built for k = 2 with LS (level-set) ordering; COLOR needs Boost's colouring and k >= 3 "
                             "the coarsened level search, both out of scope")
        self.kernel = kernelCalled
        if inCoarsenType != "HAND":
            raise HspmvError("only HAND coarsening is supported: every reference driver hard-codes it (spmv-csrk/spmv.cpp:125)")
        assert len(rVec) == nRows + 1 and len(cVec) == nnz
        self.k = int(inK)
        self.sizes = list(inSupRowSizes)
        self.A = Matrix.from_csr(rVec, cVec, value, n_cols=nCols, index_base=0, keep_host=True)
        self._y = None
        self._algo = ALGO_AUTO

    def putInCSRkFormat(self):
        if self.kernel == "STS":
            self.A.build_sts()
        elif self.k >= 2:
            self.A.build_csrk(self.k, self.sizes)

    # STS flavour (spmv-csrk/csrk.h:340-344): the L / U split needs no separate step here (L = A' up to the diagonal)
    def incomplete_choloskey(self):
        self.A.sts()

    def compute_b(self):
        self.A.sts_compute_b()

    def lowerSTS(self):
        self.A.sts_solve()

    def checkError(self):
        """Total Error of csrk.cpp:833-836: sum of (x_i - 1) after compute_b + lowerSTS"""
        return float(np.sum(self.A.sts_get_x().astype(np.float64) - 1.0))

    def getPermutation(self):
        if self.kernel == "STS":
            return self.A.sts()["perm"]
        return self.A.csrk()["perm"] if self.k >= 2 else np.arange(self.A.info()["m"], dtype=np.uint32)

    def getNumCoarsestRows(self):
        i = self.A.info()
        return i["n_ssr"] if i["k"] == 3 else i["n_sr"] if i["k"] == 2 else i["m"]

    def setX(self, x):
        self.A.set_x(x)

    def setY(self, y):
        self._y = y
        self.A.set_y(y)

    def SpMV(self, y=None):
        """SpMV() launches on the device-resident vectors (CUDA flavour, csrk.cuh:347).
        SpMV(y) additionally copies the result into y in the PERMUTED row order, like the CPU
        flavour's SpMV(float*&y) (spmv-csrk/csrk.cpp:149; checked against y_csrser[perm[i]] at
        spmv-csrk/spmv.cpp:203)."""
        self.A.spmv(self._algo)
        if y is not None:
            y[:] = self.A.get_y(original_order=False)

    def getY(self):
        y = self.A.get_y(original_order=False)
        if self._y is not None:
            self._y[:] = y
            return self._y
        return y
