"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the index rules the product adds on top of
the reference (multi-GPU row partition, halo ranges, row-length bins, merge tile table).

The reference has no multi-GPU, binning or merge-path code, so there is nothing
to compile and diff against; parity for these rules is "unpinned by the reference" and pinned
instead by this independent restatement of the rule stated in include/hspmv.h:
  partition_rows   row_lo[p] = first row r with row_ptr[r] >= floor(p*nnz/P); optional snap up
                   to the next boundary of align_ptr; monotone; row_lo[0]=0, row_lo[P]=m
  col_range        min / max column referenced by a row block
  bin_of_len       0: empty, 1: <=4, 2: <=8, 3: <=16, 4: <=32, 5: <=64, 6: <=4096, 7: longer
  tile_rows        tile_row[t] = last row r with row_ptr[r] <= t*tile; tile_row[n_tiles] = m
"""
import numpy as np


def partition_rows(row_ptr, n_parts, align_ptr=None):
    rp = np.asarray(row_ptr, dtype=np.int64)
    m = len(rp) - 1
    nnz = int(rp[m])
    out = [0]
    for p in range(1, n_parts):
        target = (nnz * p) // n_parts
        r = int(np.searchsorted(rp, target, side="left"))
        r = min(r, m)
        if align_ptr is not None:
            ap = np.asarray(align_ptr, dtype=np.int64)
            j = int(np.searchsorted(ap, r, side="left"))
            r = m if j >= len(ap) else int(ap[j])
        out.append(max(r, out[-1]))
    out.append(m)
    return np.array(out, dtype=np.int64)


def col_range(row_ptr, col_idx, row_lo, row_hi):
    e0, e1 = int(row_ptr[row_lo]), int(row_ptr[row_hi])
    if e1 <= e0:
        return 0, -1
    seg = np.asarray(col_idx[e0:e1])
    return int(seg.min()), int(seg.max())


def bin_of_len(length):
    length = np.asarray(length, dtype=np.int64)
    edges = np.array([0, 4, 8, 16, 32, 64, 4096], dtype=np.int64)
    return np.searchsorted(edges, length, side="left").astype(np.int64)


def bin_rows(row_ptr):
    rp = np.asarray(row_ptr, dtype=np.int64)
    b = bin_of_len(np.diff(rp))
    order = np.argsort(b, kind="stable")
    cnt = np.bincount(b, minlength=8)
    ptr = np.zeros(9, dtype=np.int64)
    np.cumsum(cnt, out=ptr[1:])
    return order.astype(np.int32), ptr


def tile_rows(row_ptr, tile):
    rp = np.asarray(row_ptr, dtype=np.int64)
    m = len(rp) - 1
    nnz = int(rp[m])
    n_tiles = (nnz + tile - 1) // tile
    t = np.arange(n_tiles, dtype=np.int64) * tile
    out = np.searchsorted(rp, t, side="right") - 1
    return np.concatenate([out, [m]]).astype(np.int32)


def interior_rows(row_ptr, col_idx, own_lo, own_hi):
    """[row_a,row_b): row_a = 1 + last row of the first half with a column outside [own_lo,own_hi),
    row_b = first such row of the second half (include/hspmv.h: hspmv_interior_rows)."""
    rp = np.asarray(row_ptr, dtype=np.int64)
    ci = np.asarray(col_idx, dtype=np.int64)
    m = len(rp) - 1
    bad = np.zeros(m, dtype=bool)
    rows = np.repeat(np.arange(m), np.diff(rp))
    out = (ci < own_lo) | (ci >= own_hi)
    bad[np.unique(rows[out])] = True
    half = m // 2
    first = np.nonzero(bad[:half])[0]
    second = np.nonzero(bad[half:])[0]
    a = int(first[-1]) + 1 if first.size else 0
    b = int(second[0]) + half if second.size else m
    return a, max(a, b)


def halo_plan(row_lo, col_ranges, rank):
    """recv/send segments (peer, global_start, global_end) for `rank`: the parts of its column
    window [cmin,cmax] owned by others, and the parts of its own rows that others' windows cover."""
    P = len(row_lo) - 1
    recv, send = [], []
    lo, hi = int(row_lo[rank]), int(row_lo[rank + 1])
    cmin, cmax = col_ranges[rank]
    for q in range(P):
        if q == rank:
            continue
        qlo, qhi = int(row_lo[q]), int(row_lo[q + 1])
        a, b = max(cmin, qlo), min(cmax + 1, qhi)
        if a < b:
            recv.append((q, a, b))
        qmin, qmax = col_ranges[q]
        a, b = max(qmin, lo), min(qmax + 1, hi)
        if a < b:
            send.append((q, a, b))
    return recv, send


def row_len_percentile(rp, pct):
    """smallest L with at least pct % of the rows of length <= L (include/hspmv.h:
    hspmv_row_len_percentile_host)"""
    lens = np.sort(np.diff(np.asarray(rp, dtype=np.int64)))
    m = len(lens)
    if m == 0:
        return 0
    need = (m * pct + 99) // 100
    return int(lens[need - 1]) if need > 0 else 0


def block_rows(rp, cap, max_rows):
    """CSR_BLOCK row blocks (csrc/partition.cpp:block_rows): greedy consecutive rows with <= cap nonzeros and
    <= max_rows rows per block; a row longer than cap is alone in its block.  Returns blk_row (first rows + m)."""
    m = len(rp) - 1
    out, start = [0], 0
    for r in range(m):
        if r > start and (int(rp[r + 1]) - int(rp[start]) > cap or r - start >= max_rows):
            out.append(r)
            start = r
    if m > 0:
        out.append(m)
    return np.asarray(out, dtype=np.int32)
