"""Synthetic stand-ins for the reference's matrix suite (no SuiteSparse offline).

The reference takes its inputs from Octave scripts (helpers/converter.m:10-49,
helpers/sparse2csr.m:1-7) that turn a MatrixMarket file into the `.mtx.rcm.csr` text layout.
Octave is not in this image, so these generators produce matrices of the five shapes that
BASELINE.json `configs` names, in host CSR arrays (int32 indices, 0-based, sorted columns):

    cfg 1  laplace2d_5pt(1000, rcm=True)      n=1M,   nnz=4,996,000
    cfg 2  stencil27(100)                     n=1M,   nnz=26,463,592 (=298^3)
    cfg 3  shell_blocks(1_500_000)            n=1.5M, 35 nnz/row in 5x5 node blocks, RCM-banded
           (banded_shell(1_500_000): the SURVEY's "3-wide runs" recipe -- every row an independent
            random column set, no row-to-row similarity; kept as the hostile variant)
    cfg 4  rmat(24, 16)                       n=16M,  nnz~256M, power law, not symmetric
    cfg 5  stencil27(400)                     n=64M,  nnz=1,719,374,392 (=1198^3)

Every value is a multiple of 2^-6 with |v| < 2^10, hence exact in fp32 and fp64 and exact
through the text format's "%f" / "%.6f" six-decimal round trip (helpers/converter.m:48,
reformat-csr-to-csr3/spmv-auto.cpp:60).

The stencil generators are closed-form so that csrc/gen.cu can produce the identical arrays
on the device for the sizes where a host copy is impractical (cfg 5: 22 GB).
"""
import numpy as np

try:  # scipy is only needed for the RCM'd / random shapes
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
except Exception:  # pragma: no cover
    sp = None


def value_of(rows, cols, mode):
    """Deterministic matrix values.
    mode 'classic': d on the diagonal (set by caller), -1 off it.
    mode 'varied' : off-diagonal -(1 + ((31 r + 17 c) mod 64)/64); diagonal 32 + ((7 r) mod 64)/64
    Same formula as gen.cu:stencil_value()."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    if mode == "varied":
        off = -(1.0 + ((31 * rows + 17 * cols) % 64) / 64.0)
        dia = 32.0 + ((7 * rows) % 64) / 64.0
        return np.where(rows == cols, dia, off)
    raise ValueError(mode)


def vector_x(n, dtype=np.float64):
    """Deterministic non-constant x (x_i = 1 + (i mod 97)/97, here /128 so that it
    is exact in fp32 too): x_i = 1 + (i mod 97)/128."""
    i = np.arange(n, dtype=np.int64)
    return (1.0 + (i % 97) / 128.0).astype(dtype)


def _line_len(t, g):
    return 3 - (t == 0).astype(np.int64) - (t == g - 1).astype(np.int64) if g > 1 else np.ones_like(t)


def stencil27_row_ptr(g):
    """Closed-form row pointer of the g^3 27-point stencil in lexicographic (x fastest) order."""
    n = g ** 3
    i = np.arange(n, dtype=np.int64)
    x, y, z = i % g, (i // g) % g, i // (g * g)
    ln = _line_len(z, g) * _line_len(y, g) * _line_len(x, g)
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(ln, out=rp[1:])
    return rp


def stencil27(g, dtype=np.float64, mode="classic", row_lo=0, row_hi=None):
    """3-D 27-point stencil on a g^3 grid, natural order (already banded, half bandwidth
    g^2+g+1).  classic: 26 on the diagonal, -1 elsewhere.  Rows [row_lo,row_hi) only if given
    (row_ptr is rebased to start at 0)."""
    n = g ** 3
    row_hi = n if row_hi is None else row_hi
    i = np.arange(row_lo, row_hi, dtype=np.int64)
    x, y, z = i % g, (i // g) % g, i // (g * g)
    cols, rows = [], []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = ((z + dz >= 0) & (z + dz < g) & (y + dy >= 0) & (y + dy < g) &
                      (x + dx >= 0) & (x + dx < g))
                rows.append(i[ok])
                cols.append(i[ok] + (dz * g + dy) * g + dx)
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    cnt = np.bincount(rows - row_lo, minlength=row_hi - row_lo)
    rp = np.zeros(row_hi - row_lo + 1, dtype=np.int64)
    np.cumsum(cnt, out=rp[1:])
    if mode == "classic":
        val = np.where(rows == cols, 26.0, -1.0)
    else:
        val = value_of(rows, cols, mode)
    return rp.astype(np.int32), cols.astype(np.int32), val.astype(dtype)


def laplace2d_5pt(g, dtype=np.float64, mode="classic", rcm=False):
    """2-D 5-point Laplacian on a g x g grid: diag 4, off-diagonal -1, neighbour order
    (-1,0),(0,-1),(0,0),(0,+1),(+1,0).  rcm=True applies scipy's reverse
    Cuthill-McKee (symmetric_mode=True) and re-sorts the columns (cfg 1)."""
    n = g * g
    i = np.arange(n, dtype=np.int64)
    c, r = i % g, i // g
    rows, cols = [], []
    for dr, dc in ((-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)):
        ok = (r + dr >= 0) & (r + dr < g) & (c + dc >= 0) & (c + dc < g)
        rows.append(i[ok])
        cols.append(i[ok] + dr * g + dc)
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    if rcm:
        A = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
        p = reverse_cuthill_mckee(A, symmetric_mode=True).astype(np.int64)
        inv = np.empty(n, dtype=np.int64)
        inv[p] = np.arange(n)
        rows, cols = inv[rows], inv[cols]
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rp[1:])
    if mode == "classic":
        val = np.where(rows == cols, 4.0, -1.0)
    else:
        val = value_of(rows, cols, mode)
    return rp.astype(np.int32), cols.astype(np.int32), val.astype(dtype)


def shell_blocks(n, dtype=np.float64, dofs=5, node_nbrs=6, node_band=600, seed=12345, rcm=True):
    """cfg 3 (af_shell / Queen analog): a finite-element shell look.  n/dofs nodes, every node
    coupled to itself and ~node_nbrs neighbour nodes inside a band of the node graph (symmetric),
    every node-node coupling a dense dofs x dofs block -> (node_nbrs+1)*dofs = 35 entries per row
    for the defaults, and the rows of one node share one column set, as in af_shell*.
    RCM is applied to the NODE graph (what a mesh-based reordering does), then expanded."""
    nodes = n // dofs
    assert nodes * dofs == n, "n must be a multiple of dofs"
    rng = np.random.default_rng(seed)
    half = max(1, node_nbrs // 2)
    base = np.repeat(np.arange(nodes, dtype=np.int64), half)
    nb = base + rng.integers(1, node_band, size=base.size, dtype=np.int64)
    ok = nb < nodes
    base, nb = base[ok], nb[ok]
    r = np.concatenate([base, nb, np.arange(nodes, dtype=np.int64)])
    c = np.concatenate([nb, base, np.arange(nodes, dtype=np.int64)])
    Gn = sp.csr_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(nodes, nodes))
    Gn.sum_duplicates()
    Gn.data[:] = 1
    if rcm:
        p = reverse_cuthill_mckee(Gn, symmetric_mode=True)
        Gn = Gn[p][:, p].tocsr()
    Gn.sort_indices()
    nrp = Gn.indptr.astype(np.int64)
    ncol = Gn.indices.astype(np.int64)
    ndeg = np.diff(nrp)
    # expand: row (node i, dof a) -> columns (node j, dof b) for every neighbour j, b = 0..dofs-1
    row_len = np.repeat(ndeg * dofs, dofs)
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(row_len, out=rp[1:])
    node_of_row = np.repeat(np.arange(nodes, dtype=np.int64), dofs)
    # for each row, its node's neighbour list repeated per dof block
    starts = nrp[node_of_row]
    idx_in_row = np.arange(rp[-1], dtype=np.int64) - np.repeat(rp[:-1], row_len)
    nbr = ncol[np.repeat(starts, row_len) + idx_in_row // dofs]
    cols = nbr * dofs + idx_in_row % dofs
    rows = np.repeat(np.arange(n, dtype=np.int64), row_len)
    val = value_of(rows, cols, "varied")
    return rp.astype(np.int32), cols.astype(np.int32), val.astype(dtype)


def banded_shell(n, dtype=np.float64, upper_per_row=17, band=3000, run=3, seed=12345, rcm=True):
    """cfg 3: random structurally symmetric banded matrix with a shell-element look: each row
    draws ~upper_per_row upper neighbours inside +-band in clustered runs of `run` consecutive
    columns, symmetrised (-> ~2*upper_per_row+1 per row), then RCM.  Values: multiples of 2^-6."""
    rng = np.random.default_rng(seed)
    nruns = max(1, upper_per_row // run)
    base = np.repeat(np.arange(n, dtype=np.int64), nruns)
    off = rng.integers(1, band, size=base.size, dtype=np.int64)
    r_list, c_list = [], []
    for t in range(run):
        c = base + off + t
        ok = c < n
        r_list.append(base[ok])
        c_list.append(c[ok])
    r = np.concatenate(r_list)
    c = np.concatenate(c_list)
    rows = np.concatenate([r, c, np.arange(n, dtype=np.int64)])
    cols = np.concatenate([c, r, np.arange(n, dtype=np.int64)])
    A = sp.csr_matrix((np.ones(len(rows), dtype=np.int8), (rows, cols)), shape=(n, n))
    A.sum_duplicates()
    A.data[:] = 1
    if rcm:
        p = reverse_cuthill_mckee(A, symmetric_mode=True)
        A = A[p][:, p].tocsr()
    A.sort_indices()
    rp = A.indptr.astype(np.int64)
    cols = A.indices.astype(np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    val = value_of(rows, cols, "varied")
    return rp.astype(np.int32), cols.astype(np.int32), val.astype(dtype)


def rmat(scale, edge_factor=16, dtype=np.float64, abcd=(0.57, 0.19, 0.19, 0.05), seed=12345,
         chunk=1 << 24):
    """cfg 4: R-MAT graph, n = 2^scale, edge_factor*n edge draws, duplicates removed, NOT
    symmetrised.  Values in (-1,1) on the 2^-6 grid, never 0."""
    n = 1 << scale
    m = edge_factor * n
    rng = np.random.default_rng(seed)
    a, b, c, _ = abcd
    keys = []
    for lo in range(0, m, chunk):
        cnt = min(chunk, m - lo)
        r = np.zeros(cnt, dtype=np.int64)
        cidx = np.zeros(cnt, dtype=np.int64)
        for _ in range(scale):
            u = rng.random(cnt)
            rbit = u >= a + b
            cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
            r = (r << 1) | rbit
            cidx = (cidx << 1) | cbit
        keys.append(np.unique((r << scale) | cidx))
    key = np.unique(np.concatenate(keys))
    rows = key >> scale
    cols = key & (n - 1)
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rp[1:])
    q = ((31 * rows + 17 * cols) % 127) - 63          # -63..63
    q = np.where(q == 0, 1, q)
    val = q / 64.0
    return rp.astype(np.int32), cols.astype(np.int32), val.astype(dtype)


def rmat_torch(scale, edge_factor=16, dtype=np.float64, abcd=(0.57, 0.19, 0.19, 0.05), seed=12345, device="cuda",
               chunk=1 << 26):
    """Same construction as rmat() with torch on a device: the scale-24 case takes ~10 s on a GPU
    instead of minutes in numpy.  Same distribution and value rule, a different random stream, so
    the matrix differs from rmat()'s; test data only (torch is not used by the library)."""
    import torch
    n = 1 << scale
    m = edge_factor * n
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a, b, c, _ = abcd
    keys = []
    for lo in range(0, m, chunk):
        cnt = min(chunk, m - lo)
        r = torch.zeros(cnt, dtype=torch.int64, device=device)
        cidx = torch.zeros(cnt, dtype=torch.int64, device=device)
        for _ in range(scale):
            u = torch.rand(cnt, generator=gen, device=device, dtype=torch.float64)
            rbit = (u >= a + b).to(torch.int64)
            cbit = (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64)
            r = (r << 1) | rbit
            cidx = (cidx << 1) | cbit
        keys.append(torch.unique((r << scale) | cidx))
        del r, cidx, u
    key = torch.unique(torch.cat(keys))
    del keys
    rows = key >> scale
    cols = key & (n - 1)
    rp = torch.zeros(n + 1, dtype=torch.int64, device=device)
    rp[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    q = ((31 * rows + 17 * cols) % 127) - 63
    q = torch.where(q == 0, torch.ones_like(q), q)
    val = q.to(torch.float64) / 64.0
    return (rp.to(torch.int32).cpu().numpy(), cols.to(torch.int32).cpu().numpy(), val.cpu().numpy().astype(dtype))


def random_ragged(n, max_len=40, dtype=np.float64, seed=7, empty_frac=0.1, diag=False):
    """Small irregular test matrix: ragged rows, some empty, unsorted-free, unique columns."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, max_len + 1, size=n)
    lens[rng.random(n) < empty_frac] = 0
    lens = np.minimum(lens, n)
    rows, cols = [], []
    for i in range(n):
        c = rng.choice(n, size=lens[i], replace=False)
        if diag and i not in c:
            c = np.append(c, i)
        c.sort()
        rows.append(np.full(len(c), i, dtype=np.int64))
        cols.append(c.astype(np.int64))
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    rp = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rp[1:])
    q = ((31 * rows + 17 * cols) % 127) - 63
    q = np.where(q == 0, 1, q)
    return rp.astype(np.int32), cols.astype(np.int32), (q / 64.0).astype(dtype)


def write_csr_text(path, rp, ci, val, base=0, n_cols=None):
    """`.mtx.rcm.csr` grammar (helpers/converter.m:42-48): "m n nnz", row_ptr, col_ind, val,
    each block space separated with a trailing space on its own line; values "%f".
    base=0 for spmv-csr / cuda-spmv-csr (spmv-csr/spmv.c:37-45); base=1 for spmv-csrk and
    cuda-spmv-csrk (spmv-csrk/spmv.cpp:60,67)."""
    m = len(rp) - 1
    n_cols = m if n_cols is None else n_cols
    with open(path, "w") as f:
        f.write("%d %d %d\n" % (m, n_cols, len(ci)))
        f.write(" ".join(map(str, (np.asarray(rp, dtype=np.int64) + base).tolist())) + " \n")
        f.write(" ".join(map(str, (np.asarray(ci, dtype=np.int64) + base).tolist())) + " \n")
        f.write(" ".join("%f" % v for v in np.asarray(val, dtype=np.float64).tolist()) + " \n")
