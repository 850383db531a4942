"""One-box multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

The reference has no multi-GPU code; this is the new layer north_star asks for:
rows are split into contiguous nnz-balanced blocks (hspmv_partition_rows), every rank keeps only
its block with row pointers rebased and column ids localised to the window [cmin,cmax] of x it
references, and before each local SpMV the parts of that window owned by other ranks (the halo)
are fetched with NCCL send/recv over NVLink.  Rows that reference no halo column (the interior,
hspmv_interior_rows) are launched before the halo has arrived.  Power-law matrices, whose window
is all of x, use replicated x instead (no per-step traffic while x is static).

Nothing here computes on the CPU: `compute` defaults to the library's CUDA kernels; the CPU
(gloo) tests inject the oracle only to exercise the plan and the exchange.
"""
import torch
import torch.distributed as dist


def halo_plan(row_lo, col_ranges, rank):
    """(recv, send): lists of (peer, global_start, global_end).  recv = pieces of this rank's
    column window owned by peers; send = pieces of this rank's rows inside peers' windows."""
    P = len(row_lo) - 1
    lo, hi = int(row_lo[rank]), int(row_lo[rank + 1])
    cmin, cmax = (int(t) for t in col_ranges[rank])
    recv, send = [], []
    for q in range(P):
        if q == rank:
            continue
        qlo, qhi = int(row_lo[q]), int(row_lo[q + 1])
        a, b = max(cmin, qlo), min(cmax + 1, qhi)
        if a < b:
            recv.append((q, a, b))
        qmin, qmax = (int(t) for t in col_ranges[q])
        a, b = max(qmin, lo), min(qmax + 1, hi)
        if a < b:
            send.append((q, a, b))
    return recv, send


def gather_col_ranges(col_range, group=None):
    """all_gather of every rank's (cmin, cmax)."""
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor(list(col_range), dtype=torch.int64, device=dev)
    out = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [tuple(int(v) for v in t.tolist()) for t in out]


class ShardedSpMV:
    """y_block = A_block * x for one rank's row block.

    x_local : 1-D tensor over the global columns [cmin, cmax]; the slice owned by this rank
              (rows [lo,hi) of a square matrix) sits at offset lo - cmin.
    compute(row_a, row_b) : enqueues the local SpMV of block rows [row_a,row_b) on the current
              stream (GPU) or runs it (CPU tests).
    """

    def __init__(self, row_lo, col_ranges, rank, x_local, compute, interior, group=None):
        self.rank = rank
        self.group = group
        self.lo, self.hi = int(row_lo[rank]), int(row_lo[rank + 1])
        self.cmin, self.cmax = (int(t) for t in col_ranges[rank])
        assert x_local.numel() == self.cmax - self.cmin + 1
        self.x_local = x_local
        self.compute = compute
        self.m = self.hi - self.lo
        self.row_a, self.row_b = interior            # interior rows need no halo
        self.recv, self.send = halo_plan(row_lo, col_ranges, rank)
        self.halo_elems = sum(b - a for _, a, b in self.recv)

    def owned(self):
        return self.x_local[self.lo - self.cmin: self.hi - self.cmin]

    def post_exchange(self):
        ops = []
        for q, a, b in self.send:
            ops.append(dist.P2POp(dist.isend, self.x_local[a - self.cmin: b - self.cmin], q, group=self.group))
        for q, a, b in self.recv:
            ops.append(dist.P2POp(dist.irecv, self.x_local[a - self.cmin: b - self.cmin], q, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def step(self, overlap=True):
        """halo exchange + local SpMV.  With overlap, interior rows run while the halo is in flight."""
        works = self.post_exchange()
        if overlap and self.row_b > self.row_a:
            self.compute(self.row_a, self.row_b)
            for w in works:
                w.wait()
            if self.row_a > 0:
                self.compute(0, self.row_a)
            if self.row_b < self.m:
                self.compute(self.row_b, self.m)
        else:
            for w in works:
                w.wait()
            self.compute(0, self.m)
