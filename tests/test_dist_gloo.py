"""World-size-2 gloo test of the multi-GPU layer's host logic (partition -> halo plan -> exchange ->
interior/boundary split), on CPU.  The product has no CPU SpMV, so the local compute is the
oracle, injected here only to exercise the plan and the exchange."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, g, out_dir, overlap):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as O
    import partition_oracle as PO
    import heterogeneous_spmv_b200 as hs
    from heterogeneous_spmv_b200 import dist as hd
    from heterogeneous_spmv_b200 import generators as G
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rp, ci, v = G.stencil27(g, mode="varied")
        n = len(rp) - 1
        row_lo = hs.partition_rows(rp, world)
        assert list(row_lo) == bench.partition_stencil(g, world)          # closed form == library rule
        assert np.array_equal(row_lo, PO.partition_rows(rp, world))
        lo, hi = int(row_lo[rank]), int(row_lo[rank + 1])
        cmin, cmax = hs.block_col_range(rp, ci, lo, hi)
        half_bw = g * g + g + 1
        assert max(0, lo - half_bw) <= cmin <= lo and hi - 1 <= cmax <= min(n - 1, hi - 1 + half_bw)
        assert (cmin, cmax) == PO.col_range(rp, ci, lo, hi)
        ranges = hd.gather_col_ranges((cmin, cmax))
        assert ranges[rank] == (cmin, cmax)
        assert hd.halo_plan(row_lo, ranges, rank) == PO.halo_plan(row_lo, ranges, rank)
        # local block: rebased row pointers, localised columns
        b_rp = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
        b_ci = (ci[rp[lo]:rp[hi]] - cmin).astype(np.int32)
        b_v = v[rp[lo]:rp[hi]]
        interior = PO.interior_rows(b_rp, b_ci, lo - cmin, hi - cmin)
        x = G.vector_x(n)
        x_local = torch.full((cmax - cmin + 1,), float("nan"), dtype=torch.float64)
        x_local[lo - cmin: hi - cmin] = torch.from_numpy(x[lo:hi])          # only the owned slice is known
        y_local = np.full(hi - lo, np.nan)
        calls = []

        def compute(a, b):
            calls.append((a, b))
            sub_rp = (b_rp[a:b + 1] - b_rp[a]).astype(np.int32)
            sub = slice(b_rp[a], b_rp[b])
            xs = x_local.numpy()
            if (a, b) == interior and overlap:
                # interior rows must not touch the halo: poison check
                used = b_ci[sub]
                assert used.min() >= lo - cmin and used.max() < hi - cmin
            y_local[a:b] = O.csr_spmv(sub_rp, b_ci[sub], b_v[sub], xs)

        sh = hd.ShardedSpMV(row_lo, ranges, rank, x_local, compute, interior)
        sh.step(overlap=overlap)
        assert not torch.isnan(x_local).any(), "halo not filled"
        assert np.array_equal(x_local.numpy(), x[cmin:cmax + 1])
        y_ref = O.csr_spmv(rp, ci, v, x)[lo:hi]
        assert np.array_equal(y_local, y_ref)
        if overlap and interior[1] > interior[0]:
            assert calls[0] == tuple(interior) and sorted(calls) == sorted(
                [t for t in [(0, interior[0]), tuple(interior), (interior[1], hi - lo)] if t[1] > t[0]])
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("%d %d %d" % (lo, hi, sh.halo_elems))
    finally:
        dist.destroy_process_group()


def _worker_replicated(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as O
    import partition_oracle as PO
    import heterogeneous_spmv_b200 as hs
    from heterogeneous_spmv_b200 import dist as hd
    from heterogeneous_spmv_b200 import generators as G
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rp, ci, v = G.rmat(12, 8)                      # power-law: every block's window is all of x
        n = len(rp) - 1
        row_lo = hs.partition_rows(rp, world)
        assert np.array_equal(row_lo, PO.partition_rows(rp, world))
        lo, hi = int(row_lo[rank]), int(row_lo[rank + 1])
        nnz_blocks = [int(rp[row_lo[q + 1]] - rp[row_lo[q]]) for q in range(world)]
        assert max(nnz_blocks) - min(nnz_blocks) <= int(np.diff(rp).max())       # nnz-balanced, not row-balanced
        b_rp = (rp[lo:hi + 1] - rp[lo]).astype(np.int32)
        b_ci, b_v = ci[rp[lo]:rp[hi]], v[rp[lo]:rp[hi]]
        x_full = torch.from_numpy(G.vector_x(n).copy())
        y_local = torch.full((hi - lo,), float("nan"), dtype=torch.float64)

        def compute():
            y_local[:] = torch.from_numpy(O.csr_spmv(b_rp, b_ci, b_v, x_full.numpy()))

        sh = hd.ReplicatedSpMV(row_lo, rank, compute)
        x0 = x_full.numpy().copy()
        sh.step()
        y1 = O.csr_spmv(rp, ci, v, x0)
        assert np.array_equal(y_local.numpy(), y1[lo:hi])
        sh.all_gather_y(y_local, x_full)               # y feeds the next x
        assert np.array_equal(x_full.numpy(), y1)
        sh.step()
        assert np.array_equal(y_local.numpy(), O.csr_spmv(rp, ci, v, y1)[lo:hi])
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_replicated_x_for_power_law(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker_replicated, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(str(tmp_path / "ok0")) and os.path.exists(str(tmp_path / "ok1"))


@pytest.mark.parametrize("overlap", [True, False])
def test_two_rank_halo_exchange(tmp_path, overlap):
    port = 29600 + (os.getpid() % 200) + (1 if overlap else 0)
    mp.spawn(_worker, args=(2, port, 12, str(tmp_path), overlap), nprocs=2, join=True)
    got = [open(str(tmp_path / ("ok%d" % r))).read().split() for r in range(2)]
    assert int(got[0][1]) == int(got[1][0])                      # contiguous blocks
    assert 12 * 12 <= int(got[0][2]) <= 12 * 12 + 12 + 1 and 12 * 12 <= int(got[1][2]) <= 12 * 12 + 12 + 1   # halo ~ half bandwidth


def test_partition_closed_form_matches_library():
    sys.path.insert(0, ROOT)
    import bench
    import heterogeneous_spmv_b200 as hs
    from heterogeneous_spmv_b200 import generators as G
    for g in (5, 9, 16):
        rp = G.stencil27_row_ptr(g)
        assert [bench.stencil27_before(i, g) for i in range(0, g ** 3 + 1, 7)] == rp[::7].tolist()
        for parts in (1, 2, 3, 4, 8):
            assert bench.partition_stencil(g, parts) == hs.partition_rows(rp.astype(np.int32), parts).tolist()
    # BASELINE config 5 numbers
    assert bench.stencil27_before(400 ** 3, 400) == 1719374392
    lo = bench.partition_stencil(400, 8)
    nnz = [bench.stencil27_before(lo[i + 1], 400) - bench.stencil27_before(lo[i], 400) for i in range(8)]
    assert max(nnz) - min(nnz) <= 2 * 27


def test_three_rank_halo_exchange_middle_rank_has_two_neighbours(tmp_path):
    """world size 3: the middle rank receives a halo on both sides and pushes to both neighbours; the outer ranks
    exchange with the middle one only (same assertions as the two-rank worker: plan == oracle plan, halo filled
    exactly, y == whole-matrix product, interior rows never touch the halo)."""
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker, args=(3, port, 13, str(tmp_path), True), nprocs=3, join=True)
    got = [open(str(tmp_path / ("ok%d" % r))).read().split() for r in range(3)]
    assert int(got[0][1]) == int(got[1][0]) and int(got[1][1]) == int(got[2][0]) and int(got[2][1]) == 13 ** 3
    half_bw = 13 * 13 + 13 + 1
    assert 13 * 13 <= int(got[0][2]) <= half_bw and 13 * 13 <= int(got[2][2]) <= half_bw
    assert 2 * 13 * 13 <= int(got[1][2]) <= 2 * half_bw                  # both sides
