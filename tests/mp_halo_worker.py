"""Worker of tests/test_multi_gpu.py (one process per GPU, launched with torch.distributed.run).

The N > 1 data path on hardware: a 27-point stencil with row- and column-dependent values is sharded
over the ranks exactly as bench.py does it, the halo regions of every x buffer start as NaN, x is a
DIFFERENT vector at every step (only the owned slice is written locally; everything else must arrive
through the exchange), and after every step the rank's whole block of y is compared with the oracle
product of that step's x (north_star bound).  Modes:

    csr  / peer          PeerHalo + one fused launch per step (hspmv_spmv_halo): the bench default
    csr  / peer-unfused  same transport, front-end flag waits, interior / boundary launches
    csr  / nccl          batch_isend_irecv (the baseline exchange)
    csr2 / peer-unfused  a Band-k-reordered CSR-2 shard cut at super-row boundaries (hspmv_slice_rows)
    csr3 / nccl          the same for k = 3

Prints one line `MP_HALO_OK ranks=<N> modes=<n> steps=<n>` on rank 0 when everything passed.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import heterogeneous_spmv_b200 as hs  # noqa: E402
from heterogeneous_spmv_b200 import dist as hd  # noqa: E402
from heterogeneous_spmv_b200 import generators as G  # noqa: E402
import oracle as O  # noqa: E402

STEPS = 7


def x_of_step(n, t):
    """a different non-constant vector per step, exact in fp64"""
    i = np.arange(n, dtype=np.int64)
    return 1.0 + ((i * (t + 3) + 11 * t) % 97) / 128.0


def run_mode(fmt, exchange, g, rank, world):
    rp, ci, v = G.stencil27(g, mode="varied")
    n = len(rp) - 1
    k = {"csr": 1, "csr2": 2, "csr3": 3}[fmt]
    full = hs.Matrix.from_csr(rp, ci, v)
    perm = np.arange(n, dtype=np.int64)
    if k > 1:
        sizes, _, _ = hs.auto_params(len(ci), n, hs.ARCH_B200)
        full.build_csrk(k, sizes)
        maps = full.csrk()
        perm = maps["perm"].astype(np.int64)
        align = maps["sr_ptr"] if k == 2 else maps["sr_ptr"][maps["ssr_ptr"]]
        row_lo = [int(t) for t in hs.partition_rows(full.row_ptr(), world, align_ptr=align)]
    else:
        row_lo = [int(t) for t in hs.partition_rows(rp, world)]
    lo, hi = row_lo[rank], row_lo[rank + 1]
    assert hi > lo, "empty shard: the test matrix is too small for %d ranks" % world
    wide = full.slice_rows(lo, hi)
    cmin, cmax = wide.col_range()
    wide.close()
    A = full.slice_rows(lo, hi, cmin, cmax - cmin + 1)
    full.close()
    win, m_loc = cmax - cmin + 1, hi - lo
    ranges = hd.gather_col_ranges((cmin, cmax))
    interior = A.interior_rows(lo - cmin, hi - cmin) if k == 1 else (0, 0)
    stream = torch.cuda.current_stream().cuda_stream
    algo = hs.ALGO_CSR_STREAM if k == 1 else hs.ALGO_CSRK
    x0 = hd.wrap_device_ptr(A.x_ptr(), win, torch.float64)
    y_local = torch.full((m_loc,), float("nan"), dtype=torch.float64, device="cuda")
    x_bufs = [x0]
    peer = fused = None

    def compute(a, b, x_ptr=None):
        if k == 1:
            A.spmv_range(a, b, algo, d_x=x_ptr or x0.data_ptr(), d_y=y_local.data_ptr(), stream=stream)
        else:
            A.spmv(algo, d_x=x_ptr or x0.data_ptr(), d_y=y_local.data_ptr(), stream=stream)

    if exchange.startswith("peer"):
        assert hd.halo_is_symmetric(row_lo, ranges)
        x2_ptr = hs.peer_alloc((win + 16) * 8)
        x_bufs.append(hd.wrap_device_ptr(x2_ptr, win, torch.float64))
        peer = hd.PeerHalo(hs, [A.x_ptr(), x2_ptr], 8, row_lo, ranges, rank)
        if exchange == "peer":
            def fused(a, b, flags, nw, step, timeout, x_ptr):
                A.spmv_halo(a, b, d_x=x_ptr, d_y=y_local.data_ptr(), wait_flags=flags, n_wait=nw, wait_step=step,
                            timed_out=timeout, stream=stream)
    sh = hd.ShardedSpMV(row_lo, ranges, rank, x0, compute, interior, peer=peer, fused=fused)
    for xb in x_bufs:
        xb.fill_(float("nan"))           # halo AND owned slice: nothing may survive from an earlier step by accident
    torch.cuda.synchronize()
    dist.barrier()
    for t in range(1, STEPS + 1):
        xt = x_of_step(n, t)                                   # in the ORIGINAL numbering
        xt_k = xt[perm]                                        # in the kernel's (Band-k) numbering
        if peer is not None:
            peer.reuse_buffer()
            xb = x_bufs[(peer.step + 1) % 2]
        else:
            xb = x0
        xb[lo - cmin: hi - cmin] = torch.from_numpy(xt_k[lo:hi]).cuda()      # the owned slice only
        sh.step(overlap=True)
        torch.cuda.synchronize()
        y = y_local.cpu().numpy()
        y_ref = O.csr_spmv(rp, ci, v, xt)[perm][lo:hi]
        bound = 1e-12 * O.abs_rowsum(rp, ci, v, xt)[perm][lo:hi]
        bad = np.nonzero(~(np.abs(y - y_ref) <= bound))[0]
        assert bad.size == 0, "%s/%s rank %d step %d: %d rows outside the bound, first local row %d (y %r, ref %r)" % (
            fmt, exchange, rank, t, bad.size, bad[0], y[bad[0]], y_ref[bad[0]])
        # the halo really changed hands: my copy of the neighbours' entries equals x_t there
        got = xb.cpu().numpy()
        for _, a, b in sh.recv:
            assert np.array_equal(got[a - cmin: b - cmin], xt_k[a:b]), "%s/%s rank %d step %d: halo [%d,%d) is not x_t" % (
                fmt, exchange, rank, t, a, b)
        y_local.fill_(float("nan"))
    if peer is not None:
        assert not peer.timed_out()
        peer.close()
        hs.peer_free(x2_ptr)
    A.close()
    dist.barrier()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    hs.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = int(os.environ.get("MP_HALO_G", "40"))
    modes = [("csr", "peer"), ("csr", "peer-unfused"), ("csr", "nccl"), ("csr2", "peer-unfused"), ("csr3", "nccl")]
    for fmt, exchange in modes:
        run_mode(fmt, exchange, g, rank, world)
    dist.barrier()
    if rank == 0:
        print("MP_HALO_OK ranks=%d modes=%d steps=%d" % (world, len(modes), STEPS), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
