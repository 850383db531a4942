"""One-box multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

The reference has no multi-GPU code; this is the new layer north_star asks for:
rows are split into contiguous nnz-balanced blocks (hspmv_partition_rows), every rank keeps only
its block with row pointers rebased and column ids localised to the window [cmin,cmax] of x it
references, and before each local SpMV the parts of that window owned by other ranks (the halo)
arrive over NVLink: copy-engine writes into IPC-mapped neighbour buffers (PeerHalo, the default)
or NCCL send/recv (the baseline).  Rows that reference no halo column (the interior,
hspmv_interior_rows) run before the halo has arrived.  Power-law matrices, whose window is all
of x, use replicated x instead (ReplicatedSpMV: no per-step traffic while x is static).

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


def wrap_device_ptr(ptr, count, dtype=torch.float64):
    """Zero-copy torch view of device memory owned by the library (CUDA array interface)."""
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4"}[dtype]

    class _Holder:
        __cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 3}

    return torch.as_tensor(_Holder(), device="cuda")


class AsymmetricHalo(RuntimeError):
    """the ack-less ping-pong protocol needs senders == receivers on every rank"""


def halo_is_symmetric(row_lo, col_ranges):
    """True when, on every rank, the set of ranks it receives from equals the set it sends to."""
    for r in range(len(row_lo) - 1):
        recv, send = halo_plan(row_lo, col_ranges, r)
        if {q for q, _, _ in recv} != {q for q, _, _ in send}:
            return False
    return True


class PeerHalo:
    """Halo exchange straight into the neighbours' x buffers over NVLink (CUDA IPC), no NCCL and
    no SM: the local SpMV kernel is persistent and owns every SM, so everything that must run
    NEXT to it is a copy-engine or command-processor operation.

    x is ping-ponged between two buffers (as an iterative solver does with x_k / x_{k+1}); step t
    uses buffer t % 2.  Per step, all stream-ordered, no host synchronisation:
      main stream : record "x of step t is ready" -> ONE kernel (hspmv_spmv_halo): interior tiles,
                    then each warp spins (bounded) on my data flags until they reach t, then the
                    boundary tiles
      comm stream : wait for that event -> peer copies of my boundary slices of x[t%2] into the
                    neighbours' halo regions of THEIR x[t%2] -> write t into their data flags
                    (cuStreamWriteValue32)
    No acknowledgement is needed: I overwrite a neighbour's halo of buffer b at step t+2 only
    after my kernel of step t+1 has finished, and that kernel waited for the neighbour's flags of
    step t+1, which the neighbour sent after ITS kernel of step t (the last reader of buffer b)
    had finished.  That argument needs every rank I push to to be one of my senders as well
    (otherwise nothing makes me wait for it): the constructor checks it on the all-gathered plans
    and raises AsymmetricHalo, and the caller falls back to an exchange with acknowledgements
    (NCCL send/recv).  begin() also makes the main stream wait for my own pushes of step t-2, the
    last readers of the buffer that the caller is about to fill with x of step t.  (An earlier version sent acks and waited for them with cuStreamWaitValue32 on
    the comm stream: a command-processor wait on a word written from another GPU woke up 0.2-0.5 ms
    late while the persistent kernel was running -- profiles/r1_scaling.md.)
    The buffers must come from cudaMalloc (hspmv_x_device / hspmv_peer_alloc).
    Flag words (int32) per rank: [0, n_senders) data flags in sorted-sender order, [32] timeout
    indicator, [33..35] spin diagnostics."""

    def __init__(self, hs, x_ptrs, itemsize, row_lo, col_ranges, rank, group=None):
        self.hs = hs
        self.rank = rank
        self.P = P = len(row_lo) - 1
        self.itemsize = itemsize
        self.x_ptrs = list(x_ptrs)
        assert len(self.x_ptrs) == 2
        self.cmin = int(col_ranges[rank][0])
        self.recv, self.send = halo_plan(row_lo, col_ranges, rank)
        self.senders = sorted({q for q, _, _ in self.recv})
        self.receivers = sorted({q for q, _, _ in self.send})
        if not halo_is_symmetric(row_lo, col_ranges):      # same answer on every rank: all raise together
            raise AsymmetricHalo("rank %d receives from %s but sends to %s" % (rank, self.senders, self.receivers))
        assert len(self.senders) <= 32
        self.flags = hs.peer_alloc(4 * 64)
        blob = hs.peer_export(self.x_ptrs[0]) + hs.peer_export(self.x_ptrs[1]) + hs.peer_export(self.flags)
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        allh = [torch.zeros(192, dtype=torch.uint8, device="cuda") for _ in range(P)]
        dist.all_gather(allh, mine, group=group)
        senders_t = torch.zeros(P, dtype=torch.int64, device="cuda")
        for q in self.senders:
            senders_t[q] = 1
        all_senders = [torch.zeros(P, dtype=torch.int64, device="cuda") for _ in range(P)]
        dist.all_gather(all_senders, senders_t, group=group)
        self.peer_x, self.peer_flags, self.my_slot_at = {}, {}, {}
        for q in self.receivers:        # only the ranks I write to need to be mapped
            raw = bytes(allh[q].cpu().numpy().tobytes())
            self.peer_x[q] = (hs.peer_open(raw[:64]), hs.peer_open(raw[64:128]))
            self.peer_flags[q] = hs.peer_open(raw[128:192])
            their_senders = [i for i in range(P) if int(all_senders[q][i])]
            self.my_slot_at[q] = their_senders.index(rank)
        self.peer_cmin = [int(c[0]) for c in col_ranges]
        self.step = 0
        self.comm = torch.cuda.Stream()
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.pushed = [None, None]      # comm-stream events: my pushes out of buffer b are complete
        self.trace = None   # set to [] to collect (kernel-start, push-done) timing events per step
        self.halo_elems = sum(b - a for _, a, b in self.recv)
        dist.barrier(group=group)

    def _w(self, base_ptr, word):
        return base_ptr + 4 * word

    def x_ptr(self, step=None):
        """Device pointer of the x buffer that step `step` (default: the current one) reads."""
        return self.x_ptrs[(self.step if step is None else step) % 2]

    def wait_words(self):
        """(device pointer, count) of the data-flag words the fused kernel polls."""
        return self.flags, len(self.senders)

    def timeout_word(self):
        return self._w(self.flags, 32)

    def reuse_buffer(self):
        """Call BEFORE writing x of the next step into its buffer: the main stream waits until my
        own pushes of two steps ago, which read that buffer, are complete."""
        ev = self.pushed[(self.step + 1) % 2]
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def begin(self):
        """Step t starts: x[t % 2] is ready at this point of the main stream."""
        self.step += 1
        self.ready[self.step % 2].record(torch.cuda.current_stream())
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            self.trace.append([e, None])

    def post(self):
        """Comm stream of step t: my slices go to the neighbours, then their flags."""
        hs, t = self.hs, self.step
        b = t % 2
        self.comm.wait_event(self.ready[b])
        cs = self.comm.cuda_stream
        for q, lo, hi in self.send:      # global columns [lo,hi) of mine -> q's window
            hs.memcpy_async(self.peer_x[q][b] + (lo - self.peer_cmin[q]) * self.itemsize,
                            self.x_ptrs[b] + (lo - self.cmin) * self.itemsize, (hi - lo) * self.itemsize, cs)
        for q in self.receivers:
            hs.stream_write(cs, self._w(self.peer_flags[q], self.my_slot_at[q]), t)
        if self.pushed[b] is None:
            self.pushed[b] = torch.cuda.Event()
        self.pushed[b].record(self.comm)
        if self.trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(self.comm)
            self.trace[-1][1] = e

    def push_delays_ms(self):
        """per traced step: time from the kernel's launch point on the main stream to the
        completion of this rank's pushes on the comm stream"""
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in (self.trace or []) if b is not None]

    def wait_data(self):
        """Unfused variant: make the main stream wait for this step's halo (front-end wait)."""
        ms = torch.cuda.current_stream().cuda_stream
        for i in range(len(self.senders)):
            self.hs.stream_wait_ge(ms, self._w(self.flags, i), self.step)

    def timed_out(self):
        flag = wrap_device_ptr(self.timeout_word(), 1, torch.int32)
        return bool(int(flag.item()))

    def spin_stats(self):
        """(longest warp spin us, mean warp spin us, histogram) at 1.965 GHz since the last reset;
        histogram buckets: < 4, < 17, < 67, < 267, < 1067 us, longer."""
        w = wrap_device_ptr(self.timeout_word(), 10, torch.int32).cpu().tolist()
        mx, cnt, tot = w[1], max(w[2], 1), w[3]
        return mx * 1024 / 1965.0, tot * 1024 / 1965.0 / cnt, w[4:10]

    def spin_reset(self):
        torch.cuda.synchronize()
        wrap_device_ptr(self._w(self.flags, 33), 9, torch.int32).zero_()
        torch.cuda.synchronize()

    def close(self):
        torch.cuda.synchronize()
        for q in list(self.peer_x):
            for ptr in self.peer_x[q]:
                self.hs.peer_close(ptr)
            self.hs.peer_close(self.peer_flags[q])
        self.hs.peer_free(self.flags)
        self.peer_x, self.peer_flags = {}, {}


class ShardedSpMV:
    """y_block = A_block * x for one rank's row block.

    x_local : 1-D tensor over the global columns [cmin, cmax]; the slice owned by this rank
              (rows [lo,hi) of a square matrix) sits at offset lo - cmin.
    compute(row_a, row_b) : enqueues the local SpMV of block rows [row_a,row_b) on the current
              stream (GPU) or runs it (CPU tests).
    """

    def __init__(self, row_lo, col_ranges, rank, x_local, compute, interior, group=None, peer=None, fused=None):
        self.peer = peer     # PeerHalo: NVLink peer copies + flags instead of NCCL send/recv
        self.fused = fused   # fused(interior_lo, interior_hi, flags_ptr, n_flags, step, timeout_ptr, x_ptr): one launch per step
        # (with a PeerHalo, compute(row_a, row_b, x_ptr) receives the x buffer of the step)
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
        if self.peer is not None and self.fused is not None:
            # the kernel goes first: it waits for the halo by itself, and whatever the host does
            # while enqueueing the exchange then happens while the kernel already runs
            self.peer.begin()
            flags, n = self.peer.wait_words()
            self.fused(self.row_a, self.row_b, flags, n, self.peer.step, self.peer.timeout_word(), self.peer.x_ptr())
            self.peer.post()
            return
        if self.peer is not None:
            self.peer.begin()
            self.peer.post()
            xp = self.peer.x_ptr()
            if self.row_b > self.row_a:
                self.compute(self.row_a, self.row_b, xp)
            self.peer.wait_data()
            if self.row_b > self.row_a:
                if self.row_a > 0:
                    self.compute(0, self.row_a, xp)
                if self.row_b < self.m:
                    self.compute(self.row_b, self.m, xp)
            else:
                self.compute(0, self.m, xp)
            return
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


class ReplicatedSpMV:
    """Power-law matrices (BASELINE config 4): a row block references columns all over x, so every
    rank keeps a full replica of x, y is row-sharded, and a step is just the local SpMV -- no
    collective while x is static.  When y feeds the next x (power iteration),
    all_gather_y() rebuilds the replica: one all_gather of the padded row blocks.

    compute() : enqueues / runs the local SpMV of the whole block against the replica.
    """

    def __init__(self, row_lo, rank, compute, group=None):
        self.row_lo = [int(t) for t in row_lo]
        self.rank = rank
        self.group = group
        self.compute = compute
        self.lo, self.hi = self.row_lo[rank], self.row_lo[rank + 1]
        self.m = self.hi - self.lo
        self.pad = max(b - a for a, b in zip(self.row_lo[:-1], self.row_lo[1:]))

    def step(self):
        self.compute()

    def all_gather_y(self, y_local, x_full):
        """x_full[row_lo[q]:row_lo[q+1]] = rank q's y block, for every q"""
        world = len(self.row_lo) - 1
        mine = torch.zeros(self.pad, dtype=y_local.dtype, device=y_local.device)
        mine[: self.m] = y_local
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=self.group)
        for q in range(world):
            a, b = self.row_lo[q], self.row_lo[q + 1]
            x_full[a:b] = parts[q][: b - a]
