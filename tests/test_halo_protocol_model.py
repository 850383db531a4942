"""Executable form of the safety argument of dist.PeerHalo (no acknowledgements, two x buffers).

A discrete-event model of the per-step protocol, with random and adversarial latencies:

  main stream of rank r :  K(r,1) K(r,2) ...        one fused kernel per step, in stream order
  comm stream of rank r :  P(r,1) P(r,2) ...        P(r,t) waits for the event recorded before K(r,t),
                                                    i.e. for the end of K(r,t-1), and for P(r,t-1)
  P(r,t) writes the halo of buffer t % B of each neighbour q, then sets flag_q[r] = t
  K(q,t) reads its halo of buffer t % B from the moment all its senders' flags show t until it ends

Invariant: while K(q,t) reads buffer t % B, nobody writes that buffer except with step-t data, and
every step-t write has finished before the read starts.  With B = 2 it must hold for any timing;
with B = 1 (what an ack-less single-buffer scheme would be) the model must find a violation --
the negative control that shows the model can see the race at all.
"""
import random

import pytest


def simulate(P, T, B, rng, skew, one_sided=False):
    """returns the list of violations (reader rank, step, writer rank, writer step).
    one_sided: rank r only receives from r-1 (a structurally unsymmetric matrix): the writer r-1 never
    waits for anything from its reader, which is what dist.halo_is_symmetric() refuses."""
    nbrs = {r: [q for q in ((r - 1,) if one_sided else (r - 1, r + 1)) if 0 <= q < P] for r in range(P)}
    k_end = {(r, 0): 0.0 for r in range(P)}        # K(r,0): nothing
    p_end = {(r, 0): 0.0 for r in range(P)}
    p_start, read_start = {}, {}
    speed = [1.0 + skew * rng.random() for _ in range(P)]
    for t in range(1, T + 1):
        # pushes of step t can be computed first: they depend only on step t-1 of their own rank
        for r in range(P):
            p_start[(r, t)] = max(k_end[(r, t - 1)], p_end[(r, t - 1)]) + rng.random() * 0.05
            p_end[(r, t)] = p_start[(r, t)] + 0.01 + rng.random() * 0.2 * speed[r]
        for r in range(P):
            start = k_end[(r, t - 1)] + rng.random() * 0.02
            interior_done = start + speed[r] * (0.8 + 0.4 * rng.random())
            flags_in = max([p_end[(q, t)] for q in nbrs[r]] + [0.0])
            read_start[(r, t)] = max(interior_done, flags_in)          # the warps spin until the flags show t
            k_end[(r, t)] = read_start[(r, t)] + speed[r] * 0.05
    bad = []
    for q in range(P):
        for t in range(1, T + 1):
            for r in nbrs[q]:
                for tw in range(1, T + 1):
                    if tw % B != t % B:
                        continue                     # other buffer
                    ws, we = p_start[(r, tw)], p_end[(r, tw)]
                    rs, re = read_start[(q, t)], k_end[(q, t)]
                    if tw == t:
                        if we > rs + 1e-12:
                            bad.append((q, t, r, tw, "step-t data not complete when the read starts"))
                    elif ws < re and we > rs:
                        bad.append((q, t, r, tw, "foreign step written during the read"))
                    elif tw < t and we > rs:
                        bad.append((q, t, r, tw, "stale write lands after the read began"))
    return bad


@pytest.mark.parametrize("P", [2, 3, 8])
@pytest.mark.parametrize("skew", [0.0, 0.5, 5.0, 50.0])
def test_ping_pong_without_acks_is_race_free(P, skew):
    for seed in range(40):
        assert simulate(P, 12, 2, random.Random(seed * 7919 + P), skew) == []


def test_single_buffer_without_acks_races_in_the_model():
    hits = sum(bool(simulate(4, 12, 1, random.Random(seed), 5.0)) for seed in range(40))
    assert hits > 0, "the model cannot see the race it is supposed to rule out"


def test_one_sided_senders_race_even_with_two_buffers():
    """negative control for ADVICE r1: with senders != receivers the ack-less ping-pong is NOT safe (a fast
    one-sided sender laps its reader); PeerHalo refuses such plans (dist.halo_is_symmetric)."""
    hits = sum(bool(simulate(4, 12, 2, random.Random(seed), 50.0, one_sided=True)) for seed in range(40))
    assert hits > 0


def test_halo_is_symmetric_accepts_banded_and_rejects_one_sided():
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from heterogeneous_spmv_b200 import dist as hd
    row_lo = [0, 100, 200, 300]
    banded = [(0, 109), (90, 209), (190, 299)]
    assert hd.halo_is_symmetric(row_lo, banded)
    lower_only = [(0, 99), (90, 199), (190, 299)]          # every rank needs only columns below its block
    assert not hd.halo_is_symmetric(row_lo, lower_only)
