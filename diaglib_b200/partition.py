"""Row partitioning of the CSR matrix over the ranks of one box and the halo plan of the
built-in SpMM.  Pure numpy host logic: covered on CPU by the world-size-2
gloo tests, consumed on GPU by diaglib_b200_set_csr / diaglib_b200_set_halo.

Rank r owns the contiguous global rows [row_range(n, r, p)).  Its local column space is
[0, n_loc) for owned rows followed by the halo: for every other rank q that owns columns it
references, the contiguous global range [lo_q, hi_q) of q's rows (min..max referenced column
inside q's block), in increasing q.  For banded matrices (3-D stencil in z-slabs, bounded
stride FCI-like) this is a thin neighbour exchange; for an unstructured matrix it degenerates
to (almost) an all-gather, which is the documented worst case.
"""
from __future__ import annotations

import numpy as np


def row_range(n: int, rank: int, size: int):
    """Contiguous block partition; multiples of 2 rows per rank where possible so that local
    leading dimensions stay 16-byte aligned for the vectorised loaders."""
    base = (n // size) & ~1
    r0 = rank * base
    r1 = n if rank == size - 1 else r0 + base
    return r0, r1


def owner_ranges(n: int, size: int):
    return [row_range(n, r, size) for r in range(size)]


def needed_ranges(col, n: int, rank: int, size: int):
    """For each other rank q: the global column range [lo, hi) of q's rows referenced by this
    rank's rows (None when nothing is referenced)."""
    col = np.asarray(col, dtype=np.int64)
    out = []
    for q, (q0, q1) in enumerate(owner_ranges(n, size)):
        if q == rank:
            out.append(None)
            continue
        sel = col[(col >= q0) & (col < q1)]
        out.append(None if sel.size == 0 else (int(sel.min()), int(sel.max()) + 1))
    return out


def union_ranges(*needed_lists):
    """Element-wise union of several `needed_ranges` results (one per matrix sharing a halo):
    the smallest range per peer that covers all of them."""
    out = []
    for per_peer in zip(*needed_lists):
        rgs = [rg for rg in per_peer if rg is not None]
        out.append(None if not rgs else (min(r[0] for r in rgs), max(r[1] for r in rgs)))
    return out


def localize(col, n: int, rank: int, size: int, needed=None):
    """Map global column indices to local ones.  Returns (col_local int32, n_halo, recv) where
    recv = [(q, lo, hi, halo_offset)] describes the halo layout.  Every column must be owned or
    lie inside one of the `needed` ranges (ValueError otherwise: a column outside the halo would
    be gathered from an arbitrary address on the device)."""
    col = np.asarray(col, dtype=np.int64)
    r0, r1 = row_range(n, rank, size)
    n_loc = r1 - r0
    needed = needed_ranges(col, n, rank, size) if needed is None else needed
    out = np.full(col.shape, -1, dtype=np.int64)
    own = (col >= r0) & (col < r1)
    out[own] = col[own] - r0
    off = 0
    recv = []
    for q, rg in enumerate(needed):
        if rg is None:
            continue
        lo, hi = rg
        sel = (col >= lo) & (col < hi) & ~own
        out[sel] = n_loc + off + (col[sel] - lo)
        recv.append((q, lo, hi, off))
        off += hi - lo
    if out.size and out.min() < 0:
        bad = col[out < 0]
        raise ValueError(f"rank {rank}: {bad.size} column indices (e.g. {int(bad[0])}) are neither owned nor inside the "
                         f"halo ranges {needed}; build the ranges from the union of all matrices that share the halo")
    if n_loc + off > np.iinfo(np.int32).max:
        raise ValueError("local column space (owned rows + halo) exceeds int32")
    return out.astype(np.int32), off, recv


def halo_plan(recv, needed_by_all, n: int, rank: int, size: int):
    """Build the exchange plan (peer, send_row0, send_cnt, recv_off, recv_cnt) for this rank.
    needed_by_all[q] is rank q's `needed_ranges` list (so needed_by_all[q][rank] is what q wants
    from us).  One entry per peer that we send to and/or receive from."""
    r0, _ = row_range(n, rank, size)
    recv_map = {q: (lo, hi, off) for (q, lo, hi, off) in recv}
    peers = sorted(set(recv_map) | {q for q in range(size) if q != rank and needed_by_all[q][rank] is not None})
    peer, s0, sc, ro, rc = [], [], [], [], []
    for q in peers:
        peer.append(q)
        want = needed_by_all[q][rank]
        if want is None:
            s0.append(0)
            sc.append(0)
        else:
            s0.append(want[0] - r0)
            sc.append(want[1] - want[0])
        if q in recv_map:
            lo, hi, off = recv_map[q]
            ro.append(off)
            rc.append(hi - lo)
        else:
            ro.append(0)
            rc.append(0)
    return (np.array(peer, np.int32), np.array(s0, np.int64), np.array(sc, np.int64), np.array(ro, np.int64),
            np.array(rc, np.int64))
