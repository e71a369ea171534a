"""CPU tests of the N>1 host logic with world_size=2 over gloo: row partition, column
localisation, halo plan and the exchange it prescribes (send/recv exactly as the library's
ncclSend/ncclRecv sequence), and the all-reduce of partial Gram matrices."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diaglib_b200 import partition, problems as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gen, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if gen == "lap3d":
            n = 16 * 16 * 8
            rows = lambda a, b: P.lap3d(16, 16, 8, a, b, delta=0.5)  # noqa: E731
            full = P.lap3d(16, 16, 8, delta=0.5)
        else:
            n = 2048
            rows = lambda a, b: P.fci_like(n, a, b, n_strides=8, bandwidth=300)  # noqa: E731
            full = P.fci_like(n, n_strides=8, bandwidth=300)
        m = 5
        r0, r1 = partition.row_range(n, rank, world)
        rowptr, col, val, diag = rows(r0, r1)
        needed = partition.needed_ranges(col, n, rank, world)
        all_needed = [None] * world
        dist.all_gather_object(all_needed, needed)
        col_loc, n_halo, recv = partition.localize(col, n, rank, world, needed)
        peer, s0, sc, ro, rc = partition.halo_plan(recv, all_needed, n, rank, world)
        x_glob = P.guess(n, m)
        x = np.ascontiguousarray(x_glob[r0:r1])
        halo = np.zeros((n_halo, m))
        # the exchange of Engine::halo_exchange: pack rows, send/recv per peer, unpack at recv_off
        reqs, bufs = [], []
        for i, q in enumerate(peer):
            if sc[i] > 0:
                t = torch.from_numpy(np.ascontiguousarray(x[s0[i]:s0[i] + sc[i]]))
                reqs.append(dist.isend(t, int(q)))
                bufs.append(t)
            if rc[i] > 0:
                t = torch.empty((int(rc[i]), m), dtype=torch.float64)
                reqs.append(dist.irecv(t, int(q)))
                bufs.append((i, t))
        for r in reqs:
            r.wait()
        for b in bufs:
            if isinstance(b, tuple):
                i, t = b
                halo[ro[i]:ro[i] + rc[i]] = t.numpy()
        xe = np.vstack([x, halo])
        a_loc = sp.csr_matrix((val, col_loc, rowptr), shape=(r1 - r0, r1 - r0 + n_halo))
        ax = a_loc @ xe
        a_full = sp.csr_matrix((full[2], full[1], full[0]), shape=(n, n))
        err_spmm = float(np.abs(ax - (a_full @ x_glob)[r0:r1]).max())
        # metric of the generalized problem: localised with the matrix's halo numbering
        # (dist.install_partitioned(metric_rows=...)), applied to the same extended block
        b_rowptr, b_col, b_val = P.metric_like((rowptr, col, val, diag), r0=r0)
        b_loc, b_halo, _ = partition.localize(b_col, n, rank, world, needed)
        b_full = P.metric_like(full)
        bx = sp.csr_matrix((b_val, b_loc, b_rowptr), shape=(r1 - r0, r1 - r0 + n_halo)) @ xe
        bx_ref = (sp.csr_matrix((b_full[2], b_full[1], b_full[0]), shape=(n, n)) @ x_glob)[r0:r1]
        err_metric = float(np.abs(bx - bx_ref).max()) + (0.0 if b_halo == n_halo else 1.0)
        # Gram all-reduce: sum of the per-rank partial X^T (A X) equals the global one
        g = torch.from_numpy(x.T @ ax)
        dist.all_reduce(g)
        gref = x_glob.T @ (a_full @ x_glob)
        err_gram = float(np.abs(g.numpy() - gref).max() / np.abs(gref).max())
        # norms: sum-of-squares all-reduce(sum) and max all-reduce(max), as in the drivers
        ss = torch.from_numpy((ax * ax).sum(axis=0))
        mx = torch.from_numpy(np.abs(ax).max(axis=0))
        dist.all_reduce(ss)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ref = a_full @ x_glob
        err_norm = float(max((np.abs(ss.numpy() - (ref * ref).sum(axis=0)) / (ref * ref).sum(axis=0)).max(),
                             np.abs(mx.numpy() - np.abs(ref).max(axis=0)).max()))
        out_q.put((rank, err_spmm, err_gram, err_norm, int(n_halo), len(peer), err_metric))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gen", ["lap3d", "fci_like"])
def test_partitioned_spmm_and_reductions_world2(gen):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gen, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, e3, n_halo, n_peer, e4 in res:
        assert e1 < 1e-12 and e2 < 1e-13 and e3 < 1e-13 and e4 < 1e-13
        assert n_halo > 0 and n_peer == 1
    if gen == "lap3d":  # z-slabs: the halo is exactly one 16x16 plane
        assert all(r[4] == 256 for r in res)


def test_row_range_covers_everything():
    for n in (10, 1000, 4097, 1 << 20):
        for size in (1, 2, 3, 4, 8):
            rr = partition.owner_ranges(n, size)
            assert rr[0][0] == 0 and rr[-1][1] == n
            assert all(rr[i][1] == rr[i + 1][0] for i in range(size - 1))
            assert all((b - a) % 2 == 0 for a, b in rr[:-1])
