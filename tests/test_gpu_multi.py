"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the row-partitioned drivers with NCCL
all-reduce of the k x k matrices and the SpMM halo exchange must reproduce the single-rank
oracle: eigenvalues to 1e-10, iteration count within +-1, residuals below tolerance."""
import os
import socket

import numpy as np
import pytest

from diaglib_b200 import problems as P

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    import torch
    import torch.distributed as dist

    import diaglib_b200 as D
    from diaglib_b200 import dist as DD, partition

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # bootstrap only
    try:
        D.init(rank)
        DD.init_comm(dist)
        if case == "lap3d":
            n, n_targ = 32 * 32 * 16, 6
            gen = lambda a, b: P.lap3d(32, 32, 16, a, b, delta=1.0)  # noqa: E731
            diag = P.lap3d_diag(np.arange(n), 14, 1.0, 1)
            n_max = P.n_eig_rule(n_targ)
            r0, r1 = partition.row_range(n, rank, world)
            guess = P.guess_lowest_diag(diag, n_max, r0, r1) + P.guess(n, n_max, r0, r1) * (0.03 / np.sqrt(n / 12.0))
        else:
            n, n_targ = 1 << 14, 8
            gen = lambda a, b: P.toy_sparse(n, a, b)  # noqa: E731
            n_max = P.n_eig_rule(n_targ)
            r0, r1 = partition.row_range(n, rank, world)
            guess = P.guess(n, n_max, r0, r1)
        gen_eig = case.endswith("gen_eig")
        metric = (lambda a, b: P.metric_like(gen(a, b), r0=a)) if gen_eig else None  # noqa: E731
        DD.install_partitioned(gen, n, rank, world, dist, metric_rows=metric)
        ev = np.asfortranarray(guess)
        eig = np.zeros(n_max)
        if gen_eig:
            ok = D.lobpcg_driver(False, True, r1 - r0, n_targ, n_max, 300, 1e-8, 0.0, None, None, None, eig, ev)
        elif case.endswith("davidson"):
            ok = D.davidson_driver(False, r1 - r0, n_targ, n_max, 300, 1e-8, 12, 0.0, None, None, eig, ev)
        else:
            ok = D.lobpcg_driver(False, False, r1 - r0, n_targ, n_max, 300, 1e-8, 0.0, None, None, None, eig, ev)
        its = len(D.last_history(n_max)["it"])
        q.put((rank, ok, its, eig.copy(), r0, r1, ev.copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["lap3d", "toy_sparse", "toy_sparse_davidson", "toy_sparse_gen_eig"])
def test_two_rank_parity(oracle, case):
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank oracle on the same global problem
    if case == "lap3d":
        n, n_targ = 32 * 32 * 16, 6
        csr = P.lap3d(32, 32, 16, delta=1.0)
        n_max = P.n_eig_rule(n_targ)
        g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (0.03 / np.sqrt(n / 12.0)))
    else:
        n, n_targ = 1 << 14, 8
        csr = P.toy_sparse(n)
        n_max = P.n_eig_rule(n_targ)
        g = P.guess(n, n_max)
    oracle.set_csr(*csr)
    gen_eig = case.endswith("gen_eig")
    if gen_eig:
        bcsr = P.metric_like(csr)
        oracle.set_csr_b(*bcsr)
        ro = oracle.lobpcg(g, n_targ, 300, 1e-8, gen_eig=True)
    else:
        ro = oracle.davidson(g, n_targ, 300, 1e-8, 12) if case.endswith("davidson") else oracle.lobpcg(g, n_targ, 300, 1e-8)
    evec = np.vstack([r[6] for r in res])
    for rank, ok, its, eig, r0, r1, _ in res:
        assert ok and ro["ok"]
        assert np.abs(eig[:n_targ] - ro["eig"][:n_targ]).max() / np.abs(ro["eig"][:n_targ]).max() < 1e-10
        assert abs(its - len(ro["it"])) <= 1
        assert np.array_equal(eig, res[0][3])  # replicated small solves: bit-identical on all ranks
    import scipy.sparse as sp
    a = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(n, n))
    x = evec[:, :n_targ]
    bx = sp.csr_matrix((bcsr[2], bcsr[1], bcsr[0]), shape=(n, n)) @ x if gen_eig else x
    resid = a @ x - bx * res[0][3][:n_targ]
    assert (np.linalg.norm(resid, axis=0) / np.sqrt(n)).max() < 2e-8
    assert np.abs(x.T @ bx - np.eye(n_targ)).max() < (1e-10 if gen_eig else 1e-11)


def _spmm_worker(rank, world, port, m, tiled, q):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import diaglib_b200 as D
    from diaglib_b200 import dist as DD, kernels as K, partition

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D.init(rank)
        DD.init_comm(dist)
        n = 32 * 32 * 16
        DD.install_partitioned(lambda a, b: P.lap3d(32, 32, 16, a, b, delta=1.0), n, rank, world, dist)
        r0, r1 = partition.row_range(n, rank, world)
        if tiled:   # grid tiles along a z-order curve inside this rank's slab of planes
            D.set_csr_row_order(P.tile_order_3d(32, 32, 16, tile=(32, 4, 2), z0=r0 // 1024, z1=r1 // 1024))
        x = np.asfortranarray(P.guess(n, m, r0, r1))
        dx, dax = K.DeviceArray.from_numpy(x), K.DeviceArray((r1 - r0, m))
        i32 = lambda v_: C.byref(C.c_int32(v_))  # noqa: E731
        D.lib().diaglib_b200_csr_matvec(i32(r1 - r0), i32(m), C.c_void_p(dx.ptr), C.c_void_p(dax.ptr))
        D.lib().diaglib_b200_sync()
        q.put((rank, r0, r1, dax.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("m,tiled", [(37, False), (8, False), (37, True), (5, True)])
def test_two_rank_spmm_bit_exact(oracle, m, tiled):
    """halo exchange (own stream and communicator, overlapped with the rows that do not touch the
    halo) + the column-chunked short-row SpMM (m > 24: two launches, halo block offset per chunk),
    in the natural and in a tiled processing order, reproduce the single-rank oracle product bit
    for bit"""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_spmm_worker, args=(r, world, port, m, tiled, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 32 * 32 * 16
    oracle.set_csr(*P.lap3d(32, 32, 16, delta=1.0))
    ref = oracle.csr_matvec(P.guess(n, m))
    got = np.vstack([r[3] for r in res])
    assert np.array_equal(got, ref)


def _wide_worker(rank, world, port, case, q):
    import torch
    import torch.distributed as dist

    import diaglib_b200 as D
    from diaglib_b200 import dist as DD, partition

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D.init(rank)
        DD.init_comm(dist)
        n = 1 << 13
        gen = lambda a, b: P.toy_sparse(n, a, b)  # noqa: E731
        r0, r1 = partition.row_range(n, rank, world)
        if case == "caslr_eff":
            n_targ = 6
            n_max = P.n_eig_rule(n_targ)
            DD.install_partitioned(gen, n, rank, world, dist, lr_rows=lambda a, b: P.caslr_like(n, a, b))
            lr = P.caslr_like(n)
            g = np.zeros((2 * n, n_max), order="F")
            g[:n] = P.guess_lowest_diag(lr["aa_diag"] / lr["sigma_diag"], n_max)
            g += P.guess(2 * n, n_max) * (0.05 / np.sqrt(2 * n / 12.0))
            ev = np.asfortranarray(np.vstack([g[r0:r1], g[n + r0:n + r1]]))      # local [Y; Z]
            eig = np.zeros(n_max)
            ok = D.caslr_eff_driver(False, r1 - r0, 2 * (r1 - r0), n_targ, n_max, 100, 1e-9, 10, None, None, None, None,
                                    None, eig, ev)
        else:
            n_targ = 8
            n_max = P.n_eig_rule(n_targ)
            DD.install_partitioned(gen, n, rank, world, dist, metric_rows=lambda a, b: P.metric_like(gen(a, b), r0=a))
            ev = np.asfortranarray(P.guess(n, n_max, r0, r1))
            eig = np.zeros(n_max)
            ok = D.gen_david_driver(False, r1 - r0, n_targ, n_max, 100, 1e-8, 10, 0.0, None, None, None, eig, ev)
        q.put((rank, ok, len(D.last_history(n_max)["it"]), eig.copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["gen_david", "caslr_eff"])
def test_two_rank_widened_drivers(oracle, case):
    """gen_david_driver and caslr_eff_driver row-partitioned over two ranks against the single-rank oracle"""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_wide_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 1 << 13
    if case == "caslr_eff":
        n_targ = 6
        n_max = P.n_eig_rule(n_targ)
        lr = P.caslr_like(n)
        oracle.set_lr(lr["apb"], lr["amb"], lr["spd"], lr["smd"], lr["aa_diag"], lr["sigma_diag"])
        g = np.zeros((2 * n, n_max), order="F")
        g[:n] = P.guess_lowest_diag(lr["aa_diag"] / lr["sigma_diag"], n_max)
        g += P.guess(2 * n, n_max) * (0.05 / np.sqrt(2 * n / 12.0))
        ro = oracle.caslr_eff(np.asfortranarray(g), n_targ, 100, 1e-9, 10)
    else:
        n_targ = 8
        n_max = P.n_eig_rule(n_targ)
        csr = P.toy_sparse(n)
        oracle.set_csr(*csr)
        oracle.set_csr_b(*P.metric_like(csr))
        ro = oracle.gen_david(P.guess(n, n_max), n_targ, 100, 1e-8, 10)
    for rank, ok, its, eig in res:
        assert ok and ro["ok"]
        assert np.abs(eig[:n_targ] - ro["eig"][:n_targ]).max() / np.abs(ro["eig"][:n_targ]).max() < 1e-10
        assert abs(its - len(ro["it"])) <= 1
        assert np.array_equal(eig, res[0][3])


def _peer_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import diaglib_b200 as D
    from diaglib_b200 import dist as DD, partition

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D.init(rank)
        n, n_targ = 1 << 14, 8
        n_max = P.n_eig_rule(n_targ)
        r0, r1 = partition.row_range(n, rank, world)
        out = []
        for mode in ("1", "0"):   # peer window, then ncclAllReduce (a second comm_init replaces the communicator)
            os.environ["DIAGLIB_B200_PEER_REDUCE"] = mode
            DD.init_comm(dist)
            DD.install_partitioned(lambda a, b: P.toy_sparse(n, a, b), n, rank, world, dist)
            info0 = D.peer_info()
            for drv in ("lobpcg", "davidson"):
                ev = np.asfortranarray(P.guess(n, n_max, r0, r1))
                eig = np.zeros(n_max)
                if drv == "lobpcg":
                    ok = D.lobpcg_driver(False, False, r1 - r0, n_targ, n_max, 300, 1e-8, 0.0, None, None, None, eig, ev)
                else:
                    ok = D.davidson_driver(False, r1 - r0, n_targ, n_max, 300, 1e-8, 12, 0.0, None, None, eig, ev)
                h = D.last_history(n_max)
                out.append(dict(mode=mode, drv=drv, ok=bool(ok), eig=eig.copy(), hist_eig=np.array(h["eig"]), its=len(h["it"]),
                                info=D.peer_info(), window=info0["window_ranks"]))
        q.put((rank, out))
    finally:
        os.environ.pop("DIAGLIB_B200_PEER_REDUCE", None)
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_rank_peer_window_matches_nccl():
    """The k x k all-reduces through the mapped peer windows (one kernel: reduction of the Gram kernel's
    partials + stores into every rank's window + sum in rank order) against the same solves with
    ncclAllReduce: on two ranks a + b is the same number in either order, so the whole iteration
    history must agree bit for bit, for both drivers; the window must really have been used."""
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in res:
        by = {(o["mode"], o["drv"]): o for o in out}
        if by[("1", "lobpcg")]["window"] == 0:
            pytest.skip("cudaIpc mapping of the peer windows is not available on this box: NCCL path only")
        for drv in ("lobpcg", "davidson"):
            a, b = by[("1", drv)], by[("0", drv)]
            assert a["ok"] and b["ok"]
            assert a["window"] == 2 and a["info"]["calls"] > 0 and a["info"]["error"] == 0
            assert b["window"] == 0 and b["info"]["calls"] == 0
            assert a["its"] == b["its"]
            assert np.array_equal(a["hist_eig"], b["hist_eig"])
            assert np.array_equal(a["eig"], res[0][1][out.index(a)]["eig"])   # bit-identical across the ranks
