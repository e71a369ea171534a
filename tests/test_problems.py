"""CPU tests of the stateless problem generators."""
import numpy as np
import pytest

from diaglib_b200 import partition, problems as P


def test_bijection_is_bijective():
    for bits in (6, 12, 15):
        idx = np.arange(1 << bits)
        out = P.bijection(idx, bits, seed=3)
        assert np.array_equal(np.sort(out), idx.astype(np.uint64))


def test_guess_shards_agree():
    full = P.guess(1000, 7)
    part = P.guess(1000, 7, 300, 640)
    assert np.array_equal(full[300:640], part)
    assert full.min() >= -0.5 and full.max() < 0.5
    assert full.flags.f_contiguous


def test_toy_dense_matches_reference_formula():
    a = P.toy_dense(6)
    assert a[2, 2] == 4.0 and a[0, 1] == 1.0 / 3.0 and np.array_equal(a, a.T)


def _assemble(gen, n, parts):
    rps, cols, vals, diags = [], [], [], []
    for (r0, r1) in parts:
        rp, c, v, d = gen(r0, r1)
        rps.append(rp)
        cols.append(c)
        vals.append(v)
        diags.append(d)
    return rps, cols, vals, diags


def test_generators_symmetric_and_shardable():
    gens = {
        "toy_sparse": (lambda r0, r1: P.toy_sparse(512, r0, r1), 512),
        "lap3d": (lambda r0, r1: P.lap3d(8, 8, 8, r0, r1, delta=0.5), 512),
        "fci_like": (lambda r0, r1: P.fci_like(512, r0, r1, n_strides=10, bandwidth=64), 512),
    }
    for name, (gen, n) in gens.items():
        rp, c, v, d = gen(0, n)
        a = P.csr_to_dense(n, rp, c, v)
        assert np.array_equal(a, a.T), name
        assert np.array_equal(np.diag(a), d), name
        # columns sorted within each row
        for i in range(n):
            assert np.all(np.diff(c[rp[i]:rp[i + 1]]) > 0), name
        # shards reproduce the global matrix bit for bit
        r0, r1 = 100, 333
        rp2, c2, v2, d2 = gen(r0, r1)
        assert np.array_equal(c2, c[rp[r0]:rp[r1]]) and np.array_equal(v2, v[rp[r0]:rp[r1]]), name
        assert np.array_equal(d2, d[r0:r1]), name


def test_lap3d_diagonal_is_permuted_progression():
    n = 4096
    _, _, _, d = P.lap3d(16, 16, 16, delta=1.0)
    assert np.array_equal(np.sort(d), 6.0 + 1.0 + np.arange(n))


def test_partition_localize_roundtrip():
    n, size = 512, 4
    rp, c, v, d = P.lap3d(8, 8, 8)
    a = P.csr_to_dense(n, rp, c, v)
    x = P.guess(n, 3)
    ref = a @ x
    needed_all = []
    shards = []
    for r in range(size):
        r0, r1 = partition.row_range(n, r, size)
        rpl, cl, vl, dl = P.lap3d(8, 8, 8, r0, r1)
        needed_all.append(partition.needed_ranges(cl, n, r, size))
        shards.append((r0, r1, rpl, cl, vl))
    for r in range(size):
        r0, r1, rpl, cl, vl = shards[r]
        col_loc, n_halo, recv = partition.localize(cl, n, r, size, needed_all[r])
        plan = partition.halo_plan(recv, needed_all, n, r, size)
        # emulate the exchange: what each peer would send us
        halo = np.zeros((n_halo, 3))
        for (q, lo, hi, off) in recv:
            halo[off:off + hi - lo] = x[lo:hi]
        xe = np.vstack([x[r0:r1], halo])
        al = P.csr_to_dense(r1 - r0 + n_halo, rpl, col_loc, vl)
        assert np.abs(al @ xe - ref[r0:r1]).max() < 1e-12
        # send side of the plan is consistent with what the peers expect
        peer, s0, sc, ro, rc = plan
        for i, q in enumerate(peer):
            want = needed_all[q][r]
            if want is not None:
                assert s0[i] + r0 == want[0] and sc[i] == want[1] - want[0]


def test_localize_rejects_columns_outside_the_halo_and_union_covers_them():
    """a second matrix (metric, linear-response) whose remote columns reach beyond the first
    matrix's halo: localising it with the first matrix's ranges must fail loudly; with the
    union of both it must succeed and address the shared halo consistently"""
    n, size, rank = 512, 4, 1
    r0, r1 = partition.row_range(n, rank, size)
    _, c_a, _, _ = P.lap3d(8, 8, 8, r0, r1)                    # reaches one z-plane (64 rows) into the neighbours
    c_b = np.concatenate([c_a, np.array([r0 - 100, r1 + 99], dtype=c_a.dtype)])   # reaches 100 rows
    need_a = partition.needed_ranges(c_a, n, rank, size)
    need_b = partition.needed_ranges(c_b, n, rank, size)
    with pytest.raises(ValueError):
        partition.localize(c_b, n, rank, size, need_a)
    need = partition.union_ranges(need_a, need_b)
    assert need[rank] is None and need[0] == (r0 - 100, r0) and need[2] == (r1, r1 + 100)
    la, halo_a, recv_a = partition.localize(c_a, n, rank, size, need)
    lb, halo_b, recv_b = partition.localize(c_b, n, rank, size, need)
    assert halo_a == halo_b == 200 and recv_a == recv_b
    # the same global column maps to the same local index through either matrix
    assert np.array_equal(la, lb[:len(la)])
    assert lb[-2] == (r1 - r0) + 0 and lb[-1] == (r1 - r0) + 100 + 99


@pytest.mark.parametrize("curve", ["morton", "sweep"])
def test_tile_order_is_a_permutation_with_tile_blocks(curve):
    nx, ny, nz = 64, 16, 8
    o = P.tile_order_3d(nx, ny, nz, tile=(32, 4, 2), z0=2, z1=8, curve=curve)
    n_loc = nx * ny * 6
    assert o.dtype == np.int32 and np.array_equal(np.sort(o), np.arange(n_loc))
    # every 256 consecutive entries are one 32x4x2 tile, x fastest
    blk = o[:256].astype(np.int64)
    x, y, z = blk % nx, (blk // nx) % ny, blk // (nx * ny)
    assert x.max() - x.min() == 31 and y.max() - y.min() == 3 and z.max() - z.min() == 1
    assert np.array_equal(blk[:32], blk[0] + np.arange(32))
