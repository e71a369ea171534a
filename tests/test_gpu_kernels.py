"""GPU parity tests of the individual sm_100a kernels against numpy / the CPU oracle, all
through the C-ABI (include/diaglib_b200_kernels.h)."""
import numpy as np
import pytest

from diaglib_b200 import problems as P

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


@pytest.fixture(scope="module")
def K(gpu_lib):
    from diaglib_b200 import kernels
    return kernels


def rnd(n, m, seed):
    return np.asfortranarray(np.random.default_rng(seed).standard_normal((n, m)))


# ---- gram: dgemm('t','n') replacement -------------------------------------------------------
@pytest.mark.parametrize("n,p,q", [(1000, 15, 15), (4096, 37, 37), (4099, 74, 37), (10000, 111, 111), (33, 5, 3),
                                   (7, 3, 2), (20000, 128, 128), (3001, 1, 1), (5000, 133, 133), (2048, 300, 15)])
def test_gram_full(K, n, p, q):
    a, b = rnd(n, p, 1), rnd(n, q, 2)
    da, db = K.DeviceArray.from_numpy(a), K.DeviceArray.from_numpy(b)
    c = K.gram(da, p, db, q)
    ref = a.T @ b
    assert np.abs(c - ref).max() <= 1e-12 * np.sqrt(n) * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("n,p", [(1000, 15), (4097, 37), (10000, 111), (9000, 74), (6000, 133), (3000, 8), (2500, 260)])
def test_gram_symmetric_lower(K, n, p):
    """V^T (A V) with symmetric A: only the lower triangle is computed, then mirrored"""
    v = rnd(n, p, 3)
    d = np.linspace(1.0, 2.0, n)
    av = np.asfortranarray(v * d[:, None])
    dv, dav = K.DeviceArray.from_numpy(v), K.DeviceArray.from_numpy(av)
    c = K.gram(dv, p, dav, p, sym_lower=True)
    ref = v.T @ av
    assert np.abs(c - ref).max() <= 1e-12 * np.sqrt(n) * np.abs(ref).max()
    assert np.array_equal(c, c.T)
    c2 = K.gram(dv, p, dv, p, sym_lower=True)  # same operand (ortho_cd metric)
    assert np.abs(c2 - v.T @ v).max() <= 1e-12 * np.sqrt(n) * np.abs(ref).max()


def test_gram_column_offsets_and_odd_ld(K):
    """sub-blocks of a wider block with an odd leading dimension exercise the 8-byte loader"""
    n, w = 3001, 40
    a = rnd(n, w, 4)
    da = K.DeviceArray.from_numpy(a)
    c = K.gram(da, 7, da, 11, a_off=3, b_off=20)
    assert np.abs(c - a[:, 3:10].T @ a[:, 20:31]).max() < 1e-10


def test_gram_is_deterministic(K):
    a = rnd(50000, 37, 5)
    da = K.DeviceArray.from_numpy(a)
    c1 = K.gram(da, 37, da, 37, sym_lower=True)
    c2 = K.gram(da, 37, da, 37, sym_lower=True)
    assert np.array_equal(c1, c2)


# ---- block_mul: dgemm('n','n') / dtrmm replacement -------------------------------------------
@pytest.mark.parametrize("n,p,q", [(1000, 15, 15), (4096, 111, 37), (4099, 74, 37), (130, 30, 5), (7, 3, 2),
                                   (5000, 111, 74), (5001, 128, 128), (3000, 300, 15), (2000, 40, 133), (999, 1, 1)])
def test_block_mul(K, n, p, q):
    v, c, y0 = rnd(n, p, 6), rnd(p, q, 7), rnd(n, q, 8)
    dv, dy = K.DeviceArray.from_numpy(v), K.DeviceArray.from_numpy(y0)
    K.block_mul(dv, p, c, dy, alpha=1.0, beta=0.0)
    ref = v @ c
    assert np.abs(dy.numpy() - ref).max() <= 1e-13 * p * max(1.0, np.abs(ref).max())
    dy2 = K.DeviceArray.from_numpy(y0)
    K.block_mul(dv, p, c, dy2, alpha=-1.0, beta=1.0)  # U -= X (X^T U), diaglib.f90:3544
    ref2 = y0 - v @ c
    assert np.abs(dy2.numpy() - ref2).max() <= 1e-13 * p * max(1.0, np.abs(ref2).max())


def test_block_mul_in_place_row_local(K):
    n, m = 5003, 37
    u = rnd(n, m, 9)
    t = np.asfortranarray(np.triu(rnd(m, m, 10)))
    du = K.DeviceArray.from_numpy(u)
    K.block_mul(du, m, t, du)
    assert np.abs(du.numpy() - u @ t).max() < 1e-11


def test_block_mul_writes_into_own_columns(K):
    """P = space * u_p written into columns of space itself (diaglib.f90:495-496)"""
    n, w, p, q = 3000, 30, 20, 6
    s = rnd(n, w, 11)
    c = rnd(p, q, 12)
    ds = K.DeviceArray.from_numpy(s)
    K.block_mul(ds, p, c, ds, y_off=22)
    out = ds.numpy()
    assert np.array_equal(out[:, :22], s[:, :22])
    assert np.abs(out[:, 22:28] - s[:, :p] @ c).max() < 1e-11


@pytest.mark.parametrize("n,p,q,tri,inplace", [(4096, 37, 37, True, True), (10000, 74, 37, False, False), (5002, 15, 15, True, True),
                                                 (3000, 111, 8, False, False), (700, 37, 37, True, True), (4001, 37, 37, True, True),
                                                 (6000, 60, 50, False, False)])
def test_block_mul_gram_fused(K, n, p, q, tri, inplace):
    """dtrmm / projection fused with the metric of the result (and its fallbacks: odd n, q > 40)"""
    v, y0 = rnd(n, p, 31), rnd(n, q, 32)
    c = rnd(p, q, 33)
    if tri:
        c = np.asfortranarray(np.triu(c))
    dv = K.DeviceArray.from_numpy(v)
    if inplace:
        g = K.block_mul_gram(dv, p, c, dv, upper_tri=tri)
        ref = v @ c
        out = dv.numpy()
    else:
        dy = K.DeviceArray.from_numpy(y0)
        g = K.block_mul_gram(dv, p, c, dy, alpha=-1.0, beta=1.0)
        ref = y0 - v @ c
        out = dy.numpy()
    assert np.abs(out - ref).max() <= 1e-13 * p * max(1.0, np.abs(ref).max())
    gref = ref.T @ ref
    assert np.abs(g - gref).max() <= 1e-12 * np.sqrt(n) * np.abs(gref).max()
    assert np.array_equal(g, g.T)


# ---- residual + norms (dcopy/daxpy/dnrm2/maxval fusion) --------------------------------------
def test_residual_norms(K):
    n, m = 10007, 9
    ax, x = rnd(n, m, 13), rnd(n, m, 14)
    theta = np.linspace(0.5, 3.0, m)
    active = np.array([1, 1, 0, 1, 0, 1, 1, 1, 0], np.int32)
    dax, dx, dr = K.DeviceArray.from_numpy(ax), K.DeviceArray.from_numpy(x), K.DeviceArray((n, m))
    ss, mx = K.residual(dax, dx, theta, active, dr)
    r = dr.numpy()
    for j in range(m):
        if active[j]:
            ref = ax[:, j] - theta[j] * x[:, j]
            assert np.abs(r[:, j] - ref).max() < 1e-14
            assert abs(ss[j] - ref @ ref) < 1e-10 * (ref @ ref)
            assert abs(mx[j] - np.abs(ref).max()) < 1e-14
        else:
            assert np.array_equal(r[:, j], ax[:, j])


# ---- symmetric eigensolver (dsyev replacement) ------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 3, 15, 30, 45, 74, 111, 112, 118, 119, 133, 150, 300, 399, 700])
def test_sym_eig_vs_lapack(K, oracle, k):
    rng = np.random.default_rng(k)
    s = rng.standard_normal((k, k))
    a = np.diag(np.arange(1.0, k + 1)) + 0.3 * (s + s.T)
    w_ref, _ = oracle.dsyev(a)
    for upper in (False, True):
        tri = np.triu(a) if upper else np.tril(a)  # the other triangle must not be referenced
        tri = tri + (np.tril(np.full((k, k), np.nan), -1) if upper else np.triu(np.full((k, k), np.nan), 1))
        w, z, sweeps = K.sym_eig(tri, upper=upper)
        scale = np.abs(w_ref).max()
        assert np.abs(w - w_ref).max() <= 50 * EPS * scale * max(1, np.sqrt(k))
        assert np.abs(a @ z - z * w).max() <= 200 * EPS * scale * np.sqrt(k)
        assert np.abs(z.T @ z - np.eye(k)).max() <= 100 * EPS * np.sqrt(k)
        assert np.all(np.diff(w) >= 0)


def test_sym_eig_lobpcg_like_matrix(K, oracle):
    """a_red of a converging LOBPCG: diag(eig) block plus small couplings; some exact zeros"""
    rng = np.random.default_rng(5)
    k = 45
    a = np.diag(np.concatenate([np.arange(2.0, 17.0), 50 + 10 * rng.random(30)]))
    c = 1e-6 * rng.standard_normal((k, k))
    a += c + c.T
    a[20, :5] = a[:5, 20] = 0.0
    w_ref, _ = oracle.dsyev(a)
    w, z, _ = K.sym_eig(a)
    assert np.abs(w - w_ref).max() < 1e-13 * 60
    # sign convention: largest component positive (deterministic across ranks)
    assert np.all(z[np.abs(z).argmax(axis=0), np.arange(k)] > 0)


def test_sym_eig_indefinite_and_plus_minus_pairs(K, oracle):
    """indefinite input and exact +-lambda pairs (the case a one-sided Jacobi cannot resolve)"""
    rng = np.random.default_rng(11)
    k = 40
    q, _ = np.linalg.qr(rng.standard_normal((k, k)))
    lam = np.linspace(-3.0, 5.0, k)
    a = (q * lam) @ q.T
    a = 0.5 * (a + a.T)
    w, z, sweeps = K.sym_eig(a)
    assert np.abs(w - np.sort(lam)).max() < 1e-13 * 10
    assert np.abs(a @ z - z * w).max() < 1e-13 * 10
    lam2 = np.concatenate([np.linspace(1.0, 2.0, k // 2), -np.linspace(1.0, 2.0, k // 2)])
    b = (q * lam2) @ q.T
    b = 0.5 * (b + b.T)
    w2, z2, sweeps2 = K.sym_eig(b)
    assert np.abs(w2 - np.sort(lam2)).max() < 1e-13 * 10
    assert np.abs(b @ z2 - z2 * w2).max() < 1e-13 * 10
    assert np.abs(z2.T @ z2 - np.eye(k)).max() < 1e-13
    c = np.array([[0.0, 1.0], [1.0, 0.0]])
    w3, z3, sweeps3 = K.sym_eig(c)
    assert np.allclose(w3, [-1.0, 1.0]) and np.abs(c @ z3 - z3 * w3).max() < 1e-15


def graded_lobpcg_like(k=111, seed=12):
    rng = np.random.default_rng(seed)
    d = np.concatenate([np.arange(7.0, 44.0), 50 + 1e3 * rng.random(37), 1e6 + 1e7 * rng.random(37)])
    cpl = rng.standard_normal((k, k))
    cpl = 1e-3 * (cpl + cpl.T) * np.sqrt(np.outer(d, d)) / d.max() ** 0.5
    a = np.diag(d) + cpl
    np.fill_diagonal(a, d)
    return a


def rayleigh_ld(a, z):
    """Rayleigh quotients of the columns of z in extended precision: within |r|^2 / gap of an
    eigenvalue, i.e. an independent reference for eigenvalues that LAPACK only delivers to
    eps |A| absolute"""
    al, zl = a.astype(np.longdouble), z.astype(np.longdouble)
    return ((zl * (al @ zl)).sum(0) / (zl * zl).sum(0)).astype(np.float64)


@pytest.mark.parametrize("mode", [0, 1])
def test_sym_eig_graded_matrix_relative_accuracy(K, oracle, mode):
    """LOBPCG-like reduced matrix: Ritz values 7..44 next to 1e6-1e7 (W block): the small
    eigenvalues must keep ~1e-13 RELATIVE accuracy (the parity bar is 1e-10 relative).
    mode 0: one-sided Jacobi on the Cholesky factor; mode 1: two-sided Jacobi."""
    a = graded_lobpcg_like()
    prev = K.set_eig_mode(mode)
    try:
        w, z, sweeps = K.sym_eig(a)
        path = K.sym_eig.last_path
    finally:
        K.set_eig_mode(prev)
    assert path == (1 if mode == 0 else 2)
    # rigorous check: for symmetric A an eigenvalue lies within |A z - w z| / |z| of w.  Residual in
    # extended precision.  (LAPACK's tridiagonal QR is only absolutely accurate, eps |A| ~ 1e-9 here,
    # so it is not a usable reference for the small eigenvalues of this matrix.)
    al, zl, wl = a.astype(np.longdouble), z.astype(np.longdouble), w.astype(np.longdouble)
    res = np.linalg.norm((al @ zl - zl * wl).astype(np.float64), axis=0) / np.linalg.norm(z, axis=0)
    assert np.abs(z.T @ z - np.eye(len(w))).max() < 1e-14
    # both solvers: residuals small relative to EACH eigenvalue (LAPACK only delivers eps |A|)
    assert (res[:37] / np.abs(w[:37])).max() < 1e-12
    assert (res / np.abs(w)).max() < 1e-11
    rq = rayleigh_ld(a, z)
    assert (np.abs(w - rq) / np.abs(rq)).max() < 1e-13
    assert sweeps <= 8


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 9, 15, 37, 63, 64, 74, 111, 112, 133, 158, 159, 210, 266, 399, 700, 1330])
def test_sym_eig_one_sided_positive_definite(K, oracle, k):
    """positive definite reduced matrices take the one-sided solver (Cholesky factor + block
    Jacobi): all block sizes / CTA counts, whole-matrix and panel Cholesky (k <= 158 / beyond)"""
    rng = np.random.default_rng(1000 + k)
    s = rng.standard_normal((k, k))
    a = np.diag(np.arange(1.0, k + 1)) + 0.05 * (s + s.T) + (s @ s.T) / (4 * k)
    w_ref = np.linalg.eigvalsh(a)
    assert w_ref[0] > 0
    for upper in (False, True):
        tri = np.triu(a) if upper else np.tril(a)
        tri = tri + (np.tril(np.full((k, k), np.nan), -1) if upper else np.triu(np.full((k, k), np.nan), 1))
        w, z, sweeps = K.sym_eig(tri, upper=upper)
        assert K.sym_eig.last_path == 1
        scale = np.abs(w_ref).max()
        assert np.abs(w - w_ref).max() <= 50 * EPS * scale * max(1, np.sqrt(k))
        assert np.abs(a @ z - z * w).max() <= 200 * EPS * scale * np.sqrt(k)
        assert np.abs(z.T @ z - np.eye(k)).max() <= 100 * EPS * np.sqrt(k)
        assert np.all(np.diff(w) >= 0)
        assert np.all(z[np.abs(z).argmax(axis=0), np.arange(k)] > 0)
        assert sweeps <= 12


@pytest.mark.parametrize("block", [4, 8])
def test_sym_eig_one_sided_block_sizes_agree(K, block):
    a = graded_lobpcg_like()
    prev = K.set_eig_mode(0, block)
    try:
        w, z, _ = K.sym_eig(a)
        assert K.sym_eig.last_path == 1
    finally:
        K.set_eig_mode(prev, 0)
    rq = rayleigh_ld(a, z)
    assert (np.abs(w - rq) / np.abs(rq)).max() < 1e-13


def test_sym_eig_not_positive_definite_falls_back(K):
    """zero / negative diagonal, indefinite with positive diagonal, NaN: the one-sided solver
    declines on the device and the two-sided solver delivers in the same call"""
    rng = np.random.default_rng(3)
    k = 37
    s = rng.standard_normal((k, k))
    a = np.diag(np.arange(1.0, k + 1)) + 2.0 * (s + s.T)      # positive diagonal, indefinite
    assert np.linalg.eigvalsh(a)[0] < 0
    w, z, _ = K.sym_eig(a)
    assert K.sym_eig.last_path == 2
    assert np.abs(a @ z - z * w).max() < 1e-12
    b = a.copy()
    b[5, 5] = -1.0
    w, z, _ = K.sym_eig(b)
    assert K.sym_eig.last_path == 2 and np.abs(b @ z - z * w).max() < 1e-12
    w, z, _ = K.sym_eig(np.zeros((4, 4)))
    assert K.sym_eig.last_path == 2 and np.abs(w).max() == 0.0


def test_small_kernel_timing_table(K, gpu_lib):
    """not a pass/fail test of speed: device time of the replicated single-CTA kernels (pytest -s)"""
    L = gpu_lib.lib()
    for m in (13, 21, 37, 74, 112, 133):
        t_blk = 1e3 * L.diaglib_b200_k_time_small(0, m, 0, 0, 50)
        L.diaglib_b200_k_set_tuning(b"chol_blocked", 0)
        try:
            t_col = 1e3 * L.diaglib_b200_k_time_small(0, m, 0, 0, 50)
        finally:
            L.diaglib_b200_k_set_tuning(b"chol_blocked", 1)
        print(f"chol_inv m={m}: {t_blk:.1f} us blocked (8 columns per step), {t_col:.1f} us column by column")
    for len_u, n_max, n_act in ((111, 37, 37), (111, 37, 20), (74, 37, 37), (63, 21, 21), (399, 133, 133)):
        row = []
        for threads in (1024, 512, 256):
            prev = L.diaglib_b200_k_set_tuning(b"coeffs_threads", threads)
            try:
                row.append(1e3 * L.diaglib_b200_k_time_small(1, len_u, n_max, n_act, 20))
            finally:
                L.diaglib_b200_k_set_tuning(b"coeffs_threads", prev)
        prev = L.diaglib_b200_k_set_tuning(b"coeffs_smem", 0)
        try:
            glob = 1e3 * L.diaglib_b200_k_time_small(1, len_u, n_max, n_act, 20)
        finally:
            L.diaglib_b200_k_set_tuning(b"coeffs_smem", prev)
        print(f"get_coeffs len_u={len_u} n_max={n_max} n_act={n_act}: {row[0]:.1f} us (1024 threads), {row[1]:.1f} (512), "
              f"{row[2]:.1f} (256); working set in global memory (round 1): {glob:.1f} us")
        assert row[0] > 0


def test_sym_eig_timing_table(K):
    """not a pass/fail test of speed: prints the per-solve time of both solvers (pytest -s)"""
    a = graded_lobpcg_like()
    t = []
    for mode in (0, 1):
        prev = K.set_eig_mode(mode)
        try:
            t.append(K.sym_eig_time_ms(a, reps=20))
            _, _, sw = K.sym_eig(a)
            t.append(sw)
        finally:
            K.set_eig_mode(prev)
    print(f"sym_eig graded LOBPCG-like k=111: one-sided {t[0]:.3f} ms ({t[1]} sweeps), two-sided {t[2]:.3f} ms ({t[3]} sweeps)")
    for k in (37, 74, 111, 210, 399, 700, 1330):
        rng = np.random.default_rng(k)
        s = rng.standard_normal((k, k))
        a = np.diag(np.arange(1.0, k + 1)) + 0.02 * (s + s.T) + (s @ s.T) / (4 * k)
        out = []
        for mode in (0, 1):
            if mode == 1 and k > 700:
                out.append(float("nan"))
                continue
            prev = K.set_eig_mode(mode)
            try:
                out.append(K.sym_eig_time_ms(a, reps=3 if k > 300 else 10))
            finally:
                K.set_eig_mode(prev)
        _, _, sw = K.sym_eig(a)
        print(f"sym_eig k={k}: one-sided {out[0]:.3f} ms ({sw} sweeps), two-sided {out[1]:.3f} ms")
        assert out[0] > 0


# ---- Cholesky factor + inverse + norm estimates (one ortho_cd pass) ---------------------------
@pytest.mark.parametrize("m", [1, 5, 15, 37, 64, 100, 133])
def test_chol_inv(K, oracle, m):
    import ctypes as C
    u = rnd(4 * m + 10, m, m)
    g = u.T @ u
    t, st = K.chol_inv(g)
    L = np.linalg.cholesky(g)
    assert st["info_first"] == 0 and st["n_shifts"] == 0 and st["hard_fail"] == 0
    assert np.abs(np.tril(t, -1)).max() == 0.0
    assert np.abs(t - np.linalg.inv(L).T).max() <= 1e-10 * np.abs(t).max()
    mi = C.byref(C.c_int32(m))
    Lf = np.asfortranarray(L)
    Li = np.asfortranarray(np.linalg.inv(L))
    assert abs(st["l_norm"] - oracle.lib().oracle_norm_est(mi, Lf.ctypes.data_as(C.c_void_p))) < 1e-10 * st["l_norm"]
    assert abs(st["linv_norm"] - oracle.lib().oracle_norm_est(mi, Li.ctypes.data_as(C.c_void_p))) < 1e-8 * st["linv_norm"]
    q = u @ t
    assert np.abs(q.T @ q - np.eye(m)).max() < 1e-10


def test_chol_inv_level_shift(K):
    """indefinite metric: dpotrf fails, the level-shift loop (diaglib.f90:3265-3295) rescues"""
    u = rnd(200, 12, 21)
    g = u.T @ u
    w, z = np.linalg.eigh(g)
    unorm = np.sqrt(np.trace(g))
    w[0] = -50 * EPS * unorm  # slightly negative direction, as produced by a rank-deficient block
    g2 = (z * w) @ z.T
    g2 = 0.5 * (g2 + g2.T)
    t, st = K.chol_inv(g2)
    assert st["info_first"] != 0 and st["n_shifts"] >= 1 and st["hard_fail"] == 0
    # shift = max(eps*alpha*unorm, tol_ortho), alpha = 100, 1000, ... (3267, 3287, 3291)
    un2 = np.sqrt(np.trace(g2))
    assert abs(st["shift"] - EPS * 100 * 10 ** (st["n_shifts"] - 1) * un2) <= 1e-3 * st["shift"]
    # T = L^-T of the SHIFTED metric: T^T (G + shift I) T = I
    gs = g2 + st["shift"] * np.eye(12)
    # (the rescued direction has a pivot of order sqrt(shift): only loosely accurate)
    assert np.abs(t.T @ gs @ t - np.eye(12)).max() < 0.1
    assert np.abs((t.T @ gs @ t - np.eye(12))[:11, :11]).max() < 1e-6


def test_chol_inv_hard_fail(K):
    """a strongly indefinite metric exhausts the 10 shifts (3276-3284)"""
    g = -np.eye(6)
    t, st = K.chol_inv(g)
    assert st["hard_fail"] == 1 and st["n_shifts"] == 10


# ---- get_coeffs (P coefficients) ----------------------------------------------------------------
@pytest.mark.parametrize("n_max,n_act,first", [(15, 15, True), (15, 9, False), (37, 37, False), (37, 20, False), (6, 1, False)])
def test_get_coeffs_vs_oracle(K, oracle, n_max, n_act, first):
    rng = np.random.default_rng(n_max * 100 + n_act)
    len_u = 2 * n_max if first else n_max + 2 * n_act
    s = rng.standard_normal((len_u, len_u))
    a = np.diag(np.arange(1.0, len_u + 1)) + 0.02 * (s + s.T)
    _, z, _ = K.sym_eig(a)
    z = np.asfortranarray(z)
    u_p, st = K.get_coeffs(z, len_u, n_max, n_act)
    u_x_ref, u_p_ref = oracle.get_coeffs(z, len_u, n_max, n_act)
    assert st["fail"] == 0 and st["qr"] == 0
    u_x = z[:, :n_max]
    assert np.linalg.norm(u_x.T @ u_p) < 1e-14 * np.sqrt(len_u) * 10
    assert np.linalg.norm(u_p.T @ u_p - np.eye(n_act)) < 1e-13
    assert np.abs(u_p - u_p_ref).max() < 1e-9


# ---- built-in callbacks -------------------------------------------------------------------------
@pytest.mark.parametrize("gen,m", [("lap3d", 37), ("toy_sparse", 13), ("fci_like", 5), ("lap3d", 3)])
def test_spmm_bit_exact_vs_oracle(K, gpu_lib, oracle, gen, m):
    import ctypes as C
    if gen == "lap3d":
        n = 16 ** 3
        rp, c, v, d = P.lap3d(16, 16, 16, delta=0.25)
    elif gen == "toy_sparse":
        n = 5000
        rp, c, v, d = P.toy_sparse(n)
    else:
        n = 4096
        rp, c, v, d = P.fci_like(n, n_strides=20, bandwidth=512)
    oracle.set_csr(rp, c, v, d)
    gpu_lib.set_csr(rp, c, v, d)
    x = P.guess(n, m)
    ref = oracle.csr_matvec(x)
    dx, dax = K.DeviceArray.from_numpy(x), K.DeviceArray((n, m))
    i32 = lambda v_: C.byref(C.c_int32(v_))  # noqa: E731
    gpu_lib.lib().diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(dx.ptr), C.c_void_p(dax.ptr))
    gpu_lib.lib().diaglib_b200_sync()
    assert np.array_equal(dax.numpy(), ref)  # same summation order, FMA on both sides
    # preconditioner incl. the |d+fac| <= 1e-5 guard (main.f90:161-166)
    for fac in (-1.5, -float(d[7])):
        refp = oracle.diag_precnd(x, fac)
        dpx = K.DeviceArray((n, m))
        gpu_lib.lib().diaglib_b200_diag_precnd(i32(n), i32(m), C.byref(C.c_double(fac)), C.c_void_p(dx.ptr),
                                              C.c_void_p(dpx.ptr))
        gpu_lib.lib().diaglib_b200_sync()
        assert np.array_equal(dpx.numpy(), refp)


# ---- row order of the built-in matvec, device-resident matrices, synthetic generator -----------
def _csr_matvec(gpu_lib, K, n, m, x):
    import ctypes as C
    dx, dax = K.DeviceArray.from_numpy(x), K.DeviceArray((n, m))
    i32 = lambda v_: C.byref(C.c_int32(v_))  # noqa: E731
    gpu_lib.lib().diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(dx.ptr), C.c_void_p(dax.ptr))
    gpu_lib.lib().diaglib_b200_sync()
    out = dax.numpy()
    dx.free()
    dax.free()
    return out


@pytest.mark.parametrize("tile,curve,m", [((32, 4, 2), "morton", 37), ((16, 4, 4), "sweep", 8), ((32, 8, 1), "morton", 5)])
def test_spmm_row_order_is_bit_exact(K, gpu_lib, oracle, tile, curve, m):
    """a locality-preserving processing order changes no bit of the result (every row sum is the
    same CSR-order FMA chain); a non-permutation is refused"""
    nx, ny, nz = 64, 16, 16
    n = nx * ny * nz
    csr = P.lap3d(nx, ny, nz, delta=0.25)
    oracle.set_csr(*csr)
    x = P.guess(n, m)
    ref = oracle.csr_matvec(x)
    gpu_lib.set_csr(*csr, row_order=P.tile_order_3d(nx, ny, nz, tile=tile, curve=curve))
    assert np.array_equal(_csr_matvec(gpu_lib, K, n, m, x), ref)
    bad = np.zeros(n, np.int32)
    with pytest.raises(gpu_lib.DiaglibError):
        gpu_lib.set_csr_row_order(bad)
    gpu_lib.set_csr_row_order(None)
    assert np.array_equal(_csr_matvec(gpu_lib, K, n, m, x), ref)


@pytest.mark.parametrize("bits,r0f,r1f", [(12, 0.0, 1.0), (13, 0.25, 0.5), (12, 0.75, 1.0), (12, 0.0, 0.25)])
def test_gen_fci_on_device_matches_the_host_generator(K, gpu_lib, bits, r0f, r1f):
    """tools/c4_run.py generates C4's matrix in HBM (10 GB per rank at n = 2^26): same integers and
    the same bits as diaglib_b200.problems.fci_like + partition.localize for any row block"""
    import ctypes as C
    from diaglib_b200 import partition
    n = 1 << bits
    r0, r1 = int(n * r0f), int(n * r1f)
    kw = dict(n_strides=12, bandwidth=1 << (bits - 3), big_delta=0.1, seed=1)
    rp, col, val, diag = P.fci_like(n, r0, r1, **kw)
    strides = P.fci_strides(12, min(1 << (bits - 3), n // 2), 1)
    lo_prev, hi_next = max(0, r0 - int(strides[-1])), min(n, r1 + int(strides[-1]))
    # expected local numbering: owned, then [lo_prev, r0), then [r1, hi_next)
    n_loc = r1 - r0
    exp = np.where((col >= r0) & (col < r1), col.astype(np.int64) - r0,
                   np.where(col < r0, n_loc + (col.astype(np.int64) - lo_prev), n_loc + (r0 - lo_prev) + (col.astype(np.int64) - r1)))
    d_rp = K.DeviceArray((len(rp), 1))          # 8-byte slots: int64 row pointers
    gpu_lib.lib().diaglib_b200_h2d(d_rp.ptr, rp.ctypes.data_as(C.c_void_p), rp.nbytes)
    nnz = int(rp[-1])
    d_col = K.DeviceArray(((nnz + 1) // 2 + 1, 1))
    d_val, d_diag = K.DeviceArray((nnz, 1)), K.DeviceArray((n_loc, 1))
    st = np.ascontiguousarray(strides, dtype=np.int64)
    rc = gpu_lib.lib().diaglib_b200_k_gen_fci(n, r0, r1, len(st), st.ctypes.data_as(C.c_void_p), 0.1, 1, lo_prev, hi_next,
                                             C.c_void_p(d_rp.ptr), C.c_void_p(d_col.ptr), C.c_void_p(d_val.ptr),
                                             C.c_void_p(d_diag.ptr))
    assert rc == 0
    got_col = np.zeros(nnz, np.int32)
    gpu_lib.lib().diaglib_b200_d2h(got_col.ctypes.data_as(C.c_void_p), C.c_void_p(d_col.ptr), got_col.nbytes)
    assert np.array_equal(got_col, exp.astype(np.int32))
    assert np.array_equal(d_val.numpy()[:, 0], val)
    assert np.array_equal(d_diag.numpy()[:, 0], diag)
    if (r0, r1) == (0, n):
        # the adopted device arrays drive the built-in callbacks like a host-installed matrix
        gpu_lib.set_csr_device(n, 0, nnz, d_rp.ptr, d_col.ptr, d_val.ptr, d_diag.ptr)
        x = P.guess(n, 5)
        got = _csr_matvec(gpu_lib, K, n, 5, x)
        import scipy.sparse as sp
        ref = sp.csr_matrix((val, col, rp), shape=(n, n)) @ x
        assert np.abs(got - ref).max() < 1e-13
        theta = np.arange(1.0, 6.0)
        norms = np.zeros(10)
        dx = K.DeviceArray.from_numpy(x)
        rc = gpu_lib.lib().diaglib_b200_k_true_residual(n, 5, C.c_void_p(dx.ptr), theta.ctypes.data_as(C.c_void_p),
                                                       norms.ctypes.data_as(C.c_void_p))
        assert rc == 0
        r = ref - x * theta
        assert np.allclose(norms[:5], (r * r).sum(0), rtol=1e-12) and np.allclose(norms[5:], np.abs(r).max(0), rtol=1e-12)
        dx.free()
        gpu_lib.set_csr(rp, col, val, diag)   # release the adopted pointers before they are freed
    for a in (d_rp, d_col, d_val, d_diag):
        a.free()
