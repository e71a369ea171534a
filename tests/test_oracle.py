"""CPU tests: the oracle against the known-answer data the reference's own test defines
(dense LAPACK on the toy matrix, main.f90:311-342) and against structural invariants."""
import json
import os

import numpy as np
import pytest

from diaglib_b200 import problems as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")
EPS = np.finfo(float).eps

# BASELINE.md section 2 (printed digits of the reference's lapack.txt check)
BASELINE_EIGS = [1.869398101309, 3.000476106191, 4.017712612105, 5.016812067990, 6.013523333955,
                 7.010610707515, 8.008385419234, 9.006729203366, 10.005490234949, 11.004549919231,
                 12.003824214451, 13.003254791546, 14.002801027083, 15.002434279316, 16.002134042057]


def gold_eigs():
    return np.array(json.load(open(os.path.join(GOLD, "toy_dense_eigs.json")))["eig"])


def test_golden_matches_baseline_md():
    assert np.abs(gold_eigs()[:15] - np.array(BASELINE_EIGS)).max() < 5e-12


def test_dense_lapack_crosscheck(oracle):
    a = P.toy_dense(1000)
    w, z = oracle.dsyev(a)
    assert np.abs(w[:20] - gold_eigs()).max() < 1e-10
    assert np.abs(a @ z[:, :5] - z[:, :5] * w[:5]).max() < 1e-9


@pytest.mark.parametrize("driver", ["lobpcg", "davidson"])
def test_c1_toy_driver(oracle, driver):
    """config C1: n=1000, 10 roots of 15, tol 1e-8, itmax 100, m_max 20 (main.f90:14-18)"""
    n, n_want, tol = 1000, 10, 1e-8
    n_eig = P.n_eig_rule(n_want)
    a = P.toy_dense(n)
    oracle.set_dense(a)
    ev = P.guess(n, n_eig)
    if driver == "lobpcg":
        r = oracle.lobpcg(ev, n_want, 100, tol, matvec="oracle_dense_matvec")
    else:
        r = oracle.davidson(ev, n_want, 100, tol, 20, matvec="oracle_dense_matvec")
    assert r["ok"] and r["status"] == 0
    g = gold_eigs()[:n_want]
    assert np.abs(r["eig"][:n_want] - g).max() / np.abs(g).max() < 1e-10
    # returned vectors: residuals below the requested tolerance, orthonormal
    x = ev[:, :n_want]
    res = a @ x - x * r["eig"][:n_want]
    assert (np.linalg.norm(res, axis=0) / np.sqrt(n)).max() < tol
    assert np.abs(res).max() < 10 * tol
    assert np.abs(x.T @ x - np.eye(n_want)).max() < 1e-12
    hist = json.load(open(os.path.join(GOLD, "c1_oracle_history.json")))[driver]
    assert abs(len(r["it"]) - hist["iterations"]) <= 1
    assert r["stats"]["qr_fallbacks"] == 0


def test_lobpcg_shift_is_returned_in_eig(oracle):
    """quirk: LOBPCG returns eig INCLUDING shift (diaglib.f90:416 vs 461)"""
    n, n_want = 300, 4
    a = P.toy_dense(n)
    oracle.set_dense(a)
    w = np.linalg.eigvalsh(a)
    ev = P.guess(n, 8)
    r = oracle.lobpcg(ev, n_want, 200, 1e-8, shift=2.5, matvec="oracle_dense_matvec")
    assert r["ok"]
    assert np.abs(r["eig"][:n_want] - (w[:n_want] + 2.5)).max() < 1e-8


def test_ortho_cd_invariants(oracle):
    rng = np.random.default_rng(0)
    u = np.asfortranarray(rng.standard_normal((5000, 24)))
    u[:, 3] = u[:, 2] + 1e-7 * rng.standard_normal(5000)  # nearly dependent columns: forces extra passes
    span = u.copy()
    growth, ok = oracle.ortho_cd(u)
    assert ok and growth >= 1.0
    assert np.linalg.norm(u.T @ u - np.eye(24)) < 1e-13
    # same span
    q, _ = np.linalg.qr(span)
    assert np.linalg.norm(u - q @ (q.T @ u)) < 1e-6


def test_ortho_vs_x_invariants(oracle):
    rng = np.random.default_rng(1)
    x, _ = np.linalg.qr(rng.standard_normal((4000, 30)))
    x = np.asfortranarray(x)
    u = np.asfortranarray(rng.standard_normal((4000, 12)) + x[:, :12] * 50.0)
    oracle.ortho_vs_x(x, u)
    assert np.linalg.norm(x.T @ u) < 1e-13
    assert np.linalg.norm(u.T @ u - np.eye(12)) < 1e-13


def test_ortho_qr(oracle):
    rng = np.random.default_rng(2)
    u = np.asfortranarray(rng.standard_normal((500, 9)))
    oracle.ortho(u)
    assert np.linalg.norm(u.T @ u - np.eye(9)) < 1e-13


def test_norm_est_bound(oracle):
    import ctypes as C
    rng = np.random.default_rng(3)
    L = np.asfortranarray(np.tril(rng.standard_normal((17, 17))))
    est = oracle.lib().oracle_norm_est(C.byref(C.c_int32(17)), L.ctypes.data_as(C.c_void_p))
    assert est >= np.linalg.norm(L, 2) - 1e-12
    assert abs(est - (np.abs(np.diag(L)).max() + np.linalg.norm(np.tril(L, -1)))) < 1e-12


def test_get_coeffs_orthogonal(oracle):
    rng = np.random.default_rng(4)
    n_max, n_act = 7, 5
    len_u = n_max + 2 * n_act
    s = rng.standard_normal((len_u, len_u))
    s = np.diag(np.arange(1, len_u + 1.0)) + 0.05 * (s + s.T)
    _, z = np.linalg.eigh(s)
    a_red = np.asfortranarray(z)
    u_x, u_p = oracle.get_coeffs(a_red, len_u, n_max, n_act)
    assert np.linalg.norm(u_x.T @ u_p) < 1e-14
    assert np.linalg.norm(u_p.T @ u_p - np.eye(n_act)) < 1e-14


def test_csr_callbacks_match_dense(oracle):
    n = 257
    rp, c, v, d = P.toy_sparse(n)
    a = P.csr_to_dense(n, rp, c, v)
    assert np.abs(a - a.T).max() == 0.0
    oracle.set_csr(rp, c, v, d)
    x = P.guess(n, 5)
    ax = oracle.csr_matvec(x)
    assert np.abs(ax - a @ x).max() < 1e-13
    px = oracle.diag_precnd(x, -1.5)
    assert np.abs(px - x / (d[:, None] - 1.5)).max() < 1e-15
    # guard |d + fac| <= 1e-5 -> copy (main.f90:161-166)
    px = oracle.diag_precnd(x, -d[3])
    assert np.all(px[3] == x[3])


def test_sparse_lobpcg_small(oracle):
    nx = 16
    n = nx ** 3
    rp, c, v, d = P.lap3d(nx, nx, nx, delta=256.0 / n)
    oracle.set_csr(rp, c, v, d)
    a = P.csr_to_dense(n, rp, c, v)
    w = np.linalg.eigvalsh(a)
    ev = P.guess(n, 9)
    r = oracle.lobpcg(ev, 4, 300, 1e-8)
    assert r["ok"]
    assert np.abs(r["eig"][:4] - w[:4]).max() / abs(w[3]) < 1e-10


def test_oracle_gen_eig_matches_dense_pencil(oracle):
    """gen_eig branch of the restatement (diaglib.f90:299-302, 357-364, 422-436, 500-526) against
    LAPACK's dense generalized eigensolver on the same pencil"""
    import scipy.linalg as sl
    import scipy.sparse as sp
    from diaglib_b200 import problems as P
    n, n_targ, n_max = 600, 4, 9
    csr = P.toy_sparse(n)
    bcsr = P.metric_like(csr)
    oracle.set_csr(*csr)
    oracle.set_csr_b(*bcsr)
    ev = P.guess(n, n_max)
    r = oracle.lobpcg(ev, n_targ, 300, 1e-8, gen_eig=True)
    a = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(n, n)).toarray()
    b = sp.csr_matrix((bcsr[2], bcsr[1], bcsr[0]), shape=(n, n)).toarray()
    w = sl.eigh(a, b, eigvals_only=True, subset_by_index=[0, n_targ - 1])
    assert r["ok"]
    assert np.abs(r["eig"][:n_targ] - w).max() / np.abs(w).max() < 1e-10
    x = ev[:, :n_targ]
    assert np.abs(x.T @ b @ x - np.eye(n_targ)).max() < 1e-12


def test_oracle_b_ortho(oracle):
    from diaglib_b200 import problems as P
    n, m = 800, 7
    bcsr = P.metric_like(P.toy_sparse(n))
    oracle.set_csr_b(*bcsr)
    rng = np.random.default_rng(3)
    u = np.asfortranarray(rng.standard_normal((n, m)))
    bu = oracle.csr_bvec(u)
    span = u.copy()
    oracle.b_ortho(u, bu)
    assert np.abs(u.T @ bu - np.eye(m)).max() < 1e-12
    assert np.abs(bu - oracle.csr_bvec(u)).max() < 1e-12          # bu stays B u
    assert np.linalg.matrix_rank(np.hstack([span, u]), tol=1e-8) == m   # same span


def test_oracle_gen_david_and_the_reference_restart(oracle):
    """gen_david_driver restated (diaglib.f90:1855-2250).  With the evident intent at the restart
    (keep B times the restart vectors) it agrees with LAPACK's generalized eigensolver; with the
    reference's literal `bspace = zero` (2200) a run that restarts returns wrong eigenvalues while
    reporting ok -- the reason the product keeps the intended behaviour."""
    import scipy.linalg as sl
    import scipy.sparse as sp
    from diaglib_b200 import problems as P
    n, n_targ, n_max = 600, 4, 9
    csr = P.toy_sparse(n)
    bcsr = P.metric_like(csr)
    oracle.set_csr(*csr)
    oracle.set_csr_b(*bcsr)
    a = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(n, n)).toarray()
    b = sp.csr_matrix((bcsr[2], bcsr[1], bcsr[0]), shape=(n, n)).toarray()
    w = sl.eigh(a, b, eigvals_only=True, subset_by_index=[0, n_targ - 1])
    r = oracle.gen_david(P.guess(n, n_max), n_targ, 100, 1e-8, 10)
    assert r["ok"] and len(r["it"]) > 10                      # went through a restart
    assert np.abs(r["eig"][:n_targ] - w).max() / np.abs(w).max() < 1e-10
    lit = oracle.gen_david(P.guess(n, n_max), n_targ, 100, 1e-8, 10, reference_restart=True)
    assert np.abs(lit["eig"][:n_targ] - w).max() > 1e-3       # the literal reference statement is wrong
    # identical up to the restart
    k = 10
    assert np.allclose(lit["hist_eig"][:k, :n_targ], r["hist_eig"][:k, :n_targ], rtol=0, atol=1e-12)


def test_oracle_caslr_eff_matches_dense_pencil(oracle):
    """caslr_eff_driver restated (diaglib.f90:1024-1481) against LAPACK on the full 2n x 2n pencil,
    the check the reference's own test_caslr does by hand (main.f90:601-625)"""
    import scipy.linalg as sl
    import scipy.sparse as sp
    from diaglib_b200 import problems as P
    n, n_targ, n_max = 400, 4, 9
    lr = P.caslr_like(n)
    m = {k: sp.csr_matrix((lr[k][2], lr[k][1], lr[k][0]), shape=(n, n)).toarray() for k in ("apb", "amb", "spd", "smd")}
    a, b = 0.5 * (m["apb"] + m["amb"]), 0.5 * (m["apb"] - m["amb"])
    sg, dl = 0.5 * (m["spd"] + m["smd"]), 0.5 * (m["spd"] - m["smd"])
    af, sf = np.block([[a, b], [b, a]]), np.block([[sg, dl], [-dl, -sg]])
    w = sl.eigh(sf, af, eigvals_only=True)
    ref = np.sort(1.0 / w[w > 0])[:n_targ]
    oracle.set_lr(lr["apb"], lr["amb"], lr["spd"], lr["smd"], lr["aa_diag"], lr["sigma_diag"])
    ev = P.guess(2 * n, n_max)
    r = oracle.caslr_eff(ev, n_targ, 100, 1e-8, 10)
    assert r["ok"]
    assert np.abs(r["eig"][:n_targ] - ref).max() / ref.max() < 1e-10
    x = ev[:, :n_targ]
    res = af @ x - (sf @ x) * r["eig"][:n_targ]
    assert (np.linalg.norm(res, axis=0) / np.linalg.norm(af @ x, axis=0)).max() < 1e-6


@pytest.mark.parametrize("threads", [1, 8])
def test_c3_workload_golden_nx32(oracle, threads):
    """The benchmark workload (bench.py, C3) at 32^3: the oracle reproduces the committed fixture
    written by `bench.py --impl reference --nx 32`, whatever the BLAS thread count (the fixture is
    what bench.py's parity block and the GPU parity tests compare with at 128^3 and 256^3)."""
    import bench
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c3_oracle_nx32.json")))
    nx = gold["nx"]
    n_max = P.n_eig_rule(bench.N_TARG)
    csr = P.lap3d(nx, nx, nx, delta=bench.DELTA)
    oracle.set_threads(threads)
    try:
        oracle.set_csr(*csr)
        ev = bench.make_guess(csr[3], nx ** 3, n_max, 0, nx ** 3)
        r = oracle.lobpcg(ev, bench.N_TARG, bench.MAX_ITER, bench.TOL)
    finally:
        oracle.set_threads(min(8, os.cpu_count() or 1))
    assert r["ok"] and gold["ok"]
    assert abs(len(r["it"]) - gold["iterations"]) <= 1
    e, eg = r["eig"][:bench.N_TARG], np.array(gold["eig"][:bench.N_TARG])
    assert np.max(np.abs(e - eg) / np.abs(eg)) < 1e-10
    assert r["rms"][-1][:bench.N_TARG].max() < bench.TOL


def test_bench_parity_block_logic():
    import bench
    gold = bench.load_oracle_result(32)
    assert gold is not None and gold["_source"].endswith("c3_oracle_nx32.json")
    e = np.array(gold["eig"])
    ok = bench.parity_block(gold, gold["iterations"] + 1, e * (1 + 5e-11), np.array(gold["rms"]), np.array(gold["max"]))
    assert ok["ok"] and ok["its_oracle"] == gold["iterations"]
    bad_e = bench.parity_block(gold, gold["iterations"], e * (1 + 1e-9), np.array(gold["rms"]), np.array(gold["max"]))
    bad_i = bench.parity_block(gold, gold["iterations"] + 2, e, np.array(gold["rms"]), np.array(gold["max"]))
    bad_r = bench.parity_block(gold, gold["iterations"], e, np.array(gold["rms"]) + 1e-7, np.array(gold["max"]))
    assert not bad_e["ok"] and not bad_i["ok"] and not bad_r["ok"]


def test_oracle_accurate_reduced_eig_is_a_diagnostic_not_the_default(oracle):
    """the oracle's reduced eigenproblems go through LAPACK dsyev like the reference
    (diaglib.f90:315,406,1708); oracle_set_accurate_eig(1) switches them to dpotrf + dgesvj for ONE
    purpose: to measure how much of an iteration-count difference is dsyev's absolute accuracy
    (bench.parity_block).  On a graded LOBPCG-like matrix dsyev leaves residuals ~eps*|A|, the
    diagnostic route residuals relative to each eigenvalue."""
    import ctypes as C
    rng = np.random.default_rng(12)
    k = 111
    d = np.concatenate([np.arange(7.0, 44.0), 50 + 1e3 * rng.random(37), 1e6 + 1e7 * rng.random(37)])
    cpl = rng.standard_normal((k, k))
    cpl = 1e-3 * (cpl + cpl.T) * np.sqrt(np.outer(d, d)) / d.max() ** 0.5
    a = np.diag(d) + cpl
    np.fill_diagonal(a, d)

    def red(upper):
        z = np.asfortranarray((np.triu(a) if upper else np.tril(a)).copy())
        w, info = np.zeros(k), C.c_int32(0)
        oracle.lib().oracle_reduced_eig(C.byref(C.c_int32(k)), z.ctypes.data_as(C.c_void_p), C.byref(C.c_int32(k)),
                                        w.ctypes.data_as(C.c_void_p), C.byref(info), C.byref(C.c_int32(1 if upper else 0)))
        assert info.value == 0
        al, zl, wl = a.astype(np.longdouble), z.astype(np.longdouble), w.astype(np.longdouble)
        res = np.linalg.norm((al @ zl - zl * wl).astype(np.float64), axis=0)
        return w, z, res

    w0, z0, r0 = red(False)                      # default = dsyev
    oracle.set_accurate_eig(True)
    try:
        for upper in (False, True):
            w1, z1, r1 = red(upper)
            assert (r1[:37] / w1[:37]).max() < 1e-12
            assert np.abs(z1.T @ z1 - np.eye(k)).max() < 1e-13
            assert np.abs(w1 - w0).max() < 1e-8 * 10 and np.all(np.diff(w1) > 0)
    finally:
        oracle.set_accurate_eig(False)
    assert (r0[:37] / w0[:37]).max() > 1e-11     # what the reference's dsyev delivers on this matrix
    w2, _, _ = red(False)
    assert np.array_equal(w2, w0)                # the switch is off again
