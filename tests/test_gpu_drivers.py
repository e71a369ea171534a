"""GPU parity tests of the drivers and the public block routines against the CPU oracle on
identical inputs, through the reference-shaped interface (diaglib_b200.lobpcg_driver etc.).

Bar (BASELINE.json north_star): eigenvalues to 1e-10 relative, residual norms below the
requested tolerance, iteration count within +-1 of the reference algorithm."""
import json
import os

import numpy as np
import pytest

from diaglib_b200 import problems as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL = 1e-10


def dense_as_csr(a):
    n = a.shape[0]
    rowptr = np.arange(0, n * n + 1, n, dtype=np.int64)
    col = np.tile(np.arange(n, dtype=np.int32), n)
    return rowptr, col, np.ascontiguousarray(a).reshape(-1), np.diag(a).copy()


def install(D, oracle, csr):
    rp, c, v, d = csr
    oracle.set_csr(rp, c, v, d)
    D.set_csr(rp, c, v, d)


def check_solution(csr, eig, evec, n_targ, tol):
    rp, c, v, d = csr
    import scipy.sparse as sp
    n = len(rp) - 1
    a = sp.csr_matrix((v, c, rp), shape=(n, n))
    x = evec[:, :n_targ]
    res = a @ x - x * eig[:n_targ]
    # the drivers test the RECURRENCE residual (aspace*C - theta*space*C); the true residual
    # recomputed here may exceed it by rounding, hence the factor 2
    assert (np.linalg.norm(res, axis=0) / np.sqrt(n)).max() < 2 * tol
    assert np.abs(res).max() < 2 * 10 * tol
    assert np.abs(x.T @ x - np.eye(n_targ)).max() < 1e-11


def noisy_unit_guess(csr, n_max, eps=0.1):
    """lowest-diagonal unit vectors (guess_evec(1), main.f90:1337-1347) plus eps relative noise"""
    n = len(csr[0]) - 1
    return np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (eps / np.sqrt(n / 12.0)))


def run_both(D, oracle, driver, csr, n_targ, n_max, tol=1e-8, max_iter=200, max_dav=20, shift=0.0, seed=1, guess=None):
    n = len(csr[0]) - 1
    install(D, oracle, csr)
    ev_o = P.guess(n, n_max, seed=seed) if guess is None else guess.copy(order="F")
    ev_g = ev_o.copy(order="F")
    eig_g = np.zeros(n_max)
    if driver == "lobpcg":
        ro = oracle.lobpcg(ev_o, n_targ, max_iter, tol, shift=shift)
        ok = D.lobpcg_driver(False, False, n, n_targ, n_max, max_iter, tol, shift, None, None, None, eig_g, ev_g)
    else:
        ro = oracle.davidson(ev_o, n_targ, max_iter, tol, max_dav, shift=shift)
        ok = D.davidson_driver(False, n, n_targ, n_max, max_iter, tol, max_dav, shift, None, None, eig_g, ev_g)
    hg = D.last_history(n_max)
    return ro, ok, eig_g, ev_g, hg, ev_o


def assert_parity(ro, ok, eig_g, hg, n_targ, it_slack=1):
    print(f"iterations: gpu {len(hg['it'])} oracle {len(ro['it'])}; ok gpu {ok} oracle {ro['ok']}")
    assert ok == ro["ok"]
    scale = np.abs(ro["eig"][:n_targ]).max()
    assert np.abs(eig_g[:n_targ] - ro["eig"][:n_targ]).max() / scale < REL
    assert abs(len(hg["it"]) - len(ro["it"])) <= it_slack


def assert_history(ro, hg, n_targ, upto=None):
    """per-iteration Ritz values agree while both runs are in their common prefix"""
    L = min(len(hg["it"]), len(ro["it"])) if upto is None else upto
    a, b = hg["eig"][:L, :n_targ], ro["hist_eig"][:L, :n_targ]
    assert (np.abs(a - b) / np.abs(b)).max() < 1e-4
    # the first iterations are still in lock-step: rounding-level differences (summation order of the
    # Gram kernels, the Cholesky's division order) grow ~10x per iteration from a random start and
    # sit at 0.3-1.3e-6 after three iterations across builds of this library
    assert (np.abs(a[:3] - b[:3]) / np.abs(b[:3])).max() < 1e-5
    assert (np.abs(a[:1] - b[:1]) / np.abs(b[:1])).max() < 1e-8


@pytest.mark.parametrize("driver", ["lobpcg", "davidson"])
def test_c1_toy_matrix(gpu_lib, oracle, driver):
    """config C1 = the reference's own test (main.f90:14-18, 283-401)"""
    n, n_want, tol = 1000, 10, 1e-8
    n_eig = P.n_eig_rule(n_want)
    csr = dense_as_csr(P.toy_dense(n))
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, driver, csr, n_want, n_eig, tol=tol, max_iter=100, max_dav=20)
    assert ok
    assert_parity(ro, ok, eig_g, hg, n_want)
    gold = np.array(json.load(open(os.path.join(GOLD, "toy_dense_eigs.json")))["eig"])[:n_want]
    assert np.abs(eig_g[:n_want] - gold).max() / gold.max() < REL
    check_solution(csr, eig_g, ev_g, n_want, tol)
    hist = json.load(open(os.path.join(GOLD, "c1_oracle_history.json")))[driver]
    assert abs(len(hg["it"]) - hist["iterations"]) <= 1
    st = gpu_lib.last_stats()
    assert st["qr_fallbacks"] == 0 and st["launches"] > 0


@pytest.mark.parametrize("driver", ["lobpcg", "davidson"])
@pytest.mark.parametrize("gen", ["toy_sparse", "lap3d", "lap3d_32roots", "fci_like"])
def test_sparse_configs_small(gpu_lib, oracle, driver, gen):
    """scaled-down C2 / C3 / C4 with well-conditioned starts: iteration count within +-1.
    The configurations are the ones on which the oracle's own count does not move when the
    guess is scaled by (1 +- 1e-14) ('iteration-count parity')."""
    guess, slack = None, 1
    if gen == "toy_sparse":      # C2: random guess, as the reference's own test (guess_evec(4))
        csr, n_targ = P.toy_sparse(1 << 14), 8
    elif gen == "lap3d":         # C3 generator, delta=1, lowest-diagonal start with 3% noise
        csr, n_targ = P.lap3d(32, 32, 16, delta=1.0), 6
        guess = noisy_unit_guess(csr, P.n_eig_rule(n_targ), eps=0.03)
    elif gen == "lap3d_32roots":  # C3 at the bench's block size (32 roots of 37) and 10% noise
        csr, n_targ = P.lap3d(32, 32, 16, delta=1.0), 32
        guess = noisy_unit_guess(csr, P.n_eig_rule(n_targ), eps=0.1)
        slack = 2                 # the oracle itself gives 21 or 22 under 1e-14 perturbations
    else:                        # C4: lowest-diagonal start (guess_evec(1)), the FCI practice
        csr, n_targ = P.fci_like(1 << 14, n_strides=12, bandwidth=1 << 10, big_delta=0.01), 4
        guess = P.guess_lowest_diag(csr[3], P.n_eig_rule(n_targ))
    n_max = P.n_eig_rule(n_targ)
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, driver, csr, n_targ, n_max, max_iter=400, max_dav=12,
                                          guess=guess)
    assert ok and ro["ok"]
    assert_parity(ro, ok, eig_g, hg, n_targ, it_slack=slack)
    check_solution(csr, eig_g, ev_g, n_targ, 1e-8)
    # per-iteration Ritz values: whole run for the strict configs; for the 32-root case the
    # not-yet-converged upper roots wander at the 1e-2 level between two oracle runs as well
    assert_history(ro, hg, n_targ, upto=5 if gen == "lap3d_32roots" else None)


@pytest.mark.parametrize("driver", ["lobpcg", "davidson"])
def test_many_roots_c5_scaled(gpu_lib, oracle, driver):
    """scaled-down C5 (many-root stress): 128 roots of 133.  LOBPCG reduced problems reach
    len_u = 399, Davidson lda = 1330 with a restart: exercises the multi-CTA reduced eigensolver,
    ortho_cd at m = 133 (in-place trmm in column blocks) and block products wider than one tile."""
    csr = P.toy_sparse(1 << 13)
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, driver, csr, 128, 133, max_iter=200, max_dav=10)
    assert ok and ro["ok"]
    assert_parity(ro, ok, eig_g, hg, 128)
    check_solution(csr, eig_g, ev_g, 128, 1e-8)


def test_c2_full_size(gpu_lib, oracle):
    """config C2 at its full size: sparsified toy matrix, n = 2^20, 8 roots of 13, LOBPCG"""
    n = 1 << 20
    csr = P.toy_sparse(n)
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, "lobpcg", csr, 8, 13, max_iter=100)
    assert ok and ro["ok"]
    assert_parity(ro, ok, eig_g, hg, 8)
    check_solution(csr, eig_g, ev_g, 8, 1e-8)
    # (no per-iteration comparison: the lowest Ritz value drops from 4.8e5 to 5.3 in one iteration
    #  at this size, a transient that amplifies rounding differences to 1e-3 before both runs
    #  converge to the same eigenvalues in the same number of iterations)


def test_slow_random_start_long_run(gpu_lib, oracle):
    """A hard start (random guess on a disordered Laplacian) needs ~100 iterations and its
    iteration count is chaotic: the ORACLE ITSELF moves by +-15% when the guess is scaled by
    (1 +- 1e-14) ('iteration-count parity').  Eigenvalues and residuals must still
    agree; the count is only required to stay inside that spread."""
    csr = P.lap3d(16, 16, 16, delta=0.1)
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, "lobpcg", csr, 4, 9, max_iter=400)
    assert ok and ro["ok"]
    scale = np.abs(ro["eig"][:4]).max()
    assert np.abs(eig_g[:4] - ro["eig"][:4]).max() / scale < REL
    assert abs(len(hg["it"]) - len(ro["it"])) <= 0.3 * len(ro["it"])
    check_solution(csr, eig_g, ev_g, 4, 1e-8)
    assert_history(ro, hg, 4, upto=10)


def test_lobpcg_shift_returned_in_eig(gpu_lib, oracle):
    """quirk: LOBPCG returns eig INCLUDING the shift (diaglib.f90:416 vs 461)"""
    csr = P.toy_sparse(3000)
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, "lobpcg", csr, 4, 8, shift=2.5)
    assert ok
    assert_parity(ro, ok, eig_g, hg, 4)
    check_solution(csr, eig_g - 2.5, ev_g, 4, 1e-8)


def test_lobpcg_ntarg_equals_nmax_and_odd_n(gpu_lib, oracle):
    csr = P.toy_sparse(2501)  # odd n: exercises the 8-byte loaders
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, "lobpcg", csr, 5, 5, max_iter=300)
    assert_parity(ro, ok, eig_g, hg, 5)
    if ok:
        check_solution(csr, eig_g, ev_g, 5, 1e-8)


def test_not_converged_returns_ok_false(gpu_lib, oracle):
    csr = P.toy_sparse(4000)
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, "lobpcg", csr, 6, 11, max_iter=3)
    assert not ok and not ro["ok"] and len(hg["it"]) == 3
    assert np.abs(eig_g[:6] - ro["eig"][:6]).max() < 1e-6 * np.abs(ro["eig"][:6]).max()


def test_davidson_restart_path(gpu_lib, oracle):
    """max_dav below min_dav=10 is raised to 10 (1595); 19 iterations force a restart (1795-1825)
    including the skipped matvecs of the locked roots (quirk 7, 1685/1817-1824)"""
    csr = P.toy_sparse(4096)
    ro, ok, eig_g, ev_g, hg, _ = run_both(gpu_lib, oracle, "davidson", csr, 10, 10, tol=1e-10, max_iter=200, max_dav=5)
    assert len(ro["it"]) > 11, "problem too easy to exercise the restart"
    assert ok and ro["ok"]
    assert_parity(ro, ok, eig_g, hg, 10)
    check_solution(csr, eig_g, ev_g, 10, 1e-10)
    assert_history(ro, hg, 10)
    assert np.array_equal(hg["n_act"], ro["n_act"][:len(hg["n_act"])])


def test_zero_guess_makes_random_start(gpu_lib):
    """check_guess: all-zero evec -> library-generated start vectors (diaglib.f90:3750-3755)"""
    csr = P.toy_sparse(2000)
    gpu_lib.set_csr(*csr)
    n, n_max = 2000, 9
    ev = np.zeros((n, n_max), order="F")
    eig = np.zeros(n_max)
    ok = gpu_lib.lobpcg_driver(False, False, n, 4, n_max, 300, 1e-8, 0.0, None, None, None, eig, ev)
    assert ok
    check_solution(csr, eig, ev, 4, 1e-8)


def test_device_resident_evec(gpu_lib, oracle):
    """evec/eig may live in HBM (detected): same result as with host buffers"""
    from diaglib_b200 import kernels as K
    csr = P.toy_sparse(3000)
    ro, ok, eig_h, ev_h, hg, _ = run_both(gpu_lib, oracle, "lobpcg", csr, 4, 9)
    dev = K.DeviceArray.from_numpy(P.guess(3000, 9))
    eig = np.zeros(9)
    ok2 = gpu_lib.lobpcg_driver(False, False, 3000, 4, 9, 200, 1e-8, 0.0, None, None, None, eig, dev)
    assert ok and ok2
    assert np.array_equal(eig, eig_h)
    assert np.array_equal(dev.numpy(), ev_h)


def test_gen_eig_without_metric_fails_loudly(gpu_lib):
    """gen_eig=.true. with the built-in bvec but no metric installed for this n"""
    gpu_lib.set_csr(*P.toy_sparse(100))
    ev = P.guess(100, 4)
    with pytest.raises(gpu_lib.DiaglibError):
        gpu_lib.lobpcg_driver(False, True, 100, 2, 4, 10, 1e-8, 0.0, None, None, None, np.zeros(4), ev)


EDGE = [(64, 1, 1, "lobpcg", 0), (64, 1, 2, "lobpcg", 0), (50, 3, 5, "lobpcg", 0), (33, 2, 4, "davidson", 10),
        (64, 1, 1, "davidson", 10), (257, 4, 4, "lobpcg", 0), (1000, 20, 25, "davidson", 3), (129, 5, 10, "lobpcg", 0),
        (1001, 45, 50, "lobpcg", 0), (513, 7, 12, "gen_david", 10), (300, 3, 8, "gen_eig", 0)]


@pytest.mark.parametrize("n,n_targ,n_max,driver,max_dav", EDGE)
def test_small_and_odd_sizes(gpu_lib, oracle, n, n_targ, n_max, driver, max_dav):
    """n below every tile size, n_max = 1, n_targ = n_max, odd n: same eigenvalues AND the same
    iteration count as the oracle on the reference's dense toy matrix"""
    csr = dense_as_csr(P.toy_dense(n))
    install(gpu_lib, oracle, csr)
    bcsr = (np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32), 1.0 + 0.2 * np.cos(np.arange(n)))
    oracle.set_csr_b(*bcsr)
    gpu_lib.set_csr_b(*bcsr)
    ev_g = P.guess(n, n_max)
    ev_o = ev_g.copy(order="F")
    eig = np.zeros(n_max)
    if driver == "lobpcg":
        ro = oracle.lobpcg(ev_o, n_targ, 200, 1e-8)
        ok = gpu_lib.lobpcg_driver(False, False, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, ev_g)
    elif driver == "gen_eig":
        ro = oracle.lobpcg(ev_o, n_targ, 200, 1e-8, gen_eig=True)
        ok = gpu_lib.lobpcg_driver(False, True, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, ev_g)
    elif driver == "gen_david":
        ro = oracle.gen_david(ev_o, n_targ, 200, 1e-8, max_dav)
        ok = gpu_lib.gen_david_driver(False, n, n_targ, n_max, 200, 1e-8, max_dav, 0.0, None, None, None, eig, ev_g)
    else:
        ro = oracle.davidson(ev_o, n_targ, 200, 1e-8, max_dav)
        ok = gpu_lib.davidson_driver(False, n, n_targ, n_max, 200, 1e-8, max_dav, 0.0, None, None, eig, ev_g)
    hg = gpu_lib.last_history(n_max)
    assert ok and ro["ok"]
    assert np.abs(eig[:n_targ] - ro["eig"][:n_targ]).max() / np.abs(ro["eig"][:n_targ]).max() < REL
    assert abs(len(hg["it"]) - len(ro["it"])) <= 1


@pytest.mark.parametrize("n,n_max,driver", [(2, 1, "lobpcg"), (16, 3, "davidson")])
def test_search_space_larger_than_n_fails_like_the_reference(gpu_lib, oracle, n, n_max, driver):
    """3*n_max > n (LOBPCG) / dim_dav*n_max > n (Davidson): the expansion block cannot be made
    orthogonal to the space; the reference stops in ortho_vs_x (diaglib.f90:3568), so do both sides"""
    csr = dense_as_csr(P.toy_dense(n))
    install(gpu_lib, oracle, csr)
    ev_o = P.guess(n, n_max)
    ev_g = ev_o.copy(order="F")
    ro = oracle.lobpcg(ev_o, 1, 50, 1e-8) if driver == "lobpcg" else oracle.davidson(ev_o, 1, 50, 1e-8, 10)
    assert ro["status"] == 4
    with pytest.raises(gpu_lib.DiaglibError):
        if driver == "lobpcg":
            gpu_lib.lobpcg_driver(False, False, n, 1, n_max, 50, 1e-8, 0.0, None, None, None, np.zeros(n_max), ev_g)
        else:
            gpu_lib.davidson_driver(False, n, 1, n_max, 50, 1e-8, 10, 0.0, None, None, np.zeros(n_max), ev_g)


# ---- generalized problem (gen_eig branch, diaglib.f90:299-302, 357-364, 422-436, 500-526) ------
def check_gen_solution(csr, bcsr, eig, evec, n_targ, tol):
    import scipy.sparse as sp
    n = len(csr[0]) - 1
    a = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(n, n))
    b = sp.csr_matrix((bcsr[2], bcsr[1], bcsr[0]), shape=(n, n))
    x = evec[:, :n_targ]
    res = a @ x - (b @ x) * eig[:n_targ]
    assert (np.linalg.norm(res, axis=0) / np.sqrt(n)).max() < 2 * tol
    assert np.abs(x.T @ (b @ x) - np.eye(n_targ)).max() < 1e-10   # B-orthonormal Ritz vectors


@pytest.mark.parametrize("gen", ["toy_sparse", "lap3d"])
def test_gen_eig_lobpcg_vs_oracle(gpu_lib, oracle, gen):
    if gen == "toy_sparse":
        csr = P.toy_sparse(4096)
        n_targ, n_max, guess = 5, 10, None
    else:
        csr = P.lap3d(32, 16, 16, delta=1.0)
        n_targ, n_max = 6, 11
        guess = noisy_unit_guess(csr, n_max, eps=0.03)
    n = len(csr[0]) - 1
    bcsr = P.metric_like(csr)
    install(gpu_lib, oracle, csr)
    oracle.set_csr_b(*bcsr)
    gpu_lib.set_csr_b(*bcsr)
    ev_o = P.guess(n, n_max) if guess is None else guess.copy(order="F")
    ev_g = ev_o.copy(order="F")
    eig_g = np.zeros(n_max)
    ro = oracle.lobpcg(ev_o, n_targ, 300, 1e-8, gen_eig=True)
    ok = gpu_lib.lobpcg_driver(False, True, n, n_targ, n_max, 300, 1e-8, 0.0, None, None, None, eig_g, ev_g)
    hg = gpu_lib.last_history(n_max)
    assert ok and ro["ok"]
    assert_parity(ro, ok, eig_g, hg, n_targ)
    assert_history(ro, hg, n_targ, upto=min(3, len(hg["it"])))
    check_gen_solution(csr, bcsr, eig_g, ev_g, n_targ, 1e-8)
    # against dense LAPACK on the pencil (A, B)
    if n <= 5000:
        import scipy.linalg as sl
        import scipy.sparse as sp
        a = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(n, n)).toarray()
        b = sp.csr_matrix((bcsr[2], bcsr[1], bcsr[0]), shape=(n, n)).toarray()
        w = sl.eigh(a, b, eigvals_only=True, subset_by_index=[0, n_targ - 1])
        assert np.abs(eig_g[:n_targ] - w).max() / np.abs(w).max() < REL


def test_gen_eig_identity_metric_matches_standard(gpu_lib, oracle):
    """B = I: the generalized branch must reproduce the standard one"""
    csr = P.toy_sparse(3000)
    n, n_targ, n_max = 3000, 4, 9
    ident = (np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32), np.ones(n))
    gpu_lib.set_csr(*csr)
    gpu_lib.set_csr_b(*ident)
    ev_s, ev_g = P.guess(n, n_max), P.guess(n, n_max)
    eig_s, eig_g = np.zeros(n_max), np.zeros(n_max)
    ok_s = gpu_lib.lobpcg_driver(False, False, n, n_targ, n_max, 300, 1e-8, 0.0, None, None, None, eig_s, ev_s)
    its_s = len(gpu_lib.last_history(n_max)["it"])
    ok_g = gpu_lib.lobpcg_driver(False, True, n, n_targ, n_max, 300, 1e-8, 0.0, None, None, None, eig_g, ev_g)
    its_g = len(gpu_lib.last_history(n_max)["it"])
    assert ok_s and ok_g
    assert np.abs(eig_s[:n_targ] - eig_g[:n_targ]).max() / np.abs(eig_s[:n_targ]).max() < REL
    assert abs(its_s - its_g) <= 1


def test_gen_eig_user_bvec_callback(gpu_lib, oracle):
    """caller-written bvec(n,m,x,bx) on the library stream (contract diaglib.f90:206)"""
    import torch
    csr = P.toy_sparse(2048)
    n, n_targ, n_max = 2048, 4, 9
    bcsr = P.metric_like(csr)
    gpu_lib.set_csr(*csr)
    gpu_lib.set_csr_b(*bcsr)
    ev_b, eig_b = P.guess(n, n_max), np.zeros(n_max)
    assert gpu_lib.lobpcg_driver(False, True, n, n_targ, n_max, 300, 1e-8, 0.0, None, None, None, eig_b, ev_b)
    import scipy.sparse as sp
    b_t = torch.as_tensor(sp.csr_matrix((bcsr[2], bcsr[1], bcsr[0]), shape=(n, n)).toarray(), device="cuda:0")
    stream = torch.cuda.ExternalStream(gpu_lib.lib().diaglib_b200_stream(), device="cuda:0")
    calls = [0]

    def bvec(nn, m, x, bx):
        calls[0] += 1
        with torch.cuda.stream(stream):
            xt = torch.as_tensor(_DevView(x, nn, m), device="cuda:0")
            torch.as_tensor(_DevView(bx, nn, m), device="cuda:0").copy_(xt @ b_t)   # B symmetric

    ev, eig = P.guess(n, n_max), np.zeros(n_max)
    assert gpu_lib.lobpcg_driver(False, True, n, n_targ, n_max, 300, 1e-8, 0.0, None, None, bvec, eig, ev)
    assert calls[0] >= 2
    assert np.abs(eig[:n_targ] - eig_b[:n_targ]).max() / np.abs(eig_b[:n_targ]).max() < REL
    check_gen_solution(csr, bcsr, eig, ev, n_targ, 1e-8)


@pytest.mark.parametrize("case", ["no_restart", "restart"])
def test_gen_david_vs_oracle(gpu_lib, oracle, case):
    """gen_david_driver (diaglib.f90:1855-2250) against the oracle's restatement; the restart case
    runs past dim_dav = 10 expansions (2188-2222)"""
    if case == "no_restart":
        csr = P.lap3d(32, 16, 16, delta=1.0)
        n_targ, n_max, tol = 6, 11, 1e-8
        guess = noisy_unit_guess(csr, n_max, eps=0.03)
    else:
        csr = P.toy_sparse(4096)
        n_targ, n_max, tol = 10, 15, 1e-10
        guess = None
    n = len(csr[0]) - 1
    bcsr = P.metric_like(csr)
    install(gpu_lib, oracle, csr)
    oracle.set_csr_b(*bcsr)
    gpu_lib.set_csr_b(*bcsr)
    ev_o = P.guess(n, n_max) if guess is None else guess.copy(order="F")
    ev_g = ev_o.copy(order="F")
    eig_g = np.zeros(n_max)
    ro = oracle.gen_david(ev_o, n_targ, 100, tol, 10)
    ok = gpu_lib.gen_david_driver(False, n, n_targ, n_max, 100, tol, 10, 0.0, None, None, None, eig_g, ev_g)
    hg = gpu_lib.last_history(n_max)
    assert ok and ro["ok"]
    assert (len(hg["it"]) > 10) == (case == "restart")
    assert_parity(ro, ok, eig_g, hg, n_targ)
    assert np.array_equal(hg["n_act"], ro["n_act"][:len(hg["n_act"])])
    check_gen_solution(csr, bcsr, eig_g, ev_g, n_targ, tol)
    import scipy.linalg as sl
    import scipy.sparse as sp
    a = sp.csr_matrix((csr[2], csr[1], csr[0]), shape=(n, n)).toarray()
    b = sp.csr_matrix((bcsr[2], bcsr[1], bcsr[0]), shape=(n, n)).toarray()
    w = sl.eigh(a, b, eigvals_only=True, subset_by_index=[0, n_targ - 1])
    assert np.abs(eig_g[:n_targ] - w).max() / np.abs(w).max() < REL


def test_gen_david_identity_metric_matches_davidson(gpu_lib):
    csr = P.toy_sparse(3000)
    n, n_targ, n_max = 3000, 4, 9
    ident = (np.arange(n + 1, dtype=np.int64), np.arange(n, dtype=np.int32), np.ones(n))
    gpu_lib.set_csr(*csr)
    gpu_lib.set_csr_b(*ident)
    ev_s, ev_g = P.guess(n, n_max), P.guess(n, n_max)
    eig_s, eig_g = np.zeros(n_max), np.zeros(n_max)
    assert gpu_lib.davidson_driver(False, n, n_targ, n_max, 100, 1e-8, 10, 0.0, None, None, eig_s, ev_s)
    its_s = len(gpu_lib.last_history(n_max)["it"])
    assert gpu_lib.gen_david_driver(False, n, n_targ, n_max, 100, 1e-8, 10, 0.0, None, None, None, eig_g, ev_g)
    its_g = len(gpu_lib.last_history(n_max)["it"])
    assert np.abs(eig_s[:n_targ] - eig_g[:n_targ]).max() / np.abs(eig_s[:n_targ]).max() < REL
    assert abs(its_s - its_g) <= 1


# ---- linear-response solver (caslr_eff_driver, diaglib.f90:1024-1481) ------------------------------
def _lr_dense(lr, n):
    import scipy.sparse as sp
    m = {k: sp.csr_matrix((lr[k][2], lr[k][1], lr[k][0]), shape=(n, n)).toarray() for k in ("apb", "amb", "spd", "smd")}
    a, b = 0.5 * (m["apb"] + m["amb"]), 0.5 * (m["apb"] - m["amb"])
    sg, dl = 0.5 * (m["spd"] + m["smd"]), 0.5 * (m["spd"] - m["smd"])
    return np.block([[a, b], [b, a]]), np.block([[sg, dl], [-dl, -sg]])


def _lr_guess(lr, n, n_max, eps=0.05):
    """Y = unit vectors on the lowest diag(A)/diag(S) sites, Z = 0, plus eps relative noise"""
    ev = np.zeros((2 * n, n_max), order="F")
    ev[:n] = P.guess_lowest_diag(lr["aa_diag"] / lr["sigma_diag"], n_max)
    return np.asfortranarray(ev + P.guess(2 * n, n_max) * (eps / np.sqrt(2 * n / 12.0)))


@pytest.mark.parametrize("n,n_targ,tol,structured", [(1024, 4, 1e-8, False), (4096, 8, 1e-9, True), (20001, 8, 1e-9, True)])
def test_caslr_eff_vs_oracle(gpu_lib, oracle, n, n_targ, tol, structured):
    """random guess: runs through restarts (1422-1457); structured guess: converges inside one cycle"""
    lr = P.caslr_like(n)
    n_max = P.n_eig_rule(n_targ)
    args = (lr["apb"], lr["amb"], lr["spd"], lr["smd"], lr["aa_diag"], lr["sigma_diag"])
    oracle.set_lr(*args)
    gpu_lib.set_lr(*args)
    ev_o = _lr_guess(lr, n, n_max) if structured else P.guess(2 * n, n_max)
    ev_g = ev_o.copy(order="F")
    eig_g = np.zeros(n_max)
    ro = oracle.caslr_eff(ev_o, n_targ, 200, tol, 10)
    ok = gpu_lib.caslr_eff_driver(False, n, 2 * n, n_targ, n_max, 200, tol, 10, None, None, None, None, None, eig_g, ev_g)
    hg = gpu_lib.last_history(n_max)
    assert ok and ro["ok"]
    assert (len(hg["it"]) > 10) == (not structured)
    assert_parity(ro, ok, eig_g, hg, n_targ, it_slack=1 if structured else 2)
    assert np.array_equal(hg["n_act"][:8], ro["n_act"][:8])
    # the returned pairs solve A_full x = w S_full x with x = [Y; Z]
    af, sf = _lr_dense(lr, n)
    x = ev_g[:, :n_targ]
    res = af @ x - (sf @ x) * eig_g[:n_targ]
    assert (np.linalg.norm(res, axis=0) / np.linalg.norm(af @ x, axis=0)).max() < 50 * tol
    if n <= 2048:
        import scipy.linalg as sl
        w = sl.eigh(sf, af, eigvals_only=True)            # S x = (1/w) A x
        ref = np.sort(1.0 / w[w > 0])[:n_targ]
        assert np.abs(eig_g[:n_targ] - ref).max() / ref.max() < REL


def test_caslr_eff_user_callbacks(gpu_lib):
    """caller-written products and lrprec (torch on the library stream) reproduce the built-ins"""
    import torch
    n, n_targ = 1024, 4
    n_max = P.n_eig_rule(n_targ)
    lr = P.caslr_like(n)
    gpu_lib.set_lr(lr["apb"], lr["amb"], lr["spd"], lr["smd"], lr["aa_diag"], lr["sigma_diag"])
    ev_b, eig_b = P.guess(2 * n, n_max), np.zeros(n_max)
    assert gpu_lib.caslr_eff_driver(False, n, 2 * n, n_targ, n_max, 200, 1e-8, 10, None, None, None, None, None, eig_b, ev_b)
    import scipy.sparse as sp
    stream = torch.cuda.ExternalStream(gpu_lib.lib().diaglib_b200_stream(), device="cuda:0")
    mats = {k: torch.as_tensor(sp.csr_matrix((lr[k][2], lr[k][1], lr[k][0]), shape=(n, n)).toarray(), device="cuda:0")
            for k in ("apb", "amb", "spd", "smd")}
    aa, sg = torch.as_tensor(lr["aa_diag"], device="cuda:0"), torch.as_tensor(lr["sigma_diag"], device="cuda:0")

    def product(k):
        def f(nn, m, x, y):
            with torch.cuda.stream(stream):
                xt = torch.as_tensor(_DevView(x, nn, m), device="cuda:0")          # X^T
                torch.as_tensor(_DevView(y, nn, m), device="cuda:0").copy_(xt @ mats[k].T)
        return f

    def lrprec(nn, m, fac, xp, xm, yp, ym):
        with torch.cuda.stream(stream):
            p = torch.as_tensor(_DevView(xp, nn, m), device="cuda:0")
            q = torch.as_tensor(_DevView(xm, nn, m), device="cuda:0")
            den = 1.0 / (fac * fac * aa * aa - sg * sg)
            torch.as_tensor(_DevView(yp, nn, m), device="cuda:0").copy_(den * (fac * aa * p + sg * q))
            torch.as_tensor(_DevView(ym, nn, m), device="cuda:0").copy_(den * (fac * aa * q + sg * p))

    ev, eig = P.guess(2 * n, n_max), np.zeros(n_max)
    assert gpu_lib.caslr_eff_driver(False, n, 2 * n, n_targ, n_max, 200, 1e-8, 10, product("apb"), product("amb"),
                                    product("spd"), product("smd"), lrprec, eig, ev)
    assert np.abs(eig[:n_targ] - eig_b[:n_targ]).max() / np.abs(eig_b[:n_targ]).max() < REL


@pytest.mark.parametrize("n,m", [(3000, 12), (20000, 37)])
def test_b_ortho_vs_oracle(gpu_lib, oracle, n, m):
    csr = P.toy_sparse(n)
    bcsr = P.metric_like(csr)
    oracle.set_csr_b(*bcsr)
    rng = np.random.default_rng(m)
    u, _ = np.linalg.qr(rng.standard_normal((n, m)))
    u = np.asfortranarray(u)
    bu = oracle.csr_bvec(u)
    uo, buo = u.copy(order="F"), bu.copy(order="F")
    oracle.b_ortho(uo, buo)
    gpu_lib.b_ortho(n, m, u, bu)
    assert np.abs(u.T @ bu - np.eye(m)).max() < 1e-13 * m
    assert np.abs(u - uo).max() < 1e-12 and np.abs(bu - buo).max() < 1e-12


@pytest.mark.parametrize("n,m,k", [(3000, 12, 12), (20000, 74, 37)])
def test_b_ortho_vs_x_vs_oracle(gpu_lib, oracle, n, m, k):
    csr = P.toy_sparse(n)
    bcsr = P.metric_like(csr)
    oracle.set_csr_b(*bcsr)
    rng = np.random.default_rng(k)
    x, _ = np.linalg.qr(rng.standard_normal((n, m)))
    x = np.asfortranarray(x)
    bx = oracle.csr_bvec(x)
    oracle.b_ortho(x, bx)                          # x B-orthonormal, bx = B x
    u = np.asfortranarray(rng.standard_normal((n, k)) + 30.0 * x[:, :k])
    uo = u.copy(order="F")
    oracle.b_ortho_vs_x(x, bx, uo)
    gpu_lib.b_ortho_vs_x(n, m, k, x, bx, u)
    assert np.linalg.norm(bx.T @ u) < 1e-13 * np.sqrt(m * k)
    assert np.linalg.norm(u.T @ u - np.eye(k)) < 1e-13 * k
    assert np.abs(u - uo).max() < 1e-9


class _DevView:
    """zero-copy torch view of a column-major (n, m) device block handed to a callback"""

    def __init__(self, ptr, n, m):
        self.__cuda_array_interface__ = dict(shape=(m, n), typestr="<f8", data=(int(ptr), False), version=2)


@pytest.mark.parametrize("driver", ["lobpcg", "davidson"])
def test_user_supplied_device_callbacks(gpu_lib, oracle, driver):
    """the reference contract matvec(n,m,x,ax) / precnd(n,m,shift,x,ax) (diaglib.f90:62-72) with
    caller-written device-aware callbacks: here torch ops enqueued on diaglib_b200_stream()"""
    import torch
    n, n_want, tol = 1000, 10, 1e-8
    n_eig = P.n_eig_rule(n_want)
    a = P.toy_dense(n)
    csr = dense_as_csr(a)
    ro, ok_b, eig_b, ev_b, hb, _ = run_both(gpu_lib, oracle, driver, csr, n_want, n_eig, tol=tol, max_iter=100)
    stream = torch.cuda.ExternalStream(gpu_lib.lib().diaglib_b200_stream(), device="cuda:0")
    a_t = torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    d_t = torch.as_tensor(np.diag(a).copy(), device="cuda:0")
    calls = dict(mv=0, pc=0, cols=0)

    def matvec(nn, m, x, ax):
        calls["mv"] += 1
        calls["cols"] += m
        with torch.cuda.stream(stream):
            xt = torch.as_tensor(_DevView(x, nn, m), device="cuda:0")      # = X^T
            torch.as_tensor(_DevView(ax, nn, m), device="cuda:0").copy_(xt @ a_t)   # A symmetric

    def precnd(nn, m, shift, x, px):
        calls["pc"] += 1
        with torch.cuda.stream(stream):
            xt = torch.as_tensor(_DevView(x, nn, m), device="cuda:0")
            den = d_t + shift
            den = torch.where(den.abs() > 1e-5, den, torch.ones_like(den))  # main.f90:146-171
            torch.as_tensor(_DevView(px, nn, m), device="cuda:0").copy_(xt / den)

    ev = P.guess(n, n_eig)
    eig = np.zeros(n_eig)
    if driver == "lobpcg":
        ok = gpu_lib.lobpcg_driver(False, False, n, n_want, n_eig, 100, tol, 0.0, matvec, precnd, None, eig, ev)
    else:
        ok = gpu_lib.davidson_driver(False, n, n_want, n_eig, 100, tol, 20, 0.0, matvec, precnd, eig, ev)
    hg = gpu_lib.last_history(n_eig)
    assert ok and ok_b
    assert calls["mv"] >= len(hg["it"])
    assert calls["pc"] >= len(hg["it"]) - 1
    assert np.abs(eig[:n_want] - eig_b[:n_want]).max() / np.abs(eig_b[:n_want]).max() < REL
    assert np.abs(eig[:n_want] - ro["eig"][:n_want]).max() / np.abs(ro["eig"][:n_want]).max() < REL
    assert abs(len(hg["it"]) - len(ro["it"])) <= 1
    check_solution(csr, eig, ev, n_want, tol)


@pytest.mark.parametrize("driver", ["lobpcg", "davidson"])
def test_verbose_prints_iteration_table(gpu_lib, driver, capfd):
    """verbose=.true. prints the per-iteration table of the reference (diaglib.f90:459-464,
    1741-1749) and the timing summary (537-552)"""
    csr = P.toy_sparse(2000)
    gpu_lib.set_csr(*csr)
    n, n_targ, n_max = 2000, 3, 6
    ev = P.guess(n, n_max)
    eig = np.zeros(n_max)
    if driver == "lobpcg":
        ok = gpu_lib.lobpcg_driver(True, False, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, ev)
    else:
        ok = gpu_lib.davidson_driver(True, n, n_targ, n_max, 200, 1e-8, 10, 0.0, None, None, eig, ev)
    out = capfd.readouterr().out
    assert ok
    its = len(gpu_lib.last_history(n_max)["it"])
    rows = [ln.split() for ln in out.splitlines() if len(ln.split()) == 6 and ln.split()[-1] in ("T", "F")]
    assert len(rows) == its * n_targ
    last = rows[-n_targ:]
    assert all(r[-1] == "T" for r in last)
    assert np.allclose([float(r[2]) for r in last], eig[:n_targ], rtol=0, atol=1e-11 * np.abs(eig[:n_targ]).max() + 1e-11)
    assert "matrix-vector multiplications" in out and "orthogonalization" in out
    # silent when verbose is false
    gpu_lib.lobpcg_driver(False, False, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, P.guess(n, n_max))
    assert capfd.readouterr().out == ""


# ---- public block routines -------------------------------------------------------------------
@pytest.mark.parametrize("n,m", [(1000, 15), (20000, 37), (5001, 8), (3000, 133)])
def test_ortho_cd_vs_oracle(gpu_lib, oracle, n, m):
    rng = np.random.default_rng(m)
    u = np.asfortranarray(rng.standard_normal((n, m)))
    u[:, m // 2] = u[:, 0] + 1e-6 * rng.standard_normal(n)
    uo = u.copy(order="F")
    go, oko = oracle.ortho_cd(uo)
    g, ok = gpu_lib.ortho_cd(n, m, u)
    assert ok and oko
    assert np.linalg.norm(u.T @ u - np.eye(m)) < 1e-13 * m
    assert abs(g - go) < 1e-3 * go  # growth ~ cond(L): sensitive to rounding at cond 1e6
    assert np.abs(u - uo).max() < 1e-7
    assert gpu_lib.last_stats()["ortho_cd_passes"] >= 2


def test_ortho_cd_rank_deficient_uses_level_shift(gpu_lib, oracle):
    rng = np.random.default_rng(0)
    u = np.asfortranarray(rng.standard_normal((4000, 10)))
    u[:, 6] = u[:, 1]
    g, ok = gpu_lib.ortho_cd(4000, 10, u)
    # whether dpotrf trips on an exactly rank-deficient Gram depends on rounding; the shift loop
    # itself is pinned in test_gpu_kernels.py::test_chol_inv_level_shift.  Either way: finite output.
    assert np.all(np.isfinite(u))


@pytest.mark.parametrize("n,m,k", [(1000, 15, 15), (20000, 74, 37), (5001, 30, 8), (4000, 300, 15)])
def test_ortho_vs_x_vs_oracle(gpu_lib, oracle, n, m, k):
    rng = np.random.default_rng(k)
    x, _ = np.linalg.qr(rng.standard_normal((n, m)))
    x = np.asfortranarray(x)
    u = np.asfortranarray(rng.standard_normal((n, k)) + 30.0 * x[:, :k])
    uo = u.copy(order="F")
    oracle.ortho_vs_x(x, uo)
    gpu_lib.ortho_vs_x(n, m, k, x, u)
    assert np.linalg.norm(x.T @ u) < 1e-13 * np.sqrt(m * k)
    assert np.linalg.norm(u.T @ u - np.eye(k)) < 1e-13 * k
    assert np.abs(u - uo).max() < 1e-9


def test_ortho_qr_fallback(gpu_lib, oracle):
    rng = np.random.default_rng(7)
    u = np.asfortranarray(rng.standard_normal((3000, 11)))
    uo = u.copy(order="F")
    oracle.ortho(uo)
    gpu_lib.ortho(3000, 11, u)
    assert np.linalg.norm(u.T @ u - np.eye(11)) < 1e-13
    # same Q up to column signs (Householder R may have negative diagonal entries)
    s = np.sign(np.sum(u * uo, axis=0))
    assert np.abs(u * s - uo).max() < 1e-10


def test_c3_bench_workload_nx128_vs_oracle_fixture(gpu_lib):
    """The benchmark workload itself (bench.py: C3, 32 roots of 37, tol 1e-8, lowest-diagonal start
    + 10 % noise) at 128^3 = 2 097 152 rows against the oracle's complete solves of the same problem
    (fixtures: the oracle needs one to two minutes each).  diaglib.f90:389-533.
      * eigenvalues 1e-10 relative against the literal reference (dsyev), residuals below tol;
      * iteration count: the dsyev oracle needs 27, the same oracle with LAPACK's accurate route for
        the reduced problem (dpotrf + dgesvj) 23 -- dsyev's eps*|a_red| eigenvector error
        (|a_red| ~ n) holds max|r| up near the 1e-7 threshold; the GPU path's Jacobi solvers do not
        have that floor.  Required: within +-1 of the accurate-eigensolver oracle and never more
        iterations than the dsyev oracle (bench.parity_block documents the same at 256^3)."""
    import bench
    gold = json.load(open(os.path.join(GOLD, "c3_oracle_nx128.json")))
    gold_acc = json.load(open(os.path.join(GOLD, "c3_oracle_acc_nx128.json")))
    nx = gold["nx"]
    n, n_targ = nx ** 3, bench.N_TARG
    n_max = P.n_eig_rule(n_targ)
    assert (gold["n_targ"], gold["tol"], gold["delta"], gold["noise"]) == (n_targ, bench.TOL, bench.DELTA, bench.NOISE)
    assert gold_acc["nx"] == nx and gold_acc["accurate_eig"] and gold_acc["ok"]
    csr = P.lap3d(nx, nx, nx, delta=bench.DELTA)
    gpu_lib.set_csr(*csr)
    ev = bench.make_guess(csr[3], n, n_max, 0, n)
    eig = np.zeros(n_max)
    ok = gpu_lib.lobpcg_driver(False, False, n, n_targ, n_max, bench.MAX_ITER, bench.TOL, 0.0, None, None, None, eig, ev)
    hg = gpu_lib.last_history(n_max)
    par = bench.parity_block(dict(gold, _source="tests/golden/c3_oracle_nx128.json"), len(hg["it"]), eig, hg["rms"][-1],
                             hg["max"][-1], gold_acc)
    print(par)
    assert ok and gold["ok"]
    assert par["max_rel_eig_err"] <= REL and par["max_rel_eig_err_vs_accurate_eig_oracle"] <= REL
    assert par["max_rms"] < bench.TOL and par["max_abs_residual"] < 10 * bench.TOL
    assert abs(par["its_gpu"] - par["its_oracle_accurate_eig"]) <= 1
    assert par["its_gpu"] <= par["its_oracle"]
    assert par["ok"]
    check_solution(csr, eig, ev, n_targ, bench.TOL)
    gpu_lib.lib().diaglib_b200_release_workspace()


@pytest.mark.parametrize("driver", ["lobpcg", "davidson"])
def test_speculative_ortho_chains_change_nothing_but_the_sync_count(gpu_lib, driver):
    """ortho_cd / ortho_vs_x decided on the device (one read-back per call) against the
    host-driven form (one per pass): same decisions, same arithmetic -> bit-identical histories
    and pass / sweep counts; at most 3 host synchronisations per LOBPCG iteration
    (diaglib.f90:3246-3333, 3533-3568)."""
    from diaglib_b200 import kernels as K
    csr = P.lap3d(32, 32, 16, delta=1.0)
    n, n_targ = 1 << 14, 6
    n_max = P.n_eig_rule(n_targ)
    gpu_lib.set_csr(*csr)
    g = noisy_unit_guess(csr, n_max, eps=0.03)
    runs = {}
    for spec in (True, False):
        prev = K.set_spec_ortho(spec)
        try:
            ev, eig = g.copy(order="F"), np.zeros(n_max)
            if driver == "lobpcg":
                ok = gpu_lib.lobpcg_driver(False, False, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, ev)
            else:
                ok = gpu_lib.davidson_driver(False, n, n_targ, n_max, 200, 1e-8, 10, 0.0, None, None, eig, ev)
            runs[spec] = (ok, eig, ev, gpu_lib.last_history(n_max), gpu_lib.last_stats())
        finally:
            K.set_spec_ortho(prev)
    (ok1, e1, v1, h1, s1), (ok0, e0, v0, h0, s0) = runs[True], runs[False]
    assert ok1 and ok0
    assert np.array_equal(h1["eig"], h0["eig"]) and np.array_equal(h1["rms"], h0["rms"]) and np.array_equal(e1, e0)
    assert np.array_equal(v1, v0)
    assert (s1["ortho_cd_passes"], s1["ortho_vs_x_sweeps"], s1["qr_fallbacks"]) == (s0["ortho_cd_passes"], s0["ortho_vs_x_sweeps"], s0["qr_fallbacks"])
    its = len(h1["it"])
    print(f"{driver}: {its} iterations, host syncs speculative {s1['host_syncs']} vs host-driven {s0['host_syncs']}")
    assert s1["host_syncs"] < s0["host_syncs"]
    if driver == "lobpcg":
        assert s1["host_syncs"] <= 3 * its + 8     # set-up + final copy on top of <= 3 per iteration


def test_gen_david_reference_restart_switch(gpu_lib, oracle):
    """DIAGLIB_B200_REFERENCE_RESTART=1 / k_set_reference_restart(1): the library executes the
    reference's literal `bspace = zero` after a restart (diaglib.f90:2200) and then reproduces the
    oracle's literal mode -- including its wrong eigenvalues; the default keeps B * restart vectors"""
    from diaglib_b200 import kernels as K
    n, n_targ, n_max = 600, 4, 9
    csr = P.toy_sparse(n)
    bcsr = P.metric_like(csr)
    oracle.set_csr(*csr)
    oracle.set_csr_b(*bcsr)
    gpu_lib.set_csr(*csr)
    gpu_lib.set_csr_b(*bcsr)
    lit = oracle.gen_david(P.guess(n, n_max), n_targ, 100, 1e-8, 10, reference_restart=True)
    good = oracle.gen_david(P.guess(n, n_max), n_targ, 100, 1e-8, 10)
    assert np.abs(lit["eig"][:n_targ] - good["eig"][:n_targ]).max() > 1e-3
    prev = K.set_reference_restart(True)
    try:
        ev, eig = P.guess(n, n_max), np.zeros(n_max)
        gpu_lib.gen_david_driver(False, n, n_targ, n_max, 100, 1e-8, 10, 0.0, None, None, None, eig, ev)
        hg = gpu_lib.last_history(n_max)
    finally:
        K.set_reference_restart(prev)
    # same (wrong) answer as the literal oracle - both collapse to ~0 here, far from the true eigenvalues -
    # and the same iteration count within 1
    assert np.abs(eig[:n_targ] - lit["eig"][:n_targ]).max() < 1e-8
    assert np.abs(eig[:n_targ] - good["eig"][:n_targ]).max() > 1e-3
    assert abs(len(hg["it"]) - len(lit["it"])) <= 1


@pytest.mark.parametrize("driver", ["davidson", "lobpcg"])
def test_c5_many_roots_n18_vs_oracle_fixture(gpu_lib, driver):
    """C5 at n = 2^18: toy_sparse, 128 roots of 133, Davidson-Liu with max_dav = 10
    (lda = 1330: ortho_cd at m = 133, reduced eigenproblems up to 1330 x 1330) and LOBPCG (len_a = 399),
    against the oracle's complete solves (tests/golden/c5_oracle_n18.json, tools/c5_oracle.py; minutes of
    CPU, hence a fixture).  diaglib.f90:1676-1828, 389-533."""
    gold = json.load(open(os.path.join(GOLD, "c5_oracle_n18.json")))
    n, n_targ, n_max, tol = gold["n"], gold["n_targ"], gold["n_max"], gold["tol"]
    csr = P.toy_sparse(n)
    gpu_lib.set_csr(*csr)
    g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (gold["noise"] / np.sqrt(n / 12.0)))
    eig = np.zeros(n_max)
    if driver == "davidson":
        ok = gpu_lib.davidson_driver(False, n, n_targ, n_max, 100, tol, gold["max_dav"], 0.0, None, None, eig, g)
    else:
        ok = gpu_lib.lobpcg_driver(False, False, n, n_targ, n_max, 100, tol, 0.0, None, None, None, eig, g)
    hg = gpu_lib.last_history(n_max)
    ref = gold[driver]
    eo = np.array(ref["eig"][:n_targ])
    rel = np.abs(eig[:n_targ] - eo).max() / np.abs(eo).max()
    print(f"C5 n=2^18 {driver}: iterations gpu {len(hg['it'])} oracle {ref['iterations']}, max rel eig err {rel:.2e}, "
          f"timers {gpu_lib.last_timers()}")
    assert ok and ref["ok"]
    assert rel < REL
    assert abs(len(hg["it"]) - ref["iterations"]) <= 1
    check_solution(csr, eig, g, n_targ, tol)
    gpu_lib.lib().diaglib_b200_release_workspace()


def test_repeated_solves_are_bit_identical(gpu_lib):
    """Run-to-run reproducibility: the same LOBPCG solve (scaled-down benchmark workload, tiled row order, speculative
    ortho chains with the deferred triangular multiply) four times, bit-identical eigenvalue history and vectors.
    Guards the epilogue of the block multiply: storing its accumulator registers directly (instead of alpha * acc in a
    register of its own) made 2-7 of 8 repetitions differ on a B200 (tools/determinism_check.py)."""
    D = gpu_lib
    nx, ny, nz, n_targ = 128, 128, 64, 32
    n = nx * ny * nz
    n_max = P.n_eig_rule(n_targ)
    csr = P.lap3d(nx, ny, nz, delta=1.0)
    D.set_csr(*csr)
    D.set_csr_row_order(P.tile_order_3d(nx, ny, nz, tile=(64, 2, 2), curve="morton"))
    try:
        g = np.asfortranarray(P.guess_lowest_diag(csr[3], n_max) + P.guess(n, n_max) * (0.1 / np.sqrt(n / 12.0)))
        ref = None
        for _ in range(4):
            ev = g.copy(order="F")
            eig = np.zeros(n_max)
            ok = D.lobpcg_driver(False, False, n, n_targ, n_max, 200, 1e-8, 0.0, None, None, None, eig, ev)
            assert ok
            got = (np.asarray(D.last_history(n_max)["eig"]).tobytes(), ev.tobytes())
            if ref is None:
                ref = got
            assert got == ref
    finally:
        D.set_csr_row_order(None)
