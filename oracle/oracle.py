"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package diaglib_b200."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MATVEC_T = C.CFUNCTYPE(None, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double))
PRECND_T = C.CFUNCTYPE(None, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                       C.POINTER(C.c_double))


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_norm_est.restype = C.c_double
        _LIB.oracle_blas_config.restype = C.c_char_p
    return _LIB


def _i(v):
    return C.byref(C.c_int32(int(v)))


def _d(v):
    return C.byref(C.c_double(float(v)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Keep:
    refs: list = []
    refs_b: list = []
    refs_lr: list = []


def set_csr(rowptr, col, val, diag):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    diag = np.ascontiguousarray(diag, dtype=np.float64)
    _Keep.refs = [rowptr, col, val, diag]
    lib().oracle_set_csr(C.c_int64(len(rowptr) - 1), _p(rowptr), _p(col), _p(val), _p(diag))


def set_csr_b(rowptr, col, val):
    """metric of the generalized problem, applied by oracle_csr_bvec"""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    _Keep.refs_b = [rowptr, col, val]
    lib().oracle_set_csr_b(C.c_int64(len(rowptr) - 1), _p(rowptr), _p(col), _p(val))


def set_dense(a, diag=None):
    a = np.asfortranarray(a, dtype=np.float64)
    diag = np.ascontiguousarray(np.diag(a) if diag is None else diag, dtype=np.float64)
    _Keep.refs = [a, diag]
    lib().oracle_set_dense(C.c_int64(a.shape[0]), _p(a), _p(diag))


def _cb(name):
    return C.cast(getattr(lib(), name), C.c_void_p)


def history():
    L = lib().oracle_history_len()
    return L


def _collect(n_max):
    L = lib().oracle_history_len()
    it = np.zeros(L, np.int32)
    n_act = np.zeros(L, np.int32)
    eig = np.zeros((L, n_max))
    rms = np.zeros((L, n_max))
    mx = np.zeros((L, n_max))
    done = np.zeros((L, n_max), np.int32)
    if L:
        lib().oracle_history_get(_p(it), _p(n_act), _p(eig), _p(rms), _p(mx), _p(done))
    t = np.zeros(4)
    lib().oracle_timers(_p(t))
    st = np.zeros(4, np.int32)
    lib().oracle_stats(_p(st))
    return dict(it=it, n_act=n_act, eig=eig, rms=rms, max=mx, done=done,
                timers=dict(mv=t[0], diag=t[1], ortho=t[2], total=t[3]),
                stats=dict(ortho_cd_passes=int(st[0]), ortho_vs_x_sweeps=int(st[1]), qr_fallbacks=int(st[2]),
                           chol_shifts=int(st[3])),
                status=int(lib().oracle_last_status()))


def lobpcg(evec, n_targ, max_iter, tol, shift=0.0, matvec="oracle_csr_matvec", precnd="oracle_diag_precnd",
           verbose=False, gen_eig=False, bvec="oracle_csr_bvec"):
    """Runs the oracle's lobpcg_driver.  evec: (n, n_max) Fortran-ordered guess, overwritten."""
    assert evec.flags.f_contiguous and evec.dtype == np.float64
    n, n_max = evec.shape
    eig = np.zeros(n_max)
    ok = C.c_int32(0)
    lib().oracle_stats_reset()
    lib().oracle_lobpcg_driver(_i(verbose), _i(1 if gen_eig else 0), _i(n), _i(n_targ), _i(n_max), _i(max_iter),
                               _d(tol), _d(shift), _cb(matvec), _cb(precnd), _cb(bvec) if gen_eig else None, _p(eig),
                               _p(evec), C.byref(ok))
    out = _collect(n_max)
    out.update(eig=eig, ok=bool(ok.value), hist_eig=out["eig"])
    return out


def davidson(evec, n_targ, max_iter, tol, max_dav, shift=0.0, matvec="oracle_csr_matvec",
             precnd="oracle_diag_precnd", verbose=False):
    assert evec.flags.f_contiguous and evec.dtype == np.float64
    n, n_max = evec.shape
    eig = np.zeros(n_max)
    ok = C.c_int32(0)
    lib().oracle_stats_reset()
    lib().oracle_davidson_driver(_i(verbose), _i(n), _i(n_targ), _i(n_max), _i(max_iter), _d(tol), _i(max_dav),
                                 _d(shift), _cb(matvec), _cb(precnd), _p(eig), _p(evec), C.byref(ok))
    out = _collect(n_max)
    out.update(eig=eig, ok=bool(ok.value), hist_eig=out["eig"])
    return out


def gen_david(evec, n_targ, max_iter, tol, max_dav, shift=0.0, matvec="oracle_csr_matvec",
              precnd="oracle_diag_precnd", bvec="oracle_csr_bvec", verbose=False, reference_restart=False):
    """gen_david_driver (diaglib.f90:1855).  reference_restart=True reproduces the reference's
    `bspace = zero` after a restart (see diaglib_oracle.cpp)."""
    assert evec.flags.f_contiguous and evec.dtype == np.float64
    n, n_max = evec.shape
    eig = np.zeros(n_max)
    ok = C.c_int32(0)
    lib().oracle_stats_reset()
    lib().oracle_set_gen_david_reference_restart(1 if reference_restart else 0)
    lib().oracle_gen_david_driver(_i(verbose), _i(n), _i(n_targ), _i(n_max), _i(max_iter), _d(tol), _i(max_dav),
                                  _d(shift), _cb(matvec), _cb(precnd), _cb(bvec), _p(eig), _p(evec), C.byref(ok))
    lib().oracle_set_gen_david_reference_restart(0)
    out = _collect(n_max)
    out.update(eig=eig, ok=bool(ok.value), hist_eig=out["eig"])
    return out


def set_lr(apb, amb, spd, smd, aa_diag, sigma_diag):
    """the four CSR matrices (rowptr, col, val) and the two diagonals of the linear-response problem"""
    keep = []
    for which, (rowptr, col, val) in enumerate((apb, amb, spd, smd)):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        keep += [rowptr, col, val]
        lib().oracle_set_csr_lr(which, _p(rowptr), _p(col), _p(val))
    aa = np.ascontiguousarray(aa_diag, dtype=np.float64)
    sg = np.ascontiguousarray(sigma_diag, dtype=np.float64)
    keep += [aa, sg]
    lib().oracle_set_lr_diag(_p(aa), _p(sg))
    _Keep.refs_lr = keep


def caslr_eff(evec, n_targ, max_iter, tol, max_dav, verbose=False):
    """caslr_eff_driver (diaglib.f90:1024).  evec: (2n, n_max) Fortran-ordered guess, overwritten."""
    assert evec.flags.f_contiguous and evec.dtype == np.float64
    n2, n_max = evec.shape
    n = n2 // 2
    eig = np.zeros(n_max)
    ok = C.c_int32(0)
    lib().oracle_stats_reset()
    lib().oracle_caslr_eff_driver(_i(verbose), _i(n), _i(n2), _i(n_targ), _i(n_max), _i(max_iter), _d(tol), _i(max_dav),
                                  _cb("oracle_csr_apbmul"), _cb("oracle_csr_ambmul"), _cb("oracle_csr_spdmul"),
                                  _cb("oracle_csr_smdmul"), _cb("oracle_lrprec"), _p(eig), _p(evec), C.byref(ok))
    out = _collect(n_max)
    out.update(eig=eig, ok=bool(ok.value), hist_eig=out["eig"])
    return out


def ortho_cd(u):
    assert u.flags.f_contiguous
    n, m = u.shape
    g = C.c_double(0)
    ok = C.c_int32(0)
    lib().oracle_ortho_cd(_i(n), _i(m), _p(u), C.byref(g), C.byref(ok))
    return g.value, bool(ok.value)


def ortho_vs_x(x, u):
    assert x.flags.f_contiguous and u.flags.f_contiguous
    n, m = x.shape
    k = u.shape[1]
    lib().oracle_ortho_vs_x(_i(n), _i(m), _i(k), _p(x), _p(u))


def b_ortho(u, bu):
    assert u.flags.f_contiguous and bu.flags.f_contiguous
    n, m = u.shape
    lib().oracle_b_ortho(_i(n), _i(m), _p(u), _p(bu))


def b_ortho_vs_x(x, bx, u):
    assert x.flags.f_contiguous and bx.flags.f_contiguous and u.flags.f_contiguous
    n, m = x.shape
    lib().oracle_b_ortho_vs_x(_i(n), _i(m), _i(u.shape[1]), _p(x), _p(bx), _p(u))


def csr_bvec(x):
    n, m = x.shape
    bx = np.zeros_like(x, order="F")
    lib().oracle_csr_bvec(_i(n), _i(m), _p(x), _p(bx))
    return bx


def ortho(u):
    n, m = u.shape
    lib().oracle_ortho(_i(n), _i(m), _p(u))


def get_coeffs(a_red, len_u, n_max, n_act):
    len_a = a_red.shape[0]
    u_x = np.zeros((len_u, n_max), order="F")
    u_p = np.zeros((len_u, n_act), order="F")
    lib().oracle_get_coeffs(_i(len_a), _i(len_u), _i(n_max), _i(n_act), _p(a_red), _p(u_x), _p(u_p))
    return u_x, u_p


def dsyev(a, upper=False):
    a = np.asfortranarray(a.copy())
    n = a.shape[0]
    w = np.zeros(n)
    info = C.c_int32(0)
    lib().oracle_dsyev(_i(n), _p(a), _i(n), _p(w), C.byref(info), _i(1 if upper else 0))
    assert info.value == 0
    return w, a


def csr_matvec(x):
    n, m = x.shape
    ax = np.zeros_like(x, order="F")
    lib().oracle_csr_matvec(_i(n), _i(m), _p(x), _p(ax))
    return ax


def diag_precnd(x, fac):
    n, m = x.shape
    px = np.zeros_like(x, order="F")
    lib().oracle_diag_precnd(_i(n), _i(m), _d(fac), _p(x), _p(px))
    return px


def set_accurate_eig(on: bool):
    """DIAGNOSTIC: reduced eigenproblems through dpotrf + dgesvj (high relative accuracy) instead of
    dsyev; never the default (see diaglib_oracle.cpp, reduced_eig)"""
    lib().oracle_set_accurate_eig(1 if on else 0)


def set_threads(n):
    lib().oracle_set_threads(int(n))


def get_threads():
    return int(lib().oracle_get_threads())


def blas_config():
    return lib().oracle_blas_config().decode()
