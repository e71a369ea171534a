"""Host-side mirror of the reference interface (diaglib.f90:171-172, 1483-1484, 3052, 3185,
3481) over the C-ABI of libdiaglib_b200.so.  Routine names and argument order are the
reference's; arrays are numpy (host, Fortran order) or anything with ``data_ptr()`` /
an integer address (device memory, e.g. a torch CUDA tensor)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MATVEC_T = C.CFUNCTYPE(None, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p)
PRECND_T = C.CFUNCTYPE(None, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p, C.c_void_p)
LRPREC_T = C.CFUNCTYPE(None, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                       C.c_void_p, C.c_void_p)

STATUS = {0: "ok", 1: "reduced eigensolver failed", 2: "device allocation failed", 3: "Cholesky shift loop exhausted",
          4: "ortho_vs_x failed", 5: "no CUDA device", 6: "bad argument", 7: "NCCL failure"}


class DiaglibError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(_HERE, "libdiaglib_b200.so")


def build(verbose: bool = False) -> None:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built: there is no
    Python/CPU fallback for any compute entry point."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise DiaglibError(f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(diaglib_b200 has no CPU fallback)")
        L = C.CDLL(p)
        L.diaglib_b200_last_message.restype = C.c_char_p
        L.diaglib_b200_stream.restype = C.c_void_p
        L.diaglib_b200_malloc.restype = C.c_void_p
        L.diaglib_b200_malloc.argtypes = [C.c_int64]
        L.diaglib_b200_free.argtypes = [C.c_void_p]
        L.diaglib_b200_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.diaglib_b200_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.diaglib_b200_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.diaglib_b200_k_fill_uniform.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64]
        L.diaglib_b200_timer_stop_ms.restype = C.c_double
        L.diaglib_b200_set_csr.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.diaglib_b200_set_csr_b.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.diaglib_b200_set_csr_device.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.diaglib_b200_set_csr_row_order.argtypes = [C.c_void_p]
        L.diaglib_b200_k_gen_fci.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_double, C.c_int64,
                                             C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.diaglib_b200_k_true_residual.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.diaglib_b200_set_csr_lr.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.diaglib_b200_set_lr_diag.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
        L.diaglib_b200_set_halo.argtypes = [C.c_int32] + [C.c_void_p] * 5
        L.diaglib_b200_k_gram.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                          C.c_void_p, C.c_int32, C.c_int32]
        L.diaglib_b200_k_block_mul.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                               C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_int64]
        L.diaglib_b200_k_block_mul_gram.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                                    C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_int64, C.c_int32,
                                                    C.c_void_p, C.c_int32]
        L.diaglib_b200_k_residual.argtypes = [C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.diaglib_b200_k_sym_eig.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.diaglib_b200_k_set_eig_mode.argtypes = [C.c_int32, C.c_int32]
        L.diaglib_b200_k_sym_eig_time_ms.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.diaglib_b200_k_sym_eig_time_ms.restype = C.c_double
        L.diaglib_b200_k_set_tuning.argtypes = [C.c_char_p, C.c_int32]
        L.diaglib_b200_k_trmm.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.diaglib_b200_k_project_out.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        L.diaglib_b200_k_trmm_oop.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
        L.diaglib_b200_k_time_small.argtypes = [C.c_int32] * 5
        L.diaglib_b200_k_time_small.restype = C.c_double
        L.diaglib_b200_k_chol_inv.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.diaglib_b200_k_get_coeffs.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 3
        L.diaglib_b200_comm_init.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        L.diaglib_b200_comm_unique_id.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().diaglib_b200_last_message()
        raise DiaglibError(f"{what}: status {code} ({STATUS.get(code, '?')}) {msg.decode() if msg else ''}")


def init(device: int | None = None) -> None:
    _check(lib().diaglib_b200_init(C.c_int32(-1 if device is None else device)), "diaglib_b200_init")


def _ptr(a):
    """address of a numpy array / torch tensor / raw integer address"""
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    if a is None:
        return C.c_void_p(0)
    return C.c_void_p(int(a))


def _i(v):
    return C.byref(C.c_int32(int(v)))


def _d(v):
    return C.byref(C.c_double(float(v)))


_keep: list = []


def _callback(cb, kind):
    """matvec / precnd argument: None or 'csr'/'diag' selects the built-in device callback
    (diaglib_b200_csr_matvec / diaglib_b200_diag_precnd); a ctypes function pointer is passed
    through; a Python callable f(n, m, [shift,] x_ptr, ax_ptr) receiving device addresses is
    wrapped."""
    L = lib()
    if cb is None or isinstance(cb, str):
        builtin = dict(matvec=L.diaglib_b200_csr_matvec, precnd=L.diaglib_b200_diag_precnd, bvec=L.diaglib_b200_csr_bvec,
                       apbmul=L.diaglib_b200_csr_apbmul, ambmul=L.diaglib_b200_csr_ambmul,
                       spdmul=L.diaglib_b200_csr_spdmul, smdmul=L.diaglib_b200_csr_smdmul, lrprec=L.diaglib_b200_lrprec)
        return C.cast(builtin[kind], C.c_void_p)
    if isinstance(cb, C._CFuncPtr):
        return C.cast(cb, C.c_void_p)
    if kind in ("matvec", "bvec", "apbmul", "ambmul", "spdmul", "smdmul"):
        f = MATVEC_T(lambda n, m, x, ax: cb(n[0], m[0], x, ax))
    elif kind == "lrprec":
        f = LRPREC_T(lambda n, m, fac, xp, xm, yp, ym: cb(n[0], m[0], fac[0], xp, xm, yp, ym))
    else:
        f = PRECND_T(lambda n, m, s, x, px: cb(n[0], m[0], s[0], x, px))
    _keep.append(f)
    return C.cast(f, C.c_void_p)


def set_halo(halo_plan) -> None:
    """halo_plan = (peer, send_row0, send_cnt, recv_off, recv_cnt), see diaglib_b200/partition.py"""
    peer, s0, sc, ro, rc = halo_plan
    peer = np.ascontiguousarray(peer, dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in (s0, sc, ro, rc)]
    _check(lib().diaglib_b200_set_halo(len(peer), _ptr(peer), *[_ptr(a) for a in arrs]), "set_halo")


def set_csr_row_order(order) -> None:
    """Processing order of the local rows in the built-in matvec (a permutation; None = natural).
    Results do not depend on it; a locality-preserving order (problems.tile_order_3d for grid
    stencils) raises the cache hit rate of the gathers."""
    if order is None:
        _check(lib().diaglib_b200_set_csr_row_order(None), "set_csr_row_order")
        return
    order = np.ascontiguousarray(order, dtype=np.int32)
    _check(lib().diaglib_b200_set_csr_row_order(_ptr(order)), "set_csr_row_order")


def set_csr(rowptr, col, val, diag, n_halo: int = 0, halo_plan=None, row_order=None) -> None:
    """Install the local rows of the matrix used by the built-in callbacks (the reference keeps
    its matrix in a module global too: utils.f90:4, main.f90:73).  Column indices are local
    (see include/diaglib_b200.h); halo_plan = (peer, send_row0, send_cnt, recv_off, recv_cnt)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    diag = np.ascontiguousarray(diag, dtype=np.float64)
    n_loc = len(rowptr) - 1
    _check(lib().diaglib_b200_set_csr(n_loc, int(n_halo), _ptr(rowptr), _ptr(col), _ptr(val), _ptr(diag)), "set_csr")
    if halo_plan is not None:
        set_halo(halo_plan)
    if row_order is not None:
        set_csr_row_order(row_order)


def set_csr_device(n_loc: int, n_halo: int, nnz: int, rowptr_dev: int, col_dev: int, val_dev: int, diag_dev: int,
                   halo_plan=None) -> None:
    """set_csr for arrays that already live in HBM (raw device addresses; the caller keeps them alive)."""
    _check(lib().diaglib_b200_set_csr_device(int(n_loc), int(n_halo), int(nnz), C.c_void_p(rowptr_dev), C.c_void_p(col_dev),
                                             C.c_void_p(val_dev), C.c_void_p(diag_dev)), "set_csr_device")
    if halo_plan is not None:
        set_halo(halo_plan)


def set_csr_b(rowptr, col, val, n_halo: int = 0) -> None:
    """Install the local rows of the metric B used by the built-in bvec (generalized problem)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    _check(lib().diaglib_b200_set_csr_b(len(rowptr) - 1, int(n_halo), _ptr(rowptr), _ptr(col), _ptr(val)), "set_csr_b")


def lobpcg_driver(verbose, gen_eig, n, n_targ, n_max, max_iter, tol, shift, matvec, precnd, bvec, eig, evec) -> bool:
    """diaglib.f90:171-172.  eig (n_max) and evec (n, n_max; guess in, vectors out) are
    overwritten.  Returns `ok`.  Raises DiaglibError where the reference would `stop`.
    gen_eig=True solves A x = lambda B x; bvec=None selects the built-in product with the metric
    installed by set_csr_b."""
    ok = C.c_int32(0)
    lib().diaglib_b200_lobpcg_driver(_i(verbose), _i(gen_eig), _i(n), _i(n_targ), _i(n_max), _i(max_iter), _d(tol),
                                     _d(shift), _callback(matvec, "matvec"), _callback(precnd, "precnd"),
                                     _callback(bvec, "bvec") if gen_eig else None, _ptr(eig), _ptr(evec), C.byref(ok))
    _check(lib().diaglib_b200_last_status(), "lobpcg_driver")
    return bool(ok.value)


def davidson_driver(verbose, n, n_targ, n_max, max_iter, tol, max_dav, shift, matvec, precnd, eig, evec) -> bool:
    """diaglib.f90:1483-1484."""
    ok = C.c_int32(0)
    lib().diaglib_b200_davidson_driver(_i(verbose), _i(n), _i(n_targ), _i(n_max), _i(max_iter), _d(tol), _i(max_dav),
                                       _d(shift), _callback(matvec, "matvec"), _callback(precnd, "precnd"), _ptr(eig),
                                       _ptr(evec), C.byref(ok))
    _check(lib().diaglib_b200_last_status(), "davidson_driver")
    return bool(ok.value)


def gen_david_driver(verbose, n, n_targ, n_max, max_iter, tol, max_dav, shift, matvec, precnd, bvec, eig, evec) -> bool:
    """diaglib.f90:1855-1856.  bvec=None selects the built-in product with the metric installed
    by set_csr_b."""
    ok = C.c_int32(0)
    lib().diaglib_b200_gen_david_driver(_i(verbose), _i(n), _i(n_targ), _i(n_max), _i(max_iter), _d(tol), _i(max_dav),
                                        _d(shift), _callback(matvec, "matvec"), _callback(precnd, "precnd"),
                                        _callback(bvec, "bvec"), _ptr(eig), _ptr(evec), C.byref(ok))
    _check(lib().diaglib_b200_last_status(), "gen_david_driver")
    return bool(ok.value)


def set_lr(apb, amb, spd, smd, aa_diag, sigma_diag, n_halo: int = 0) -> None:
    """Install the four CSR matrices (rowptr, col, val) A+B, A-B, S+D, S-D and the diagonals of A
    and S used by the built-in linear-response callbacks."""
    for which, (rowptr, col, val) in enumerate((apb, amb, spd, smd)):
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        _check(lib().diaglib_b200_set_csr_lr(which, len(rowptr) - 1, int(n_halo), _ptr(rowptr), _ptr(col), _ptr(val)),
               "set_csr_lr")
    aa = np.ascontiguousarray(aa_diag, dtype=np.float64)
    sg = np.ascontiguousarray(sigma_diag, dtype=np.float64)
    _check(lib().diaglib_b200_set_lr_diag(len(aa), _ptr(aa), _ptr(sg)), "set_lr_diag")


def caslr_eff_driver(verbose, n, n2, n_targ, n_max, max_iter, tol, max_dav, apbmul, ambmul, spdmul, smdmul, lrprec,
                     eig, evec) -> bool:
    """diaglib.f90:1024-1025.  evec is (2n, n_max): rows [0,n) = Y, [n,2n) = Z.  None selects the
    built-in callbacks on the matrices installed by set_lr."""
    ok = C.c_int32(0)
    lib().diaglib_b200_caslr_eff_driver(_i(verbose), _i(n), _i(n2), _i(n_targ), _i(n_max), _i(max_iter), _d(tol),
                                        _i(max_dav), _callback(apbmul, "apbmul"), _callback(ambmul, "ambmul"),
                                        _callback(spdmul, "spdmul"), _callback(smdmul, "smdmul"),
                                        _callback(lrprec, "lrprec"), _ptr(eig), _ptr(evec), C.byref(ok))
    _check(lib().diaglib_b200_last_status(), "caslr_eff_driver")
    return bool(ok.value)


def ortho_cd(n, m, u):
    """diaglib.f90:3185.  Returns (growth, ok)."""
    g = C.c_double(0)
    ok = C.c_int32(0)
    lib().diaglib_b200_ortho_cd(_i(n), _i(m), _ptr(u), C.byref(g), C.byref(ok))
    _check(lib().diaglib_b200_last_status(), "ortho_cd")
    return g.value, bool(ok.value)


def ortho_vs_x(n, m, k, x, u, ax=None, au=None) -> None:
    """diaglib.f90:3481."""
    lib().diaglib_b200_ortho_vs_x(_i(n), _i(m), _i(k), _ptr(x), _ptr(u), _ptr(ax), _ptr(au))
    _check(lib().diaglib_b200_last_status(), "ortho_vs_x")


def b_ortho(n, m, u, bu) -> None:
    """diaglib.f90:3094."""
    lib().diaglib_b200_b_ortho(_i(n), _i(m), _ptr(u), _ptr(bu))
    _check(lib().diaglib_b200_last_status(), "b_ortho")


def b_ortho_vs_x(n, m, k, x, bx, u) -> None:
    """diaglib.f90:3576."""
    lib().diaglib_b200_b_ortho_vs_x(_i(n), _i(m), _i(k), _ptr(x), _ptr(bx), _ptr(u))
    _check(lib().diaglib_b200_last_status(), "b_ortho_vs_x")


def ortho(n, m, u, w=None) -> None:
    """diaglib.f90:3052 (QR fallback)."""
    lib().diaglib_b200_ortho(_i(n), _i(m), _ptr(u), _ptr(w))
    _check(lib().diaglib_b200_last_status(), "ortho")


def last_history(n_max: int):
    L = lib().diaglib_b200_history_len()
    it = np.zeros(L, np.int32)
    n_act = np.zeros(L, np.int32)
    eig = np.zeros((L, n_max))
    rms = np.zeros((L, n_max))
    mx = np.zeros((L, n_max))
    done = np.zeros((L, n_max), np.int32)
    if L:
        lib().diaglib_b200_history_get(_ptr(it), _ptr(n_act), _ptr(eig), _ptr(rms), _ptr(mx), _ptr(done))
    return dict(it=it, n_act=n_act, eig=eig, rms=rms, max=mx, done=done)


def last_timers():
    t = np.zeros(12)
    lib().diaglib_b200_timers(_ptr(t))
    return dict(mv=t[0], diag=t[1], ortho=t[2], total=t[3], gram=t[4], ritz=t[5], resid=t[6], stage=t[7],
                k_gram=t[8], k_block_mul=t[9], k_copy=t[11])


def set_profile(on: bool) -> None:
    """per-kernel-family device timing (two CUDA events per launch); off by default"""
    lib().diaglib_b200_set_profile(C.c_int32(1 if on else 0))


def peer_info():
    """How the k x k all-reduces travel: ranks sharing the peer window (0 = NCCL), window calls of the last
    driver call, time-out flag, completed window calls."""
    s = np.zeros(4, np.int64)
    lib().diaglib_b200_peer_info(_ptr(s))
    return dict(window_ranks=int(s[0]), calls=int(s[1]), error=int(s[2]), epoch=int(s[3]))


def last_stats():
    s = np.zeros(8, np.int64)
    lib().diaglib_b200_stats(_ptr(s))
    return dict(ortho_cd_passes=int(s[0]), ortho_vs_x_sweeps=int(s[1]), qr_fallbacks=int(s[2]), chol_shifts=int(s[3]),
                launches=int(s[4]), launches_total=int(s[5]), host_syncs=int(s[6]),
                eig_calls=int(s[7] & 0xffff), eig_sweeps=int((s[7] >> 16) & 0xffffff), eig_two_sided_fallbacks=int(s[7] >> 40))
