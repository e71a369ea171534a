"""Kernel-level ctypes wrappers (include/diaglib_b200_kernels.h) used by tests/ and bench.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .api import DiaglibError, _check, _ptr, init, lib


class DeviceArray:
    """A column-major float64 block in HBM owned through the library's own allocator."""

    def __init__(self, shape, ld: int | None = None):
        init()
        self.shape = tuple(shape)
        rows = self.shape[0]
        cols = self.shape[1] if len(self.shape) > 1 else 1
        self.ld = rows if ld is None else ld
        self.nbytes = 8 * max(1, self.ld * cols)
        self.ptr = lib().diaglib_b200_malloc(self.nbytes)
        if not self.ptr:
            raise DiaglibError(f"device allocation of {self.nbytes} bytes failed")

    @classmethod
    def from_numpy(cls, a: np.ndarray):
        a = np.asfortranarray(a, dtype=np.float64)
        d = cls(a.shape if a.ndim == 2 else (a.shape[0], 1))
        _check(lib().diaglib_b200_h2d(d.ptr, _ptr(a), a.nbytes), "h2d")
        return d

    def numpy(self) -> np.ndarray:
        rows = self.shape[0]
        cols = self.shape[1] if len(self.shape) > 1 else 1
        out = np.empty((self.ld, cols), dtype=np.float64, order="F")
        _check(lib().diaglib_b200_d2h(_ptr(out), self.ptr, out.nbytes), "d2h")
        return out[:rows, :]

    def data_ptr(self) -> int:
        return self.ptr

    def col_ptr(self, j: int) -> int:
        return self.ptr + 8 * self.ld * j

    def free(self):
        if self.ptr:
            lib().diaglib_b200_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def gram(a: DeviceArray, p: int, b: DeviceArray, q: int, n: int | None = None, sym_lower: bool = False,
         a_off: int = 0, b_off: int = 0) -> np.ndarray:
    """C = A[:, a_off:a_off+p]^T B[:, b_off:b_off+q] (dgemm 't','n'; diaglib.f90:403,3256,3543)."""
    n = a.shape[0] if n is None else n
    c = DeviceArray((p, q))
    _check(lib().diaglib_b200_k_gram(n, a.col_ptr(a_off), a.ld, p, b.col_ptr(b_off), b.ld, q, c.ptr, p,
                                     1 if sym_lower else 0), "k_gram")
    out = c.numpy()
    c.free()
    return out


def block_mul(v: DeviceArray, p: int, cmat: np.ndarray, y: DeviceArray, alpha=1.0, beta=0.0, n: int | None = None,
              v_off: int = 0, y_off: int = 0) -> None:
    """Y[:, y_off:y_off+q] = alpha V[:, v_off:v_off+p] C + beta Y (dgemm 'n','n'; diaglib.f90:420,495,3544)."""
    n = v.shape[0] if n is None else n
    cmat = np.asfortranarray(cmat, dtype=np.float64)
    q = cmat.shape[1]
    cd = DeviceArray.from_numpy(cmat)
    _check(lib().diaglib_b200_k_block_mul(n, v.col_ptr(v_off), v.ld, p, cd.ptr, cd.ld, q, alpha, beta,
                                          y.col_ptr(y_off), y.ld), "k_block_mul")
    lib().diaglib_b200_sync()
    cd.free()


def block_mul_gram(v: DeviceArray, p: int, cmat: np.ndarray, y: DeviceArray, alpha=1.0, beta=0.0, upper_tri=False):
    """Y = alpha V[:, :p] C + beta Y and G = Y^T Y in one kernel.  Returns G."""
    n = v.shape[0]
    cmat = np.asfortranarray(cmat, dtype=np.float64)
    q = cmat.shape[1]
    cd = DeviceArray.from_numpy(cmat)
    gd = DeviceArray((q, q))
    _check(lib().diaglib_b200_k_block_mul_gram(n, v.ptr, v.ld, p, cd.ptr, cd.ld, q, alpha, beta, y.ptr, y.ld,
                                               1 if upper_tri else 0, gd.ptr, q), "k_block_mul_gram")
    lib().diaglib_b200_sync()
    out = gd.numpy()
    cd.free()
    gd.free()
    return out


def residual(ax: DeviceArray, x: DeviceArray, theta, active, r: DeviceArray):
    n, m = ax.shape
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    active = np.ascontiguousarray(active, dtype=np.int32)
    norms = np.zeros(2 * m)
    _check(lib().diaglib_b200_k_residual(n, m, ax.ptr, ax.ld, x.ptr, x.ld, _ptr(theta), _ptr(active), r.ptr, r.ld,
                                         _ptr(norms)), "k_residual")
    return norms[:m], norms[m:]


def sym_eig(a: np.ndarray, upper: bool = False):
    """dsyev('v',uplo) replacement.  Returns (w, z, sweeps)."""
    init()
    a = np.asfortranarray(a.copy(), dtype=np.float64)
    k = a.shape[0]
    w = np.zeros(k)
    sweeps = lib().diaglib_b200_k_sym_eig(k, _ptr(a), k, 1 if upper else 0, _ptr(w))
    if sweeps < 0:
        raise DiaglibError(f"sym_eig did not converge ({sweeps})")
    sym_eig.last_path = 1 if sweeps >= 1000 else 2   # 1: one-sided Jacobi on the Cholesky factor, 2: two-sided
    return w, a, sweeps % 1000


def set_eig_mode(mode: int, block: int = 0) -> int:
    """0: one-sided Jacobi on the Cholesky factor, two-sided fallback (default); 1: two-sided only"""
    init()
    return int(lib().diaglib_b200_k_set_eig_mode(int(mode), int(block)))


def set_reference_restart(on: bool) -> bool:
    """gen_david_driver: reproduce the reference's `bspace = zero` at a restart (diaglib.f90:2200) literally"""
    init()
    return bool(lib().diaglib_b200_k_set_reference_restart(1 if on else 0))


def set_spec_ortho(on: bool) -> bool:
    """speculative (device-decided) ortho_cd / ortho_vs_x chains on/off; returns the previous setting"""
    init()
    return bool(lib().diaglib_b200_k_set_spec_ortho(1 if on else 0))


def sym_eig_time_ms(a: np.ndarray, upper: bool = False, reps: int = 10) -> float:
    init()
    a = np.asfortranarray(a, dtype=np.float64)
    k = a.shape[0]
    return float(lib().diaglib_b200_k_sym_eig_time_ms(k, _ptr(a), k, 1 if upper else 0, reps))


def chol_inv(metric: np.ndarray):
    """One factor+invert step of ortho_cd.  Returns (T, dict)."""
    init()
    metric = np.asfortranarray(metric, dtype=np.float64)
    m = metric.shape[0]
    t = np.zeros((m, m), order="F")
    out = np.zeros(5)
    hard = lib().diaglib_b200_k_chol_inv(m, _ptr(metric), _ptr(t), _ptr(out))
    return t, dict(l_norm=out[0], linv_norm=out[1], shift=out[2], info_first=int(out[3]), n_shifts=int(out[4]),
                   hard_fail=int(hard))


def get_coeffs(a_red: np.ndarray, len_u: int, n_max: int, n_act: int):
    init()
    a_red = np.asfortranarray(a_red, dtype=np.float64)
    len_a = a_red.shape[0]
    u_p = np.zeros((len_u, n_act), order="F")
    out = np.zeros(4, np.int32)
    lib().diaglib_b200_k_get_coeffs(len_a, len_u, n_max, n_act, _ptr(a_red), _ptr(u_p), _ptr(out))
    return u_p, dict(sweeps=int(out[0]), cd_passes=int(out[1]), fail=int(out[2]), qr=int(out[3]))


def timer_start():
    lib().diaglib_b200_timer_start()


def timer_stop_ms() -> float:
    return float(lib().diaglib_b200_timer_stop_ms())
