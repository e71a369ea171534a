"""Diagnostic: kernel-by-kernel comparison against numpy / the oracle at a large n."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K, problems as P
from oracle import oracle as O

n = 1 << 20
D.init(0)
O.set_threads(16)
csr = P.toy_sparse(n)
O.set_csr(*csr)
D.set_csr(*csr)
m = 13
x = P.guess(n, m)
ref = O.csr_matvec(x)
dx, dax = K.DeviceArray.from_numpy(x), K.DeviceArray((n, m))
i32 = lambda v: C.byref(C.c_int32(v))
D.lib().diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(dx.ptr), C.c_void_p(dax.ptr))
D.lib().diaglib_b200_sync()
ax = dax.numpy()
print("spmm bit-exact:", np.array_equal(ax, ref), np.abs(ax - ref).max())
for (p, q, sym, same) in [(13, 13, True, True), (13, 13, True, False), (26, 26, True, False), (26, 13, False, False), (39, 39, True, False)]:
    a = np.asfortranarray(np.random.default_rng(p + q).standard_normal((n, p)))
    b = a if same else np.asfortranarray(np.random.default_rng(99).standard_normal((n, q)) + (a[:, :q] if sym else 0) * 1000)
    if sym and not same:
        d = np.linspace(1.0, 2.0, n)
        b = np.asfortranarray(a * d[:, None])
    da = K.DeviceArray.from_numpy(a)
    db = da if same else K.DeviceArray.from_numpy(b)
    c = K.gram(da, p, db, q, sym_lower=sym)
    r = a.T @ b
    print("gram", p, q, "sym" if sym else "full", "same" if same else "", "relerr", np.abs(c - r).max() / np.abs(r).max())
for (p, q) in [(13, 13), (26, 13), (39, 13)]:
    v = np.asfortranarray(np.random.default_rng(p).standard_normal((n, p)))
    cm = np.asfortranarray(np.random.default_rng(q).standard_normal((p, q)))
    dv, dy = K.DeviceArray.from_numpy(v), K.DeviceArray((n, q))
    K.block_mul(dv, p, cm, dy)
    r = v @ cm
    print("block_mul", p, q, "relerr", np.abs(dy.numpy() - r).max() / np.abs(r).max())
u = P.guess(n, m)
uo = u.copy(order="F")
go, _ = O.ortho_cd(uo)
g, ok = D.ortho_cd(n, m, u)
print("ortho_cd growth", g, go, "maxdiff", np.abs(u - uo).max(), "orth", np.abs(u.T @ u - np.eye(m)).max())
w = np.asfortranarray(np.random.default_rng(5).standard_normal((n, m)) * 1e-3 + 0.5 * u)
wo = w.copy(order="F")
O.ortho_vs_x(uo, wo)
D.ortho_vs_x(n, m, m, u, w)
print("ortho_vs_x maxdiff", np.abs(w - wo).max(), "xu", np.abs(u.T @ w).max(), "orth", np.abs(w.T @ w - np.eye(m)).max())
for mi in (1, 2):
    ev_o = P.guess(n, m)
    ev_g = ev_o.copy(order="F")
    ro = O.lobpcg(ev_o, 8, mi, 1e-8)
    eig = np.zeros(m)
    D.lobpcg_driver(False, False, n, 8, m, mi, 1e-8, 0.0, None, None, None, eig, ev_g)
    h = D.last_history(m)
    print("lobpcg max_iter", mi, "oracle eig", ro["hist_eig"][-1][:4], "gpu", h["eig"][-1][:4], "rms o", ro["rms"][-1][:3], "g", h["rms"][-1][:3])
