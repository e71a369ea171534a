"""Sequence-level reproducibility: ortho_vs_x (Gram, Cholesky, triangular multiply, projection back to back, as in
a LOBPCG iteration) repeated on the same device-resident input; every output compared with the first one.
usage: python tools/seq_repro.py [log2n=21] [reps=25]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
n = 1 << log2n
D.init(0)
lib = D.lib()
i32 = lambda v_: C.byref(C.c_int32(int(v_)))  # noqa: E731
v0 = K.DeviceArray((n, 111))
v = K.DeviceArray((n, 111))
lib.diaglib_b200_k_fill_uniform(v0.ptr, n, 111, n, 1)
# x = the first 74 columns, orthonormalised once; u = the 37 columns behind them, made nearly dependent on x so that
# several sweeps and passes are needed (as for the preconditioned residuals of a converging solve)
g, ok = C.c_double(0), C.c_int32(0)
lib.diaglib_b200_ortho_cd(i32(n), i32(74), C.c_void_p(v0.ptr), C.byref(g), C.byref(ok))
mix = K.DeviceArray.from_numpy(np.asfortranarray(np.random.default_rng(2).standard_normal((74, 37))))
eps_c = K.DeviceArray.from_numpy(np.asfortranarray(1e-6 * np.eye(37)))
tmp = K.DeviceArray((n, 37))
lib.diaglib_b200_k_block_mul(n, v0.col_ptr(74), n, 37, eps_c.ptr, 37, 37, 1.0, 0.0, tmp.ptr, n)          # 1e-6 u
lib.diaglib_b200_k_block_mul(n, v0.ptr, n, 74, mix.ptr, 74, 37, 1.0, 1.0, tmp.ptr, n)                     # + x mix
lib.diaglib_b200_d2d(v0.col_ptr(74), tmp.ptr, 8 * n * 37)
lib.diaglib_b200_sync()
ref, bad = None, 0
for r in range(reps):
    lib.diaglib_b200_d2d(v.ptr, v0.ptr, 8 * n * 111)
    lib.diaglib_b200_ortho_vs_x(i32(n), i32(74), i32(37), C.c_void_p(v.ptr), C.c_void_p(v.col_ptr(74)), None, None)
    lib.diaglib_b200_sync()
    assert lib.diaglib_b200_last_status() == 0
    got = v.numpy()[:, 74:]
    st = D.last_stats()
    if ref is None:
        ref = got.copy()
        print("passes", st["ortho_cd_passes"], "sweeps", st["ortho_vs_x_sweeps"], flush=True)
    elif not np.array_equal(got, ref):
        bad += 1
        d = np.abs(got - ref)
        idx = np.argwhere(d > 0)
        print(f"rep {r}: {len(idx)} elements differ, rows {idx[:, 0].min()}..{idx[:, 0].max()}, cols {sorted(set(idx[:, 1].tolist()))[:8]}, "
              f"max |diff| {d.max():.3e}; passes {st['ortho_cd_passes']} sweeps {st['ortho_vs_x_sweeps']}", flush=True)
print(f"ortho_vs_x 74 + 37 columns, n = 2^{log2n}: {bad} of {reps - 1} repetitions differ")
