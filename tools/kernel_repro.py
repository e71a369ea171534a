"""Kernel-level reproducibility: each dense kernel of the ortho / Ritz steps is run `reps` times on the same
input and every output is compared with the first one bit for bit.
usage: python tools/kernel_repro.py [log2n=20] [reps=40]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = 1 << log2n
D.init(0)
lib = D.lib()
v0 = K.DeviceArray((n, 111))      # master copy
v = K.DeviceArray((n, 111))       # working copy
y = K.DeviceArray((n, 37))
lib.diaglib_b200_k_fill_uniform(v0.ptr, n, 111, n, 1)
rng = np.random.default_rng(5)
tm = K.DeviceArray.from_numpy(np.asfortranarray(np.triu(rng.standard_normal((37, 37)) * 0.05 + np.eye(37))))
cm = K.DeviceArray.from_numpy(np.asfortranarray(rng.standard_normal((111, 37))))
xu = K.DeviceArray.from_numpy(np.asfortranarray(rng.standard_normal((74, 37)) * 1e-3))


def restore():
    lib.diaglib_b200_d2d(v.ptr, v0.ptr, 8 * n * 111)


def run(name, fn, out_ptr_cols):
    arr, c0, nc = out_ptr_cols
    ref, bad, worst = None, 0, 0.0
    for r in range(reps):
        restore()
        fn()
        lib.diaglib_b200_sync()
        got = arr.numpy()[:, c0:c0 + nc]
        if ref is None:
            ref = got.copy()
        elif not np.array_equal(got, ref):
            bad += 1
            d = np.abs(got - ref)
            worst = max(worst, float(d.max()))
            if bad == 1:
                idx = np.argwhere(d > 0)
                print(f"   first mismatch: {len(idx)} elements, rows {idx[:, 0].min()}..{idx[:, 0].max()}, cols {sorted(set(idx[:, 1].tolist()))[:12]}, "
                      f"max |diff| {d.max():.3e}, ref there {ref[tuple(idx[0])]:.6e} got {got[tuple(idx[0])]:.6e}")
    print(f"{name}: {bad} of {reps - 1} repetitions differ (max |diff| {worst:.3e})", flush=True)


run("trmm 37 in place", lambda: lib.diaglib_b200_k_trmm(n, v.ptr, n, 37, tm.ptr), (v, 0, 37))
run("trmm 37 out of place", lambda: lib.diaglib_b200_k_trmm_oop(n, v.ptr, n, 37, tm.ptr, y.ptr, n), (y, 0, 37))
run("projection [x u][-xu; I] in place", lambda: lib.diaglib_b200_k_project_out(n, 74, 37, v.ptr, n, xu.ptr, v.col_ptr(74), n), (v, 74, 37))
run("block_mul 111 -> 37", lambda: lib.diaglib_b200_k_block_mul(n, v.ptr, n, 111, cm.ptr, 111, 37, 1.0, 0.0, y.ptr, n), (y, 0, 37))
