"""Times (and, under ncu, exposes) the two dominant kernel families at the C3 shapes.
usage: python tools/kernel_bench.py [log2n=22] [reps=3]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diaglib_b200 as D  # noqa: E402
from diaglib_b200 import kernels as K  # noqa: E402


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n = 1 << log2n
    D.init(0)
    lib = D.lib()
    v = K.DeviceArray((n, 111))
    w = K.DeviceArray((n, 111))
    lib.diaglib_b200_k_fill_uniform(v.ptr, n, 111, n, 1)
    lib.diaglib_b200_k_fill_uniform(w.ptr, n, 111, n, 777)
    out = {}

    def t_gram(p, q, sym, a, b, tag):
        c = K.DeviceArray((p, q))
        lib.diaglib_b200_k_gram(n, a.ptr, n, p, b.ptr, n, q, c.ptr, p, sym)
        lib.diaglib_b200_sync()
        K.timer_start()
        for _ in range(reps):
            lib.diaglib_b200_k_gram(n, a.ptr, n, p, b.ptr, n, q, c.ptr, p, sym)
        ms = K.timer_stop_ms() / reps
        fl = n * p * (p + 1) if sym else 2.0 * n * p * q
        by = 8.0 * n * ((p if a is b else p + q))
        out[tag] = dict(ms=round(ms, 4), tflops=round(fl / ms / 1e9, 2), gbs=round(by / ms / 1e6, 1))
        c.free()

    def t_bmul(p, q, tag, inplace=False, tri=False):
        cm = np.asfortranarray(np.triu(np.random.default_rng(0).standard_normal((p, q))) if tri
                               else np.random.default_rng(0).standard_normal((p, q)))
        cd = K.DeviceArray.from_numpy(cm)
        y = v if inplace else K.DeviceArray((n, q))
        lib.diaglib_b200_k_block_mul(n, v.ptr, n, p, cd.ptr, p, q, 1.0, 0.0, y.ptr, n)
        lib.diaglib_b200_sync()
        K.timer_start()
        for _ in range(reps):
            lib.diaglib_b200_k_block_mul(n, v.ptr, n, p, cd.ptr, p, q, 1.0, 0.0, y.ptr, n)
        ms = K.timer_stop_ms() / reps
        fl = 2.0 * n * p * q
        by = 8.0 * n * (p + q)
        out[tag] = dict(ms=round(ms, 4), tflops=round(fl / ms / 1e9, 2), gbs=round(by / ms / 1e6, 1))
        cd.free()
        if not inplace:
            y.free()

    only = sys.argv[3] if len(sys.argv) > 3 else ""
    if only == "bmul":
        t_bmul(111, 37, "bmul_111x37")
        t_bmul(37, 37, "bmul_37x37")
        print(json.dumps(dict(n=n, **out)))
        return
    if only == "spmm":
        import ctypes as C
        from diaglib_b200 import problems as P
        nx = round(n ** (1 / 3))
        csr = P.lap3d(nx, nx, nx, delta=1.0)
        D.set_csr(*csr)
        nnz = len(csr[1])
        for m in (40, 37, 32, 24, 16, 8):
            ax = K.DeviceArray((n, m))
            i32 = lambda v_: C.byref(C.c_int32(v_))
            lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(v.ptr), C.c_void_p(ax.ptr))
            lib.diaglib_b200_sync()
            K.timer_start()
            for _ in range(reps):
                lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(v.ptr), C.c_void_p(ax.ptr))
            ms = K.timer_stop_ms() / reps
            by = 12.0 * nnz + 8.0 * (n + 1) + 16.0 * n * m
            out["spmm_m%d" % m] = dict(ms=round(ms, 4), gbs=round(by / ms / 1e6, 1), bytes=by)
            ax.free()
        print(json.dumps(dict(n=n, **out)))
        return
    if only == "spmm_toy":
        # long rows (~2 log2 n entries): the generic / pipelined kernels
        import ctypes as C
        from diaglib_b200 import problems as P
        csr = P.toy_sparse(n)
        D.set_csr(*csr)
        nnz = len(csr[1])
        for m in (37, 13, 8):
            ax = K.DeviceArray((n, m))
            i32 = lambda v_: C.byref(C.c_int32(v_))
            lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(v.ptr), C.c_void_p(ax.ptr))
            lib.diaglib_b200_sync()
            K.timer_start()
            for _ in range(reps):
                lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(v.ptr), C.c_void_p(ax.ptr))
            ms = K.timer_stop_ms() / reps
            by = 12.0 * nnz + 8.0 * (n + 1) + 16.0 * n * m
            out["spmm_toy_m%d" % m] = dict(ms=round(ms, 4), gbs=round(by / ms / 1e6, 1), bytes=by, nnz=nnz)
            ax.free()
        print(json.dumps(dict(n=n, **out)))
        return
    if only == "trmm":
        # the in-place triangular multiply of ortho_cd (diaglib.f90:3327) on a 37-column block
        tm = np.asfortranarray(np.triu(np.random.default_rng(1).standard_normal((37, 37)) * 0.05 + np.eye(37)))
        td = K.DeviceArray.from_numpy(tm)
        lib.diaglib_b200_k_trmm(n, v.ptr, n, 37, td.ptr)
        lib.diaglib_b200_sync()
        K.timer_start()
        for _ in range(reps):
            lib.diaglib_b200_k_trmm(n, v.ptr, n, 37, td.ptr)
        ms = K.timer_stop_ms() / reps
        out["trmm_37_in_place"] = dict(ms=round(ms, 4), gbs=round(16.0 * n * 37 / ms / 1e6, 1))
        tds = {}
        for m in (37, 24, 16):
            tm2 = np.asfortranarray(np.triu(np.random.default_rng(1).standard_normal((m, m)) * 0.05 + np.eye(m)))
            tds[m] = K.DeviceArray.from_numpy(tm2)
            for mode in ("in_place", "out_of_place"):
                def go():
                    if mode == "in_place":
                        lib.diaglib_b200_k_trmm(n, v.ptr, n, m, tds[m].ptr)
                    else:
                        lib.diaglib_b200_k_trmm_oop(n, v.ptr, n, m, tds[m].ptr, w.ptr, n)
                go()
                lib.diaglib_b200_sync()
                K.timer_start()
                for _ in range(reps):
                    go()
                ms = K.timer_stop_ms() / reps
                out[f"trmm_{m}_{mode}"] = dict(ms=round(ms, 4), gbs=round(16.0 * n * m / ms / 1e6, 1))
        # a plain device copy of the same block for reference (read + write, different buffers / same buffer region)
        lib.diaglib_b200_d2d(w.ptr, v.ptr, 8 * n * 37)
        K.timer_start()
        for _ in range(reps):
            lib.diaglib_b200_d2d(w.ptr, v.ptr, 8 * n * 37)
        ms = K.timer_stop_ms() / reps
        out["memcpy_d2d_37cols"] = dict(ms=round(ms, 4), gbs=round(16.0 * n * 37 / ms / 1e6, 1))
        print(json.dumps(dict(n=n, **out)))
        return
    if only == "ortho":
        # the two wide kernels of an ortho_vs_x sweep at the C3 shape: xu = x^T u (3543) and u -= x xu (3544) with
        # x = the first 74 columns and u the 37 columns behind them, as in the solver
        t_gram(74, 37, 0, v, w, "gram_74x37")
        xu = K.DeviceArray.from_numpy(np.asfortranarray(np.random.default_rng(3).standard_normal((74, 37)) * 1e-3))
        u_ptr = v.col_ptr(74)
        lib.diaglib_b200_k_project_out(n, 74, 37, v.ptr, n, xu.ptr, u_ptr, n)
        lib.diaglib_b200_sync()
        K.timer_start()
        for _ in range(reps):
            lib.diaglib_b200_k_project_out(n, 74, 37, v.ptr, n, xu.ptr, u_ptr, n)
        ms = K.timer_stop_ms() / reps
        out["project_out_74_37"] = dict(ms=round(ms, 4), gbs=round(8.0 * n * (111 + 37) / ms / 1e6, 1))
        print(json.dumps(dict(n=n, **out)))
        return
    if only == "gram":
        t_gram(111, 111, 1, v, w, "gram_sym_111")
        t_gram(74, 37, 0, v, w, "gram_74x37")
        print(json.dumps(dict(n=n, **out)))
        return
    t_gram(111, 111, 1, v, w, "gram_sym_111")
    t_gram(74, 74, 1, v, w, "gram_sym_74")
    t_gram(37, 37, 1, v, v, "gram_sym_37_same")
    t_gram(74, 37, 0, v, w, "gram_74x37")
    t_gram(111, 37, 0, v, w, "gram_111x37")
    t_bmul(111, 37, "bmul_111x37")
    t_bmul(111, 74, "bmul_111x74")
    t_bmul(74, 37, "bmul_74x37")
    t_bmul(37, 37, "bmul_37x37")
    print(json.dumps(dict(n=n, **out)))


if __name__ == "__main__":
    main()
