import sys, json, ctypes as C, numpy as np
sys.path.insert(0, '.')
import diaglib_b200 as D
from diaglib_b200 import kernels as K, problems as P
n = 1 << 22
D.init(0)
csr = P.fci_like(n)
D.set_csr(*csr)
nnz = len(csr[1])
lib = D.lib()
out = {}
i32 = lambda v_: C.byref(C.c_int32(v_))
for m in (21, 16, 8):
    x = K.DeviceArray((n, m)); lib.diaglib_b200_k_fill_uniform(x.ptr, n, m, n, 1)
    ax = K.DeviceArray((n, m))
    lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(x.ptr), C.c_void_p(ax.ptr)); lib.diaglib_b200_sync()
    K.timer_start()
    for _ in range(5):
        lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(x.ptr), C.c_void_p(ax.ptr))
    ms = K.timer_stop_ms() / 5
    by = 12.0 * nnz + 8.0 * (n + 1) + 16.0 * n * m
    out["m%d" % m] = dict(ms=round(ms, 3), gbs=round(by / ms / 1e6, 1))
    x.free(); ax.free()
print(json.dumps(out))
