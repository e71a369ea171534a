"""ncu target: the built-in matvec on lap3d 256^3 once per case: tiled order m=37, m=32; natural order
m=37 (two launches: 24 + 13 columns)."""
import ctypes as C
import sys

sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K, problems as P

nx = 256
n = nx ** 3
D.init(0)
lib = D.lib()
D.set_csr(*P.lap3d(nx, nx, nx, delta=1.0))
i32 = lambda v: C.byref(C.c_int32(int(v)))  # noqa: E731
x, y = K.DeviceArray((n, 37)), K.DeviceArray((n, 37))
lib.diaglib_b200_k_fill_uniform(x.ptr, n, 37, n, 1)
D.set_csr_row_order(P.tile_order_3d(nx, nx, nx, tile=(64, 2, 2)))
for m in (37, 32):
    lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(x.ptr), C.c_void_p(y.ptr))
lib.diaglib_b200_sync()
D.set_csr_row_order(None)
lib.diaglib_b200_csr_matvec(i32(n), i32(37), C.c_void_p(x.ptr), C.c_void_p(y.ptr))
lib.diaglib_b200_sync()
print("done")
