"""SpMM on the C3 stencil (lap3d nx^3) under different processing orders of the rows
(diaglib_b200.set_csr_row_order): natural order against grid tiles along a z-order curve.
Prints ms per call and the fraction of the HBM roofline on algorithmic bytes."""
import ctypes as C
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import diaglib_b200 as D
from diaglib_b200 import kernels as K, problems as P

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = nx ** 3
D.init(0)
lib = D.lib()
csr = P.lap3d(nx, nx, nx, delta=1.0)
D.set_csr(*csr)
nnz = len(csr[1])
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6543.7) if __import__("os").path.exists("MEASURED_PEAKS.json") else 6543.7
i32 = lambda v: C.byref(C.c_int32(int(v)))  # noqa: E731
mmax = 37
x = K.DeviceArray((n, mmax))
y = K.DeviceArray((n, mmax))
lib.diaglib_b200_k_fill_uniform(x.ptr, n, mmax, n, 1)
out = []
# usage: spmm_order_bench.py NX [label] [tile specs like 64x2x2 ...]   (default: a sweep of shapes)
label = sys.argv[2] if len(sys.argv) > 2 else ""
orders = [("natural", None)]
if len(sys.argv) > 3:
    for spec in sys.argv[3:]:
        tile = tuple(int(v) for v in spec.split("x"))
        orders.append((f"tile{tile}-morton", (tile, "morton")))
else:
    for tile in [(32, 4, 2), (16, 4, 4), (32, 8, 1), (32, 2, 4), (64, 2, 2), (32, 4, 2)]:
        for curve in ("morton", "sweep"):
            orders.append((f"tile{tile}-{curve}", (tile, curve)))
seen = set()
for name, spec in orders:
    if name in seen:
        continue
    seen.add(name)
    D.set_csr_row_order(None if spec is None else P.tile_order_3d(nx, nx, nx, tile=spec[0], curve=spec[1]))
    for m in (37, 32, 8):
        for _ in range(3):
            lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(x.ptr), C.c_void_p(y.ptr))
        lib.diaglib_b200_sync()
        K.timer_start()
        reps = 10
        for _ in range(reps):
            lib.diaglib_b200_csr_matvec(i32(n), i32(m), C.c_void_p(x.ptr), C.c_void_p(y.ptr))
        ms = K.timer_stop_ms() / reps
        b = 12.0 * nnz + 8.0 * (n + 1) + 16.0 * n * m
        rec = {"label": label, "order": name, "m": m, "ms": round(ms, 4), "gbs": round(b / ms / 1e6, 1), "frac_hbm": round(b / ms / 1e6 / peak, 4)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
json.dump(out, open(f"gpurun_out/spmm_order_bench{('_' + label) if label else ''}.json", "w"), indent=1)
