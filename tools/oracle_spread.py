"""How far does the ORACLE's own iteration count move on the benchmark workload (C3) when only
the BLAS thread count or the last bit of the start vectors changes?  The iteration count of a
solve that stops on max|r| < 10 tol is decided by residuals within a few per cent of the
threshold, so rounding-level perturbations shift it; the spread measured here is the resolution
of the "+-1 iteration" parity bar on this problem.
usage: python tools/oracle_spread.py NX THREADS SCALE_MINUS_1 OUT.json [accurate_eig=0|1]
(accurate_eig=1: DIAGNOSTIC mode of the oracle, dpotrf + dgesvj instead of dsyev for the reduced problems)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from diaglib_b200 import problems as P  # noqa: E402
from oracle import oracle as O  # noqa: E402

nx, threads, dscale, out = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
n = nx ** 3
n_max = P.n_eig_rule(B.N_TARG)
acc = len(sys.argv) > 5 and sys.argv[5] == "1"
O.set_threads(threads)
O.set_accurate_eig(acc)
csr = P.lap3d(nx, nx, nx, delta=B.DELTA)
O.set_csr(*csr)
g = B.make_guess(csr[3], n, n_max, 0, n)
if dscale != 0.0:
    g *= (1.0 + dscale)
t0 = time.time()
r = O.lobpcg(g, B.N_TARG, B.MAX_ITER, B.TOL)
res = {"nx": nx, "accurate_eig": acc, "threads": O.get_threads(), "guess_scale_minus_1": dscale, "iterations": int(len(r["it"])), "ok": bool(r["ok"]),
       "wall_s": time.time() - t0, "n_act": [int(x) for x in r["n_act"]], "eig": [float(x) for x in r["eig"]],
       "hist_rms_max": [float(x[:B.N_TARG].max()) for x in r["rms"]], "hist_max_max": [float(x[:B.N_TARG].max()) for x in r["max"]]}
json.dump(res, open(out, "w"))
print(json.dumps({k: res[k] for k in ("nx", "accurate_eig", "threads", "guess_scale_minus_1", "iterations", "wall_s")}))
